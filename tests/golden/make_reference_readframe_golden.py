"""Golden checksums of calibrated images produced by the REFERENCE's own readFrame / correctImage code (compiled from
/root/reference, oracle/ref_glue.cpp::ref_read_frame; normalize and flip executed by the real OpenCV) on seeded synthetic
frames, + LocoMouse_TM::readFrame's imadjust (the reference's own table) for the TM variants."""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _frame_glue as G  # noqa: E402
from oracle import reference_nms as ref  # noqa: E402

CASES = (dict(method="base"), dict(method="base", flip=True, warp=True, vid_pad=5), dict(method="TM", warp=True), dict(method="TM_DE", flip=True))


def reference_image(cfg, bkg, calib, frame):
    img = ref.read_frame(frame, bkg, calib, cfg.flip)
    return ref.imadjust_lut()[img] if cfg.imadjust else img


def frames_of(kw):
    cfg, model, bkg, calib, frames, bx, bs, bb = G.problem(3, 1000, **kw)
    return cfg, bkg, calib, list(frames) + [bkg.copy()]   # + a constant frame: smax == smin


def main():
    out = {}
    for ci, kw in enumerate(CASES):
        cfg, bkg, calib, frames = frames_of(kw)
        for f, fr in enumerate(frames):
            img = reference_image(cfg, bkg, calib, fr)
            out[f"c{ci}_f{f}"] = np.array([zlib.crc32(img.tobytes()), int(img.sum()), int(img.max())], np.int64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_readframe.npz")
    np.savez_compressed(path, **out)
    print("images", len(out), {k: v.tolist() for k, v in list(out.items())[:2]})


if __name__ == "__main__":
    main()
