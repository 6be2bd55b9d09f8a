"""Golden bb_x values produced by the REFERENCE's own LocoMouse_TM_DE::computeMouseBox_DE + LocoMouse::imadjust_default code
(compiled from /root/reference, oracle/ref_glue.cpp::ref_mouse_box_de) on the calibrated side views that its own readFrame
code (ref_read_frame) produced from seeded synthetic frames; OpenCV's normalize / flip / scaled 8-bit conversion run in the
real OpenCV."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _frame_glue as G  # noqa: E402
from oracle import reference_nms as ref  # noqa: E402

CASES = (dict(method="TM_DE"), dict(method="TM_DE", flip=True, warp=True), dict(method="TM_DE", mouse_scale=0.6))
SIDE_H = 165


def frames_of(kw, n=8):
    cfg, model, bkg, calib, frames, bx, bs, bb = G.problem(n, 1000, **kw)
    extra = [bkg.copy(), np.clip(bkg.astype(int) + 3, 0, 255).astype(np.uint8)]   # constant / near-constant frames
    return cfg, bkg, calib, np.stack(list(frames) + extra)


def main():
    out = {}
    for ci, kw in enumerate(CASES):
        cfg, bkg, calib, frames = frames_of(kw)
        out[f"c{ci}"] = np.array([ref.mouse_box_de(ref.read_frame(fr, bkg, calib, cfg.flip)[:SIDE_H]) for fr in frames], np.float64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_pass1.npz")
    np.savez_compressed(path, **out)
    print({k: v.tolist() for k, v in out.items()})


if __name__ == "__main__":
    main()
