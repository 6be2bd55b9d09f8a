"""SURVEY config 5 (2x upsampled 800x3400 frames, 60x60 templates): dense exact kernel vs the tensor-core screen.
Checks bit-equality of the three modes on device-rendered frames and prints per-stage device times."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from locomouse_cpp_b200 import synth  # noqa: E402
from locomouse_cpp_b200.api import Detector  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=128)
ap.add_argument("--subbatch", type=int, default=64)
args = ap.parse_args()
spec = synth.SynthSpec(scale=2, cand_cap=128, match_cap=512)
cfg, model, bkg, calib, _, _, _, _ = synth.make_problem(spec, 4, seed=1000)
frames, bx, bs, bb = synth.make_video(spec, args.frames, 1000, "cuda", bkg)
torch.cuda.synchronize()
sums = {}
for screen in (2, 0):
    det = Detector(cfg, model, bkg, calib)
    det.set_option("screen", screen)
    det.set_option("subbatch", args.subbatch)
    det.set_option("streams", 1)
    for _ in range(3):
        r = det.detect_batch(frames, bx, bs, bb, allow_overflow=True)
    tm, nl = det.last_timing()
    print(f"screen={screen} active={det.info('screen_active')} frames={args.frames}: "
          + " ".join(f"{k}={v:.2f}" for k, v in tm.items()) + f" ms -> {args.frames / tm['total'] * 1e3:.0f} frames/s (device)")
    sums[screen] = r.checksum()
    det.close()
print("identical results:", len(set(sums.values())) == 1, sums)
