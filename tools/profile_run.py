"""Small driver for ncu: N device-resident synthetic frames through lm_detect_batch, a few times."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from locomouse_cpp_b200 import synth  # noqa: E402
from locomouse_cpp_b200.api import Detector  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=512)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--scale", type=int, default=1)
ap.add_argument("--screen", type=int, default=2)
ap.add_argument("--streams", type=int, default=2)
ap.add_argument("--layout", type=int, default=3)
ap.add_argument("--prio", type=int, default=1)
ap.add_argument("--stages", type=int, default=4)
ap.add_argument("--subbatch", type=int, default=0)
args = ap.parse_args()
spec = synth.SynthSpec(scale=args.scale) if args.scale == 1 else synth.SynthSpec(scale=args.scale, cand_cap=128, match_cap=512)
cfg, model, bkg, calib, _, _, _, _ = synth.make_problem(spec, 8, seed=1000)
frames, bx, bs, bb = synth.make_video(spec, args.frames, 1000, "cuda", bkg)
torch.cuda.synchronize()
det = Detector(cfg, model, bkg, calib)
det.set_option("screen", args.screen)
det.set_option("streams", args.streams)
det.set_option("screen_layout", args.layout)
det.set_option("screen_priority", args.prio)
det.set_option("screen_stages", args.stages)
if args.subbatch:
    det.set_option("subbatch", args.subbatch)
for _ in range(args.iters):
    r = det.detect_batch(frames, bx, bs, bb, allow_overflow=True)
    print(det.last_timing())
print("screen_active", det.info("screen_active"), "merged", det.info("screen2_merged"), "stacked", det.info("screen2_stacked"), "ms_screen", det.info("ms_screen"), "checksum", r.checksum())
print("sparse_tasks(set0)", det.info("sparse_tasks"), "positives(set0)", det.info("positives"), "subbatch", det.info("subbatch"))
print("flags", int((r.flags != 0).sum()), "n_bottom", r.n_bottom.mean(0), "n_side", r.n_side.mean(0))
