"""Device timeline of the sub-batch pipeline: start / end of every stage of every sub-batch of one call (ms since the call began)."""
import argparse
import sys

import torch

sys.path.insert(0, '/root/repo')
from locomouse_cpp_b200 import synth  # noqa: E402
from locomouse_cpp_b200.api import Detector  # noqa: E402
from locomouse_cpp_b200.types import Results  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=5120)
ap.add_argument("--streams", type=int, nargs="+", default=[2, 1])
ap.add_argument("--prio", type=int, default=1)
ap.add_argument("--skip", type=int, default=0)
args = ap.parse_args()
spec = synth.SynthSpec()
cfg, model, bkg, calib, _, _, _, _ = synth.make_problem(spec, 8, seed=1000)
N = args.frames
frames, bx, bs, bb = synth.make_video(spec, N, 1000, "cuda", bkg)
torch.cuda.synchronize()
det = Detector(cfg, model, bkg, calib)
res = Results(N, cfg.cand_cap, cfg.match_cap, cfg.n_tail_points, pinned=True)
names = ["start", "mm", "prep", "scr0", "screen", "corr", "tail", "nms", "pair", "d2h"]
import os
os.environ["LM_WHATIF_SKIP"] = str(args.skip)
for streams in args.streams:
    det.set_option("streams", streams)
    det.set_option("screen_priority", args.prio)
    for _ in range(3):
        det.detect_batch(frames, bx, bs, bb, results=res, allow_overflow=True)
    sub = int(det.info("subbatch"))
    nsub = (N + sub - 1) // sub
    print(f"streams={streams} prio={args.prio} total {det.last_timing()[0]['total']:.3f} ms, {nsub} sub-batches of {sub}")
    for i in range(nsub):
        t = [det.info(f"tl_{i}_{k}") for k in range(10)]
        order = [t[0], t[1], t[2], t[9], t[8], t[3], t[4], t[5], t[6], t[7]]
        print(f"  #{i:2d}: " + "  ".join(f"{n}={v:7.3f}" for n, v in zip(names, order)))
