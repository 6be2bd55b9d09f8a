"""Which stage bounds the overlapped pipeline?  Runs lm_detect_batch on resident frames with stages left out
(LM_WHATIF_SKIP, timing only: the results of such runs are meaningless) and prints frames/s per variant."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from locomouse_cpp_b200 import synth  # noqa: E402
from locomouse_cpp_b200.api import Detector  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5120
spec = synth.SynthSpec(det_cap=int(os.environ.get("LM_WHATIF_DETCAP", "8192")))
cfg, model, bkg, calib, _, _, _, _ = synth.make_problem(spec, 8, seed=1000)
frames, bx, bs, bb = synth.make_video(spec, n, 1000, "cuda", bkg)
det = Detector(cfg, model, bkg, calib)
from locomouse_cpp_b200.types import Results  # noqa: E402

res = Results(n, cfg.cand_cap, cfg.match_cap, cfg.n_tail_points, pinned=True)  # as bench.py: direct device -> host copies
names = {1: "minmax", 2: "prep", 4: "screen", 8: "sparse", 16: "tail", 32: "nms", 64: "pair"}
variants = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0, 1, 2, 3, 4, 8, 12, 16, 32, 64, 112, 115, 123, 4 | 8 | 16 | 32 | 64]
stream_counts = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [4, 8]
for streams in stream_counts:
    det.set_option("streams", streams)
    os.environ["LM_WHATIF_SKIP"] = "0"
    det.detect_batch(frames, bx, bs, bb, results=res, allow_overflow=True)
    for v in variants:
        os.environ["LM_WHATIF_SKIP"] = str(v)
        det.detect_batch(frames, bx, bs, bb, results=res, allow_overflow=True)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(3):
            det.detect_batch(frames, bx, bs, bb, results=res, allow_overflow=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) / 3
        skipped = "+".join(nm for bit, nm in names.items() if v & bit) or "nothing"
        print(f"streams={streams} skip={v:3d} ({skipped:40s}) {n / dt / 1e3:8.1f} k frames/s  {dt * 1e3 / (n / 512):6.3f} ms per 512 frames", flush=True)
os.environ["LM_WHATIF_SKIP"] = "0"
