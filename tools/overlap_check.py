"""Device time of the whole path for combinations of the library options (sub-batch size, stream overlap, screen mode)."""
import sys
import time

import torch

sys.path.insert(0, '/root/repo')
from locomouse_cpp_b200 import synth  # noqa: E402
from locomouse_cpp_b200.api import Detector  # noqa: E402
from locomouse_cpp_b200.types import Results  # noqa: E402

spec = synth.SynthSpec()
cfg, model, bkg, calib, _, _, _, _ = synth.make_problem(spec, 8, seed=1000)
N = 5120
frames, bx, bs, bb = synth.make_video(spec, N, 1000, "cuda", bkg)
torch.cuda.synchronize()
det = Detector(cfg, model, bkg, calib)
res = Results(N, cfg.cand_cap, cfg.match_cap, cfg.n_tail_points)
for screen in (2, 1):
    for sub in (128, 256, 512, 1024):
        for streams in (1, 2):
            det.set_option("screen", screen)
            det.set_option("subbatch", sub)
            det.set_option("streams", streams)
            for _ in range(2):
                det.detect_batch(frames, bx, bs, bb, results=res)
            t = time.perf_counter()
            for _ in range(3):
                det.detect_batch(frames, bx, bs, bb, results=res)
            dt = (time.perf_counter() - t) / 3
            tm, nl = det.last_timing()
            print(f"screen={screen} subbatch={sub:4d} streams={streams}: wall {dt * 1e3:6.2f} ms, device {tm['total']:6.2f} ms -> "
                  f"{N / tm['total'] * 1e3:8.0f} frames/s (device), checksum {res.checksum()}")
