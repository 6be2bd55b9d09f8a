"""Timing of lm_bounding_box_tm on resident synthetic frames for several thresholds (clean -> very noisy binary images)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from locomouse_cpp_b200 import synth  # noqa: E402
from locomouse_cpp_b200.api import Detector  # noqa: E402
from locomouse_cpp_b200.types import bb_tm_params  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
spec = synth.SynthSpec()
cfg, model, bkg, calib, _, _, _, _ = synth.make_problem(spec, 8, seed=1000)
frames, bx, bs, bb = synth.make_video(spec, n, 1000, "cuda", bkg)
det = Detector(cfg, model, bkg, calib)
for r in (5, 7):
    y, x = np.mgrid[-r:r + 1, -r:r + 1]
    dk = ((x * x + y * y) <= (r + 0.5) ** 2).astype(np.float64)
    dk = (dk / dk.sum()).astype(np.float32)
    for thr in (80, 40, 20, 10, 3):
        P = bb_tm_params(cfg, dk, side_h=spec.side_h, side_threshold=thr, min_pixel_count=25, sums_as_float=0)
        det.bounding_box_tm(frames[:256], P)
        torch.cuda.synchronize()
        t = time.perf_counter()
        raw, lims = det.bounding_box_tm(frames, P)
        dt = time.perf_counter() - t
        print(f"disk {2 * r + 1}x{2 * r + 1} threshold {thr:3d}: {n / dt / 1e3:8.1f} k frames/s  ({dt * 1e3:.1f} ms for {n} frames)  bb_x median {np.median(raw):.0f}", flush=True)
