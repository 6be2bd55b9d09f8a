import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, pytest
from locomouse_cpp_b200 import synth
from locomouse_cpp_b200.types import diff_results
from locomouse_cpp_b200.api import Detector
from oracle import oracle
bad = 0
for seed in range(7, 47):
    rng = np.random.Generator(np.random.PCG64(9000 + seed))
    side_h = int(rng.choice([60, 96, 140, 165, 300])); bottom_h = int(rng.choice([70, 120, 235, 280])); bb_w = int(rng.choice([150, 250, 400, 430]))
    n_cols = int(bb_w * rng.uniform(1.6, 3.0)) & ~3
    tsh = lambda: (int(rng.integers(8, 31)), int(rng.integers(8, 31)))
    shapes = tuple(tuple(tsh() for _ in range(3)) for _ in range(2))
    method = str(rng.choice(["TM", "TM_DE", "base"]))
    spec = synth.SynthSpec(method=method, n_rows=side_h + bottom_h, n_cols=n_cols, side_h=side_h, bb_w=bb_w, bb_h_side_tm=max(40, side_h - 15), tshapes=shapes,
                           mouse_scale=min(1.0, bb_w / 400, bottom_h / 235, side_h / 165), flip=bool(rng.integers(0, 2)), cand_cap=128, match_cap=1024,
                           warp=bool(rng.integers(0, 2)), conn=int(rng.choice([4, 8])))
    n = int(rng.integers(5, 12))
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, n, seed=1000 + seed)
    frames = frames.numpy()
    ref = oracle.detect(cfg, model, bkg, calib, frames, bx, bs, bb, n_threads=8)
    for layout in (3, 1, 2):
        det = Detector(cfg, model, bkg, calib, device=0)
        det.set_option("screen_layout", layout); det.set_option("subbatch", int(rng.integers(2, 8))); det.set_option("streams", int(rng.integers(1, 5)))
        got = det.detect_batch(frames, bx, bs, bb, allow_overflow=True)
        info = (det.info("screen_active"), det.info("screen2_merged"), det.info("screen2_stacked"))
        det.close()
        d = diff_results(got, ref)
        if d:
            bad += 1
            print("MISMATCH", seed, layout, info, side_h, bottom_h, bb_w, shapes, d[:3])
    print(seed, "ok", info, int(ref.n_bottom.sum()), flush=True)
print("bad", bad)
