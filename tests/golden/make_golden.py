"""Generates the committed golden fixtures under tests/golden/.

Run from the repo root:  python tests/golden/make_golden.py

Two kinds of vectors:
 * cv2_primitives.npz — outputs of REAL OpenCV (python cv2 4.13, the only executable part of the
   reference stack in this image) for the primitives the hot path calls, so the oracle stays pinned
   on machines without cv2.
 * detect_*.npz — small end-to-end problems (inputs + frozen oracle outputs).  The reference itself
   cannot be built here (no OpenCV C++ SDK) and ships no fixtures, so these freeze the oracle
   (SURVEY.md §8c item 3); the CUDA path is tested against them through the C ABI.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))

from locomouse_cpp_b200 import synth  # noqa: E402
from locomouse_cpp_b200.types import Results  # noqa: E402
from oracle import oracle  # noqa: E402

SMALL = dict(n_rows=160, n_cols=420, side_h=64, bb_w=128, bb_h_side_tm=56, tsize=12, mouse_scale=0.32, cand_cap=32,
             det_cap=2048, match_cap=128)
CASES = {
    "detect_small_tm": dict(spec=dict(method="TM", tshapes=(((12, 12), (10, 14), (9, 8)), ((11, 12), (12, 10), (8, 9))), **SMALL),
                            n=5, first=0, seed=31),
    "detect_small_tmde_flip_warp": dict(spec=dict(method="TM_DE", flip=True, warp=True, vid_pad=3, conn=4, **SMALL),
                                        n=4, first=3, seed=32),
    "detect_small_base_muladd": dict(spec=dict(method="base", fma_mode=False, **SMALL), n=3, first=0, seed=33),
}


def make_detect_case(name, spec_kw, n, first, seed):
    spec = synth.SynthSpec(**spec_kw)
    total = first + n
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, total, seed=seed)
    frames = frames.numpy()
    prev = frames[first - 1] if first > 0 else None
    fr, bx, bs, bb = frames[first:], bx[first:], bs[first:], bb[first:]
    res = oracle.detect(cfg, model, bkg, calib, fr, bx, bs, bb, prev_frame=prev, first_frame_index=first)
    assert res.rc == 0
    out = dict(spec=json.dumps(spec_kw), first=first, bkg=bkg, calib=calib, frames=fr, bb_x=bx, bb_y_side=bs,
               bb_y_bottom=bb, prev=prev if prev is not None else np.zeros((0, 0), np.uint8),
               rho=np.array(model.rho, np.float64))
    for v in range(2):
        for k in range(3):
            out[f"w_{v}_{k}"] = model.w[v][k]
    for a in Results.ARRAYS:
        out["exp_" + a] = getattr(res, a)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "frames", fr.shape, "n_bottom", res.n_bottom.sum(0), "n_side", res.n_side.sum(0), "matches",
          int(res.match_n.sum()))


def make_cv2_primitives():
    import cv2

    rng = np.random.Generator(np.random.PCG64(4242))
    out = {"cv2_version": np.array(cv2.__version__)}
    # readFrame chain
    vr, vc = 60, 90
    bkg = rng.integers(10, 70, (vr, vc), dtype=np.uint8)
    frame = np.clip(bkg.astype(int) + rng.integers(-12, 190, (vr, vc)), 0, 255).astype(np.uint8)
    calib = rng.integers(0, vr * vc, (50, 80)).astype(np.int32)
    F = cv2.normalize(cv2.subtract(frame, bkg), None, 0, 255, cv2.NORM_MINMAX, cv2.CV_8UC1)
    out.update(pre_bkg=bkg, pre_frame=frame, pre_calib=calib, pre_norm_gather_flip=cv2.flip(F.reshape(-1)[calib], 1))
    # filter2D direct path (< 50 taps): bit-exact target of the mul+add mode
    I = rng.integers(0, 256, (48, 64), dtype=np.uint8)
    pad = 12
    C = np.zeros((48 + 2 * pad, 64 + 2 * pad), np.uint8)
    C[pad:-pad, pad:-pad] = I
    out["f2d_image"] = I
    for i, (kh, kw) in enumerate([(7, 7), (4, 9), (6, 6)]):
        k = rng.normal(0, 0.02, (kh, kw)).astype(np.float32)
        full = cv2.filter2D(C, cv2.CV_32F, k, anchor=(-1, -1), delta=-0.25, borderType=cv2.BORDER_CONSTANT)
        out[f"f2d_k{i}"] = k
        out[f"f2d_out{i}"] = full[pad - 5: pad + 48 + 5, pad - 5: pad + 64 + 5]  # window origin (-5,-5)
    # connected components: largest region incl. ties
    bins, larg = [], []
    for conn in (4, 8):
        for t in range(12):
            b = (rng.random((14, 22)) < rng.choice([0.15, 0.4])).astype(np.uint8)
            n, labels, stats, _ = cv2.connectedComponentsWithStats(b, connectivity=conn, ltype=cv2.CV_16U)
            if n <= 1:
                m = np.zeros_like(b)
            else:
                best = 1 + int(np.argmax(stats[1:, cv2.CC_STAT_AREA]))  # argmax = first maximum = strict '>'
                m = ((labels == best) * 255).astype(np.uint8)
            bins.append(b)
            larg.append(m)
    out["cc_bin"] = np.stack(bins)
    out["cc_largest"] = np.stack(larg)
    np.savez_compressed(os.path.join(OUT, "cv2_primitives.npz"), **out)
    print("cv2_primitives written")


if __name__ == "__main__":
    make_cv2_primitives()
    for name, c in CASES.items():
        make_detect_case(name, c["spec"], c["n"], c["first"], c["seed"])
