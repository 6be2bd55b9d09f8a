"""Pass 1 of LocoMouse_TM_DE: the oracle's restatement of computeMouseBox_DE / imadjust_default /
firstLastOverT / vecmovingaverage (LocoMouse_TM_DE.cpp:8-113, LocoMouse_class.cpp:3244-3311, 1559-1608,
LocoMouse_class.hpp:411-442) against OpenCV 4.13 primitives (cv2) and hand-derived known answers."""
import math

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from locomouse_cpp_b200 import synth
from locomouse_cpp_b200.types import bb_de_params


# ---- literal Python transcriptions used as the second opinion ---------------------------------------------------
def py_imadjust_default(img):
    """imadjust_default with cv2.calcHist for the histogram and cv2.convertScaleAbs for the MatExpr (the same
    cvtScale float path as Mat::convertTo; identical wherever src*alpha+beta >= 0, 0 below)."""
    hist = cv2.calcHist([img], [0], None, [256], [0, 256]).reshape(-1)  # float32 counts
    total = np.float32(np.float64(hist.astype(np.float64).sum()))
    cum = np.float32(0)
    i0 = i1 = imin = imax = 0
    cmin = cmax = True
    for i in range(256):
        cum = np.float32(cum + hist[i])
        cn = np.float32(cum / total)
        if cn > np.float32(0.01) and cmin:
            i0, cmin, imin = i, False, i
        if cn >= np.float32(0.99) and cmax:
            i1, cmax, imax = i, False, i
        if not (cmin or cmax):
            break
    if imin == imax:
        i1 = 256
    r0 = np.float32(np.float32(i0) / np.float32(255))
    r1 = np.float32(np.float32(i1) / np.float32(255))
    s = float(np.float32(r1 - r0))
    alpha = 1.0 / s
    beta = -float(r0) * alpha
    if abs(alpha) == 1.0:
        return img.copy(), (i0, i1)
    out = cv2.convertScaleAbs(img, alpha=alpha, beta=beta)
    neg = (img.astype(np.float64) * np.float32(alpha) + np.float32(beta)) < -0.5   # saturate_cast clips what abs would mirror
    out[neg] = 0
    return out, (i0, i1)


def py_first_last(vals, th):
    fl = [0, 0]
    idx, has = 0, False
    for i, v in enumerate(vals):
        if v >= th:
            fl[idx] = i
            if not has:
                idx, has = 1, True
    return tuple(fl) if has else (-1, -1)


def py_movavg(v, w):
    u32 = lambda x: int(x) & 0xFFFFFFFF  # noqa: E731
    n = len(v)
    if w >= n:
        return [u32(x) for x in v]
    out = [0] * n
    half = w // 2
    for i in range(half):
        out[i] = u32(v[i])
    cur = 0.0
    for i in range(w):
        cur += v[i]
    out[half] = u32(math.floor(cur / w))
    for i in range(n - w):
        cur = cur - v[i] + v[i + w]
        out[half + 1 + i] = u32(math.floor(cur / w))
    for i in range(n - half - 1, n):
        out[i] = u32(v[i])
    return out


# ---- tests --------------------------------------------------------------------------------------------------------
def test_imadjust_default_lut_matches_cv2(oracle):
    rng = np.random.Generator(np.random.PCG64(5))
    for it in range(30):
        kind = it % 5
        if kind == 0:
            img = rng.integers(0, 256, (165, 1700), dtype=np.uint8)
        elif kind == 1:   # dark image with a bright object (the usual side view)
            img = np.clip(rng.normal(8, 4, (165, 1700)), 0, 255).astype(np.uint8)
            img[100:140, 300:600] = rng.integers(120, 256, (40, 300), dtype=np.uint8)
        elif kind == 2:   # narrow band
            img = rng.integers(90, 110, (64, 512), dtype=np.uint8)
        elif kind == 3:   # constant image: imin == imax -> indices[1] = 256
            img = np.full((50, 300), int(rng.integers(0, 256)), np.uint8)
        else:             # more than 1 % zeros and more than 1 % at 255: imin = 0, imax = 255 -> identity branch
            img = rng.integers(0, 256, (100, 800), dtype=np.uint8)
            img[:10] = 0
            img[-10:] = 255
        hist = np.bincount(img.reshape(-1), minlength=256).astype(np.uint32)
        assert np.array_equal(hist, cv2.calcHist([img], [0], None, [256], [0, 256]).reshape(-1).astype(np.uint32))
        lut, idx = oracle.imadjust_default_lut(hist)
        want, idx2 = py_imadjust_default(img)
        assert idx == idx2
        assert np.array_equal(lut[img], want), f"case {it}: imadjust_default mapping differs from cv2"


def test_first_last_over_t_known_answers(oracle):
    f = oracle.first_last_over_t
    assert f([0, 0, 0, 0], 1) == (-1, -1)
    assert f([0, 5, 0, 0], 5) == (1, 0)              # a single qualifying column leaves slot 1 at its initial 0
    assert f([9, 10, 11, 3, 10, 2], 10) == (1, 4)
    assert f([10, 0, 0], 10) == (0, 0)
    assert f([0, 0, 12, 12], 10) == (2, 3)
    assert f([3, 3, 3], 0) == (0, 2)                 # threshold 0: every column qualifies
    rng = np.random.Generator(np.random.PCG64(1))
    for _ in range(50):
        v = rng.integers(0, 50, int(rng.integers(1, 200))).astype(np.float32)
        th = int(rng.integers(0, 55))
        assert f(v, th) == py_first_last(v, th)


def test_vecmovingaverage_known_answers(oracle):
    mv = oracle.vecmovingaverage
    assert mv([3.7, 4.2], 5).tolist() == [3, 4]                       # window >= n: truncated copy
    assert mv([1, 2, 3, 4, 5, 6, 7], 5).tolist() == [1, 2, 3, 4, 5, 6, 7]
    # floor(10.9 / 5) = 2 at index 2; the reference's tail loop starts at n - half - 1 and so overwrites the last averaged
    # value (index 4) with the raw one (LocoMouse_class.cpp:1603)
    assert mv([10.9, 0, 0, 0, 0, 0, 10.9], 5).tolist() == [10, 0, 2, 0, 0, 0, 10]
    assert mv([5.5] * 9, 3).tolist() == [5] * 9
    rng = np.random.Generator(np.random.PCG64(2))
    for _ in range(40):
        n, w = int(rng.integers(1, 60)), int(rng.choice([1, 3, 5, 7, 9]))
        v = rng.uniform(0, 1800, n)
        assert mv(v, w).tolist() == py_movavg(list(v), w)
    # no mouse in a frame: bb_x = -1.1 -> (uint32_t)(-1.1) through int64 = 0xFFFFFFFF
    assert mv([-1.1, 50.0], 5).tolist() == [0xFFFFFFFF, 50]


@pytest.mark.parametrize("flip,warp", [(False, False), (True, True)])
def test_bounding_box_tm_de_matches_cv2_reenactment(oracle, flip, warp):
    spec = synth.SynthSpec(method="TM_DE", flip=flip, warp=warp, vid_pad=3 if warp else 0)
    n = 9
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, n, seed=1000)
    frames = frames.numpy()
    p = bb_de_params(cfg, side_h=spec.side_h)
    got, raw, lims = oracle.bounding_box_tm_de(cfg, bkg, calib, frames, p)
    import dataclasses

    base = dataclasses.replace(cfg, imadjust=False)
    want_raw, want_lims = [], []
    for f in range(n):
        I, _ = oracle.preprocess(base, bkg, calib, frames[f])          # base readFrame (pinned vs cv2 elsewhere)
        side = np.ascontiguousarray(I[: spec.side_h, :])
        side, _ = py_imadjust_default(side)
        side[:, :46] = 0
        side[:, 760:] = 0
        side[:100] = 0
        side[149:] = 0
        _, binary = cv2.threshold(side, 255 * 0.05, 1, cv2.THRESH_BINARY)
        colsum = cv2.reduce(binary, 0, cv2.REDUCE_SUM, dtype=cv2.CV_32F).reshape(-1)
        fl = py_first_last(colsum, 10)
        want_lims.append(fl)
        want_raw.append(min(float(cfg.n_cols - 1), float(fl[1]) * 1.1))
    assert lims.tolist() == [list(x) for x in want_lims]
    assert raw.tolist() == want_raw
    assert got.tolist() == py_movavg(want_raw, 5)
    assert (lims[:, 1] > 0).any(), "synthetic side view produced no qualifying column: the test would be vacuous"
