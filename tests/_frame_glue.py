"""Shared by tests/test_oracle_vs_reference.py and tests/golden/make_reference_frame_golden.py: one frame of the bottom / side
candidate detection through the REFERENCE's own detectBottomCandidates / detectSideCandidates / nmsMax / peakClustering
(oracle/_ref), fed with the oracle's pre-processed image, its correlation maps (each pinned against cv2.filter2D
elsewhere) and the reference's own TAIL_MASK."""
import numpy as np

from locomouse_cpp_b200 import synth
from oracle import reference_nms as ref


def problem(n=6, seed=1000, **kw):
    spec = synth.SynthSpec(**kw)
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, n, seed=seed)
    return cfg, model, bkg, calib, frames.numpy(), bx, bs, bb


def reference_frame(oracle, cfg, model, bkg, calib, frame, bx, bs, bb):
    I, _ = oracle.preprocess(cfg, bkg, calib, frame)
    W, hb, hs, tw = cfg.bb_w, cfg.bb_h_bottom, cfg.bb_h_side, cfg.tail_w
    x0, y0b, y0s = int(bx) - W + 1, int(bb) - hb + 1, int(bs) - hs + 1
    big = 2 * max(W, hb, hs)
    canvas = np.zeros((I.shape[0] + 2 * big, I.shape[1] + 2 * big), np.uint8)
    canvas[big:big + I.shape[0], big:big + I.shape[1]] = I
    crop = lambda y0, h: canvas[big + y0: big + y0 + h, big + x0: big + x0 + W]
    corr = lambda v, f, width, y0, h: oracle.correlate(I, model.w[v][f], model.rho[v][f], x0, y0, width, h, fma_mode=bool(cfg.fma_mode))
    # TAIL_MASK from the reference's own tail code on the tail score maps (left tail_w columns of each crop)
    _, tail_mask = ref.detect_tail(corr(0, 2, tw, y0b, hb), corr(1, 2, tw, y0s, hs), cfg.conn, cfg.n_tail_points)
    pad = 17   # any padding: the reference only reads the unpadded window of the filter output (class.cpp:848, 863)
    maps = []
    for v, (y0, h) in enumerate(((y0b, hb), (y0s, hs))):
        for f in range(2):
            m = np.zeros((h + 2 * pad, W + 2 * pad), np.float32)
            m[pad:pad + h, pad:pad + W] = corr(v, f, W, y0, h)
            maps.append(m)
    tsz = []
    for f in range(2):
        for v in range(2):
            tsz += [model.w[v][f].shape[1], model.w[v][f].shape[0]]
    return ref.detect_candidates(crop(y0b, hb), crop(y0s, hs), tail_mask, maps, (W + 2 * pad, hb + 2 * pad), (W + 2 * pad, hs + 2 * pad),
                                 (pad, pad), (pad, pad), tsz)


CASES = (dict(), dict(method="TM_DE", flip=True), dict(method="base", fma_mode=False, conn=4), dict(q6=True))


def case_problem(kw, n=4):
    """q6: the bottom paw bias is raised until no bottom paw candidate survives, so the reference must skip the side paw
    detection although the side scores are positive (class.cpp:820-833)."""
    from locomouse_cpp_b200.types import Model

    kw = dict(kw)
    q6 = kw.pop("q6", False)
    cfg, model, bkg, calib, frames, bx, bs, bb = problem(n, 1000, **kw)
    if q6:
        rho = [list(r) for r in model.rho]
        rho[0][0] += 1.0e3
        model = Model(w=model.w, rho=rho)
    return cfg, model, bkg, calib, frames, bx, bs, bb
