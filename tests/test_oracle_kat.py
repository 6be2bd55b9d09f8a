"""Known-answer tests for the oracle's restatement of the reference's OWN loops (nmsMax,
peakClustering, matchViews), hand-derived from the cited reference lines, plus a literal pure-Python
transliteration of the same loops used as an independent second implementation on random inputs.
CPU only.
"""
import math

import numpy as np
import pytest


def _rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def _score_map(h, w, dets):
    m = np.full((h, w), -1.0, np.float32)
    for (x, y, s) in dets:
        m[y, x] = s
    return m


# ---- literal python restatement (small inputs only) ------------------------------------------------
def _collect(scores):
    d = [(x, y, float(scores[y, x])) for y in range(scores.shape[0]) for x in range(scores.shape[1])
         if scores[y, x] > 0]
    return sorted(d, key=lambda t: -t[2])  # python sort is stable -> ties keep row-major order


def _inter(a, b, w, h):
    iw = min(a[0], b[0]) + w - max(a[0], b[0])
    ih = min(a[1], b[1]) + h - max(a[1], b[1])
    return iw * ih if iw > 0 and ih > 0 else 0


def _round_half_even(v):
    return int(np.rint(v))


def py_nms_max(scores, w, h):            # LocoMouse_class.cpp:1610-1747
    det = _collect(scores)
    n = len(det)
    discard = [False] * n
    maxima = [0] * n
    cand = []
    for i in range(n):
        if not discard[i]:
            cand.append(i)
            maxima[i] = i
        for j in range(i + 1, n):
            if discard[j]:
                continue
            ia = _inter(det[i], det[j], w, h)
            if ia == 0:
                continue
            if ia / (2.0 * w * h - ia) > 0.5:
                discard[j] = True
                maxima[j] = maxima[i]
    slot = {c: k for k, c in enumerate(cand)}
    wx = [0.0] * len(cand)
    wy = [0.0] * len(cand)
    ss = [0.0] * len(cand)
    for i in range(n):
        k = slot[maxima[i]]
        wx[k] += det[i][0] * det[i][2]
        wy[k] += det[i][1] * det[i][2]
        ss[k] += det[i][2]
    return [(_round_half_even(wx[k] / ss[k]), _round_half_even(wy[k] / ss[k]), det[c][2]) for k, c in enumerate(cand)]


def _round_half_away(v):
    return int(math.floor(abs(v) + 0.5) * (1 if v >= 0 else -1))


def py_peak_clustering(scores, w, h):    # LocoMouse_class.cpp:1749-1905
    det = _collect(scores)
    n = len(det)
    kp = [False] * n
    out = []
    for i in range(n):
        if kp[i]:
            continue
        cl = [i]
        for j in range(i + 1, n):
            if kp[j]:
                continue
            ia = _inter(det[i], det[j], w, h)
            if ia == 0:
                continue
            if ia / (2.0 * h * w - ia) > 0.5:
                kp[j] = True
                cl.append(j)
        if len(cl) > 1:
            px = py = sm = 0.0
            for k in cl:
                px += det[k][0] * det[k][2]
                py += det[k][1] * det[k][2]
                sm += det[k][2]
            out.append((_round_half_away(px / sm), _round_half_away(py / sm), det[i][2]))
        else:
            out.append(det[i])
    return out


# ---- hand-derived cases ------------------------------------------------------------------------------
def test_overlap_predicate_table(oracle):
    """inter/(2wh-inter) > 0.5  <=>  3(w-|dx|)(h-|dy|) > 2wh.  30x30: dx=10,dy=0 -> 1800 > 1800 false;
    dx=9 -> 1890 true; dx=dy=5 -> 1875 true; dx=dy=6 -> 1728 false."""
    for (dx, dy, expect) in [(10, 0, False), (9, 0, True), (0, 9, True), (0, 10, False), (5, 5, True), (6, 6, False),
                             (9, 1, True), (9, 2, False), (8, 2, True)]:
        m = _score_map(40, 40, [(5, 5, 2.0), (5 + dx, 5 + dy, 1.0)])
        n = len(oracle.nms_max(m, 30, 30))
        assert (n == 1) == expect, (dx, dy)
        assert (len(oracle.peak_clustering(m, 30, 30)) == 1) == expect
        assert (3 * (30 - dx) * (30 - dy) > 2 * 900) == expect


def test_chain_suppression_differs_between_nmsmax_and_peakclustering(oracle):
    """A(0,0,3) B(8,0,2) C(16,0,1), 30x30: A~B and B~C overlap, A~C do not.
    nmsMax: B is discarded by A but STILL suppresses C (no `continue` in the outer loop, Q3) -> one
    candidate, mean over A,B,C;  peakClustering skips clustered B -> C is its own maximum."""
    m = _score_map(10, 40, [(0, 0, 3.0), (8, 0, 2.0), (16, 0, 1.0)])
    a = oracle.nms_max(m, 30, 30)
    assert a == [(5, 0, 3.0)]  # (0*3+8*2+16*1)/6 = 5.33 -> 5
    b = oracle.peak_clustering(m, 30, 30)
    assert b == [(3, 0, 3.0), (16, 0, 1.0)]  # (0*3+8*2)/5 = 3.2 -> 3 ; singleton copied


def test_rounding_half_even_vs_half_away(oracle):
    """Equal scores at x=0 and x=1 -> mean 0.5 : nmsMax (saturate_cast) -> 0, peakClustering (round) -> 1;
    x=1,2 -> 1.5 -> both 2; x=2,3 -> 2.5 : 2 vs 3.  Also fixes the tie order (row-major index)."""
    for x, ev, aw in [(0, 0, 1), (1, 2, 2), (2, 2, 3)]:
        m = _score_map(4, 10, [(x, 1, 1.5), (x + 1, 1, 1.5)])
        assert oracle.nms_max(m, 30, 30) == [(ev, 1, 1.5)]
        assert oracle.peak_clustering(m, 30, 30) == [(aw, 1, 1.5)]


def test_empty_and_single(oracle):
    m = np.zeros((5, 5), np.float32)
    assert oracle.nms_max(m, 30, 30) == [] and oracle.peak_clustering(m, 30, 30) == []
    m[2, 3] = 0.25
    assert oracle.nms_max(m, 30, 30) == [(3, 2, 0.25)] and oracle.peak_clustering(m, 30, 30) == [(3, 2, 0.25)]


@pytest.mark.parametrize("seed", range(12))
def test_nms_random_maps_vs_python_restatement(oracle, seed):
    rng = _rng(seed)
    h, w = int(rng.integers(20, 60)), int(rng.integers(30, 90))
    bw, bh = int(rng.integers(6, 31)), int(rng.integers(6, 31))
    m = rng.normal(-1.2, 1.0, (h, w)).astype(np.float32)
    # blobs with plateaus so equal scores (ties) occur
    for _ in range(int(rng.integers(1, 6))):
        cx, cy = int(rng.integers(0, w)), int(rng.integers(0, h))
        m[max(0, cy - 2): cy + 3, max(0, cx - 3): cx + 4] = np.float32(rng.uniform(0.5, 3))
    assert oracle.nms_max(m, bw, bh) == py_nms_max(m, bw, bh)
    assert oracle.peak_clustering(m, bw, bh) == py_peak_clustering(m, bw, bh)


# ---- matchViews ----------------------------------------------------------------------------------------
def py_match_views(cb, cs, vel_check, tb, ts, T, I, Ip, x0, y0b, y0s):   # LocoMouse_class.cpp:1023-1254
    ovlp = int(tb[0] * (1 - T))
    nb, ns = len(cb), len(cs)
    out = []
    if nb == 0:
        return out
    if ns:
        D = [[abs(cb[i][0] - cs[j][0]) for j in range(ns)] for i in range(nb)]
        B = [[255 if D[i][j] <= ovlp else 0 for j in range(ns)] for i in range(nb)]
        flat = [v for r in B for v in r]
        B = [[(1 if v else 0) if max(flat) > min(flat) else 0 for v in r] for r in B]
        Wt = [[D[i][j] * (-(1.0 / ovlp)) + 1.0 for j in range(ns)] for i in range(nb)]
        col = [sum(B[i][j] for i in range(nb)) for j in range(ns)]
        row = [sum(B[i]) for i in range(nb)]

    def box(t):
        w = int(math.floor(t[0] / 2 + 0.5))
        h = int(math.floor(t[1] / 2 + 0.5))
        return -(w // 2), -(h // 2), w, h

    def px(A, x, y):
        return int(A[y, x]) if 0 <= x < A.shape[1] and 0 <= y < A.shape[0] else 0

    def vel(x, y, bx, area, alpha):
        cnt = 0
        for r in range(bx[3]):
            for c in range(bx[2]):
                d = px(I, x + bx[0] + c, y + bx[1] + r) - px(Ip, x + bx[0] + c, y + bx[1] + r)
                cnt += 1 if max(d, 0) > 25 else 0
        return cnt >= area * alpha

    need_t, mov_t = [True] * ns, [False] * ns
    for i in range(nb):
        lst = []
        if ns and row[i] != 0:
            need_b, mov_b = True, False
            for j in range(ns):
                if B[i][j] < 1:
                    continue
                match = True
                if (col[j] > 1) and vel_check:
                    if need_b:
                        mov_b = vel(x0 + cb[i][0], y0b + cb[i][1], box(tb), tb[0] * tb[1], 0.02)
                        need_b = False
                    if need_t[j]:
                        mov_t[j] = vel(x0 + cs[j][0], y0s + cs[j][1], box(ts), ts[0] * ts[1], 0.05)
                        need_t[j] = False
                    match = mov_b == mov_t[j]
                if match:
                    lst.append((cs[j][1], cs[j][2] * Wt[i][j]))
        out.append(lst)
    return out


def test_match_views_hand_cases(oracle):
    tb = ts = (30, 30)     # ovlp = int(30 * (1 - 0.7)) = 9
    # no side candidates -> every bottom candidate unmatched (sentinel)
    assert oracle.match_views([(10, 5, 1.0), (50, 6, 2.0)], [], False, tb, ts, 0.7) == [[], []]
    # Q7: 1 x 1 within overlap -> boolD all 255 -> normalised to 0 -> NO match
    assert oracle.match_views([(10, 5, 1.0)], [(12, 7, 3.0)], False, tb, ts, 0.7) == [[]]
    # mixed matrix: b0 matches s0 (D=2), b1 matches nothing (D=40, 58 > 9) ... s1 far from both
    r = oracle.match_views([(10, 5, 1.0), (50, 6, 2.0)], [(12, 7, 3.0), (108, 9, 4.0)], False, tb, ts, 0.7)
    assert r[1] == [] and len(r[0]) == 1 and r[0][0][0] == 7
    assert r[0][0][1] == 3.0 * (2 * (-(1.0 / 9)) + 1.0)
    # D == ovlp -> weight 0 -> score 0 but still a match
    r = oracle.match_views([(10, 5, 1.0), (90, 5, 1.0)], [(19, 7, 3.0)], False, tb, ts, 0.7)
    assert r[0][0][0] == 7 and abs(r[0][0][1]) < 1e-15 and r[1] == []


def test_match_views_velocity_constraint(oracle):
    """Two bottom candidates share one side candidate (colsum 2) -> velocity check decides.
    Bottom window: half template 15x15 at (x-7, y-7); 'moving' iff #(cur-prev > 25) >= 900*0.02 = 18
    (bottom) / 900*0.05 = 45 (side)."""
    tb = ts = (30, 30)
    I = np.zeros((120, 200), np.uint8)
    Ip = np.zeros_like(I)
    cb = [(40, 30, 2.0), (44, 60, 1.5), (150, 30, 1.0)]
    cs = [(42, 20, 3.0)]
    x0, y0b, y0s = 10, 50, 5
    # make bottom candidate 0 moving (full 15x15 window brightened), candidate 1 static, side moving
    I[y0b + 30 - 7: y0b + 30 + 8, x0 + 40 - 7: x0 + 40 + 8] = 200
    I[y0s + 20 - 7: y0s + 20 + 8, x0 + 42 - 7: x0 + 42 + 8] = 200
    r = oracle.match_views(cb, cs, True, tb, ts, 0.7, I, Ip, x0, y0b, y0s)
    assert [len(x) for x in r] == [1, 0, 0]
    # without the velocity check (video frame 0) both overlapping bottoms keep the side candidate
    r0 = oracle.match_views(cb, cs, False, tb, ts, 0.7, I, Ip, x0, y0b, y0s)
    assert [len(x) for x in r0] == [1, 1, 0]
    # threshold edge: exactly 17 brightened pixels in the bottom window -> not moving (needs >= 18)
    I2 = np.zeros_like(I)
    I2[y0s + 20 - 7: y0s + 20 + 8, x0 + 42 - 7: x0 + 42 + 8] = 200
    I2[y0b + 30 - 7, x0 + 40 - 7: x0 + 40 + 8] = 200   # 15 px
    I2[y0b + 30 - 6, x0 + 40 - 7: x0 + 40 - 5] = 200   # +2 = 17
    r17 = oracle.match_views(cb, cs, True, tb, ts, 0.7, I2, Ip, x0, y0b, y0s)
    assert [len(x) for x in r17] == [0, 0, 0]          # both bottoms static, side moving
    I2[y0b + 30 - 6, x0 + 40 - 5] = 200                # 18th pixel
    r18 = oracle.match_views(cb, cs, True, tb, ts, 0.7, I2, Ip, x0, y0b, y0s)
    assert [len(x) for x in r18] == [1, 0, 0]


@pytest.mark.parametrize("seed", range(10))
def test_match_views_random_vs_python_restatement(oracle, seed):
    rng = _rng(100 + seed)
    tb = (int(rng.integers(10, 31)), int(rng.integers(10, 31)))
    ts = (int(rng.integers(10, 31)), int(rng.integers(10, 31)))
    I = (rng.random((90, 160)) * 255).astype(np.uint8)
    Ip = np.where(rng.random(I.shape) < 0.5, I, (rng.random(I.shape) * 255).astype(np.uint8)).astype(np.uint8)
    nb, ns = int(rng.integers(0, 9)), int(rng.integers(0, 9))
    cb = [(int(rng.integers(0, 120)), int(rng.integers(0, 40)), float(np.float32(rng.uniform(0.1, 5)))) for _ in range(nb)]
    cs = [(int(rng.integers(0, 120)), int(rng.integers(0, 30)), float(np.float32(rng.uniform(0.1, 5)))) for _ in range(ns)]
    for vel in (False, True):
        got = oracle.match_views(cb, cs, vel, tb, ts, 0.7, I, Ip, 5, 45, 2)
        ref = py_match_views(cb, cs, vel, tb, ts, 0.7, I, Ip, 5, 45, 2)
        assert got == ref


# ---- geometry ------------------------------------------------------------------------------------------
def test_geometry_pads_and_q11(oracle):
    """30x30 templates: spre = ceil(29/2) = 15, spost_side = 14, spost_bottom = spre (move-assign quirk
    LocoMouse_class.cpp:3172-3173); canvas pads = max(box, pads) (672-682)."""
    from locomouse_cpp_b200 import synth

    spec = synth.SynthSpec()
    pads, canvas = oracle.geometry(spec.config(), synth.make_model(spec))
    assert list(pads) == [15, 15, 15, 15, 15, 15, 14, 14]
    assert list(canvas) == [400, 150, 15, 15]


def test_roi_check(oracle):
    from locomouse_cpp_b200 import synth

    spec = synth.SynthSpec()
    cfg, model = spec.config(), synth.make_model(spec)
    assert oracle.check_roi(cfg, model, 399, 164, 399) == 0
    assert oracle.check_roi(cfg, model, 14, 164, 399) == 0      # x = 14+400-15-400+1 = 0
    assert oracle.check_roi(cfg, model, 13, 164, 399) == -3     # leaves the canvas on the left
    assert oracle.check_roi(cfg, model, 1699, 164, 399) == 0
    assert oracle.check_roi(cfg, model, 1700, 164, 399) == -3
    assert oracle.check_roi(cfg, model, 800, 164, 400) == -3    # bottom edge: y+H > rows
