"""Python model of the driver stages after the per-frame loop (LocoMouse_class.cpp:2153-2482: computeBottomTracks,
bestSideViewMatch, exportPointTracks, exportLineTracks), used to check the C++ driver's output_<stem>.yml.  The tracker
and the side-view transition builder are the REFERENCE's own compiled code when oracle/_ref/ is present (it travels to
the GPU box), otherwise the host library under test."""
import ctypes as C
import os

import numpy as np

from test_match2nd import HOST_LIB, REF_LIB, Trellis, _host, _ref, run

PAW_ORDERS = [[3, 2, 1, 0], [2, 3, 1, 0], [1, 2, 3, 0], [0, 2, 1, 3]]  # rows 0..3 of the 4 x 24 PAW_PERMUTATIONS


def tracker():
    R = _ref()
    ref_nms = os.path.join(os.path.dirname(REF_LIB), "libref_nms.so")
    if R is not None and os.path.exists(ref_nms):
        return R, "ref_match2nd", C.CDLL(ref_nms), "ref_pairwise_potential_side", "reference"
    H = _host()
    return H, "lmh_match2nd", H, "lmh_pairwise_potential_side", "host"


def side_transitions(lib, name, zi, zip1, grid_mapping, spacing, nong, max_disp, alpha, occ):
    a, b = np.asarray(zi, np.uint32), np.asarray(zip1, np.uint32)
    jc = np.zeros(len(a) + nong + 1, np.int32)
    cap = (len(a) + nong) * (len(b) + nong) + 1
    ir, pr, dims = np.zeros(cap, np.int32), np.zeros(cap, np.float64), np.zeros(3, np.int32)
    fn = getattr(lib, name)
    fn.restype = C.c_int
    p = lambda x: C.c_void_p(x.ctypes.data)
    rc = fn(p(a), len(a), p(b), len(b), C.c_double(grid_mapping), C.c_double(spacing), int(nong), C.c_double(max_disp), C.c_double(alpha),
            C.c_double(occ), p(jc), p(ir), p(pr), cap, p(dims))
    assert rc == 0
    return int(dims[0]), int(dims[1]), jc, ir[:dims[2]].copy(), pr[:dims[2]].copy()


def cost_track(labels, unary, perm):
    """computeCostTrack (match2nd.cpp:162-191): unary terms only (MATSPARSE::get returns 0), in track / frame order."""
    c = 0.0
    for t in range(4):
        for f in range(labels.shape[1]):
            lab = labels[t, f]
            c += float(unary[f][lab, perm[t]]) if 0 <= lab < unary[f].shape[0] else 0.0
    return c


def bottom_tracks(unary_paw, trans_paw, unary_snout, trans_snout, nong):
    lib, name, _, _, _ = tracker()
    n_loc = [u.shape[0] for u in unary_paw]
    best, best_cost, best_perm = None, -1.0, 0
    for i, perm in enumerate(PAW_ORDERS):
        lab = run(lib, name, Trellis(4, nong, n_loc, unary_paw, trans_paw), 0.0, 0.0, perm)
        c = cost_track(lab, unary_paw, perm)
        if c > best_cost:
            best, best_cost, best_perm = lab, c, i
    paw = np.zeros_like(best)
    for r in range(4):
        paw[PAW_ORDERS[best_perm][r]] = best[r]
    snout = run(lib, name, Trellis(1, nong, [u.shape[0] for u in unary_snout], unary_snout, trans_snout), 0.0, 0.0, [0])
    return paw, snout


def side_tracks(T, p22d, side_h, spacing_side=20, max_disp_side=15, alpha_side=100.0, occ=1e-2):
    """p22d[f] = list of (bottom candidate, [(y_side, score)...]) for the feature; T = points x frames bottom labels."""
    lib, name, slib, sname, _ = tracker()
    nong = (side_h - spacing_side) // spacing_side + 1
    lowest = side_h - 1 - spacing_side // 2
    n_frames = T.shape[1]
    out = np.zeros_like(T)
    for k in range(T.shape[0]):
        unary, trans, zprev = [], [], []
        for f in range(n_frames):
            z, u = [], np.zeros((0, 1))
            if 0 <= T[k, f] < len(p22d[f]):
                side = p22d[f][T[k, f]][1]
                u = np.array([[s] for _y, s in side], np.float64).reshape(len(side), 1)
                z = [y for y, _s in side]
            unary.append(u)
            if f > 0:
                _r, _c, jc, ir, pr = side_transitions(slib, sname, zprev, z, float(lowest), float(spacing_side), nong, float(max_disp_side), alpha_side, occ)
                trans.append((jc, ir, pr))
            zprev = z
        out[k] = run(lib, name, Trellis(1, nong, [u.shape[0] for u in unary], unary, trans), 0.0, 0.0, [0])[0]
    return out


def export_points(T_bottom, T_side, p22d, bx, bs, bb, bb_w, bb_hb, bb_hs):
    mats = []
    for k in range(T_bottom.shape[0]):
        M = -np.ones((T_bottom.shape[1], 3), np.int32)
        for f in range(T_bottom.shape[1]):
            lab = T_bottom[k, f]
            if not (0 <= lab < len(p22d[f])):
                continue
            (x, y, _s), side = p22d[f][lab]
            M[f, 0] = int(bx[f]) - bb_w + 1 + x
            M[f, 1] = int(bb[f]) - bb_hb + 1 + y
            if 0 <= T_side[k, f] < len(side):
                M[f, 2] = int(bs[f]) - bb_hs + 1 + side[T_side[k, f]][0]
        mats.append(M)
    return mats


def export_tail(tail, bx, bs, bb, bb_w, bb_hb, bb_hs):
    n, _, npts = tail.shape
    L = -np.ones((3, npts * n), np.int32)
    for f in range(n):
        for k in range(npts):
            c = f * npts + k
            if tail[f, 0, k] >= 0:
                L[0, c] = int(bx[f]) - bb_w + 1 + tail[f, 0, k]
            if tail[f, 1, k] >= 0:
                L[1, c] = int(bb[f]) - bb_hb + 1 + tail[f, 1, k]
            if tail[f, 2, k] >= 0:
                L[2, c] = int(bs[f]) - bb_hs + 1 + tail[f, 2, k]
    return L
