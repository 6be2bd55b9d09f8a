"""Pins the CPU oracle against OpenCV itself (python cv2), primitive by primitive.

The reference cannot be built here (no OpenCV C++ SDK) and ships no tests, so the executable
anchor for the oracle is the library the reference calls: every OpenCV call on the hot path
 is executed through cv2 with the reference's arguments and compared with
the oracle's restatement.  CPU only.
"""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


# ---- readFrame: subtract + normalize + gather + flip (+ imadjust LUT) --------------------------
@pytest.mark.parametrize("flip,imadjust,seed", [(False, False, 1), (True, False, 2), (False, True, 3), (True, True, 4)])
def test_preprocess_matches_cv2(oracle, flip, imadjust, seed):
    from locomouse_cpp_b200.types import Config

    rng = _rng(seed)
    vr, vc, nr, nc = 97, 211, 90, 200
    cfg = Config(vid_rows=vr, vid_cols=vc, n_rows=nr, n_cols=nc, bb_w=50, bb_h_bottom=40, bb_h_side=30, flip=flip,
                 imadjust=imadjust)
    bkg = rng.integers(10, 70, (vr, vc), dtype=np.uint8)
    frame = np.clip(bkg.astype(int) + rng.integers(-12, 190, (vr, vc)), 0, 255).astype(np.uint8)
    calib = rng.integers(0, vr * vc, (nr, nc)).astype(np.int32)
    got, mm = oracle.preprocess(cfg, bkg, calib, frame)

    F = cv2.subtract(frame, bkg)                                       # class.cpp:1304
    assert (int(F.min()), int(F.max())) == (int(mm[0]), int(mm[1]))
    F = cv2.normalize(F, None, 0, 255, cv2.NORM_MINMAX, cv2.CV_8UC1)   # class.cpp:1310
    I = F.reshape(-1)[calib]                                           # class.cpp:1396
    if flip:
        I = cv2.flip(I, 1)                                             # class.cpp:1324
    if imadjust:                                                       # LocoMouse_TM.cpp:247
        lut = np.array([0 if i <= 0 else 255 if i >= 0.6 * 255 else int(np.floor(i * (255.0 / (0.6 * 255)) + 0.5))
                        for i in range(256)], np.uint8)
        assert np.array_equal(lut, oracle.imadjust_lut())
        I = cv2.LUT(np.ascontiguousarray(I), lut)
    assert np.array_equal(got, I)


def test_normalize_constant_frame(oracle):
    """smax == smin -> scale 0 -> all zeros (cv::normalize)."""
    from locomouse_cpp_b200.types import Config

    cfg = Config(vid_rows=8, vid_cols=8, n_rows=8, n_cols=8, bb_w=4, bb_h_bottom=4, bb_h_side=4, imadjust=False)
    bkg = np.full((8, 8), 10, np.uint8)
    frame = np.full((8, 8), 40, np.uint8)
    calib = np.arange(64, dtype=np.int32).reshape(8, 8)
    got, mm = oracle.preprocess(cfg, bkg, calib, frame)
    ref = cv2.normalize(cv2.subtract(frame, bkg), None, 0, 255, cv2.NORM_MINMAX, cv2.CV_8UC1)
    assert np.array_equal(got, ref) and got.max() == 0


def test_imadjust_lut_head(oracle):
    """LUT starts 0,2,3,5,7,8,10,12 and saturates at 153."""
    lut = oracle.imadjust_lut()
    assert list(lut[:8]) == [0, 2, 3, 5, 7, 8, 10, 12]
    assert lut[152] == 253 and lut[153] == 255 and lut[255] == 255


# ---- filter2D -----------------------------------------------------------------------------------
def _canvas(I, pad):
    C = np.zeros((I.shape[0] + 2 * pad, I.shape[1] + 2 * pad), np.uint8)
    C[pad:-pad, pad:-pad] = I
    return C


@pytest.mark.parametrize("kh,kw", [(7, 7), (5, 9), (6, 6), (4, 7), (3, 16), (1, 1)])
def test_correlate_bitexact_vs_filter2d_direct_path(oracle, kh, kw):
    """< 50 taps: cv2 uses its direct filter engine; the oracle's mul+add mode must match bit for bit,
    including the (cols/2, rows/2) anchor of even kernels, delta = -rho and zero extension."""
    rng = _rng(kh * 100 + kw)
    I = rng.integers(0, 256, (60, 90), dtype=np.uint8)
    k = rng.normal(0, 0.02, (kh, kw)).astype(np.float32)
    rho = 0.37
    pad = 20
    # filter the WHOLE canvas (a numpy slice is an isolated image for cv2, not an ROI), then crop
    full = cv2.filter2D(_canvas(I, pad), cv2.CV_32F, k, anchor=(-1, -1), delta=-rho, borderType=cv2.BORDER_CONSTANT)
    for (x0, y0, w, h) in [(0, 0, 90, 60), (-7, -5, 40, 30), (60, 40, 38, 27)]:
        ref = full[pad + y0: pad + y0 + h, pad + x0: pad + x0 + w]
        got = oracle.correlate(I, k, rho, x0, y0, w, h, fma_mode=False)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_correlate_30x30_within_tolerance_of_cv2(oracle):
    """>= 50 taps cv2 takes its DFT path: agreement to 1e-5 relative of the score scale for both oracle
    modes; fused and two-rounding modes agree to 1e-5 as well (north_star tolerance)."""
    rng = _rng(5)
    I = rng.integers(0, 256, (120, 160), dtype=np.uint8)
    k = rng.normal(0, 1.0 / (255 * 30), (30, 30)).astype(np.float32)
    rho = 0.1
    pad = 32
    full = cv2.filter2D(_canvas(I, pad), cv2.CV_32F, k, anchor=(-1, -1), delta=-rho, borderType=cv2.BORDER_CONSTANT)
    ref = full[pad:-pad, pad:-pad]
    a = oracle.correlate(I, k, rho, 0, 0, 160, 120, fma_mode=False)
    b = oracle.correlate(I, k, rho, 0, 0, 160, 120, fma_mode=True)
    scale = np.abs(ref).max()
    assert np.abs(a - ref).max() <= 1e-5 * scale
    assert np.abs(b - ref).max() <= 1e-5 * scale
    assert np.abs(a - b).max() <= 1e-5 * scale


def test_mask_threshold_semantics():
    """threshold(.., 25.5, 255, THRESH_BINARY_INV) on u8 masks px <= 25 (class.cpp:782,817)."""
    v = np.arange(256, dtype=np.uint8).reshape(1, -1)
    _, m = cv2.threshold(v, 25.5, 255, cv2.THRESH_BINARY_INV)
    assert np.array_equal(m[0] == 255, np.arange(256) <= 25)


# ---- connected components / largest region -------------------------------------------------------
def _cv_largest(binary, conn):
    n, labels, stats, _ = cv2.connectedComponentsWithStats(binary, connectivity=conn, ltype=cv2.CV_16U)
    if n <= 1:
        return np.zeros_like(binary)
    best, area = 1, stats[1, cv2.CC_STAT_AREA]
    for i in range(2, n):                      # strict '>' : class.cpp:2752-2757
        if stats[i, cv2.CC_STAT_AREA] > area:
            best, area = i, stats[i, cv2.CC_STAT_AREA]
    return ((labels == best) * 255).astype(np.uint8)


@pytest.mark.parametrize("conn", [4, 8])
def test_largest_region_matches_cv2_including_ties(oracle, conn):
    rng = _rng(conn)
    n_ties = 0
    for trial in range(300):
        h, w = int(rng.integers(3, 40)), int(rng.integers(3, 48))
        dens = rng.choice([0.08, 0.2, 0.45, 0.6])
        b = (rng.random((h, w)) < dens).astype(np.uint8)
        ref = _cv_largest(b, conn)
        got = oracle.largest_region(b, conn)
        assert np.array_equal(got, ref), f"trial {trial} conn {conn}"
        n, _, stats, _ = cv2.connectedComponentsWithStats(b, connectivity=conn)
        if n > 2:
            areas = sorted(stats[1:, cv2.CC_STAT_AREA])
            n_ties += areas[-1] == areas[-2]
    assert n_ties > 20  # the tie-break rule was really exercised


def test_largest_region_block_scan_order_8conn(oracle):
    """Two equal-area components: A's first pixel is (1,4) (odd row), B's is (0,8).  OpenCV's 8-conn
    labelling scans 2x2 blocks, so A (block column 2) gets the lower label although B comes first in
    pixel raster order."""
    b = np.zeros((6, 12), np.uint8)
    b[1, 4] = b[2, 4] = 1
    b[0, 8] = b[1, 8] = 1
    ref = _cv_largest(b, 8)
    assert np.array_equal(oracle.largest_region(b, 8), ref)
    assert ref[1, 4] == 255 and ref[0, 8] == 0
    ref4 = _cv_largest(b, 4)
    assert np.array_equal(oracle.largest_region(b, 4), ref4)


def test_largest_region_empty(oracle):
    assert oracle.largest_region(np.zeros((5, 7), np.uint8)).max() == 0


# ---- tail: moments / segments -------------------------------------------------------------------
def _tail_with_cv2(bin_b, bin_s, conn, n_points):
    """detectLineCandidates class.cpp:2593-2742 re-executed with cv2 primitives."""
    tracks = -np.ones((3, n_points), np.int32)
    mask = _cv_largest(bin_b, conn)
    colmax = cv2.reduce(mask, 0, cv2.REDUCE_MAX)
    side = ((bin_s > 0).astype(np.uint8) * 255) & np.repeat(colmax, bin_s.shape[0], 0)
    side = _cv_largest(side, conn)
    nz = np.flatnonzero(colmax[0] > 0)
    if nz.size == 0:
        return tracks, mask
    first, last = int(nz[0]), int(nz[-1])
    width = last - first
    rem = width % n_points
    reg = (width - rem) // n_points
    x = first
    for i in range(n_points):
        wseg = reg + (1 if i < rem else 0)
        seg = mask[:, x:x + wseg]
        if seg.size:
            M = cv2.moments(seg, True)
            if M["m00"] > 0:
                tracks[0, i] = int(M["m10"] / M["m00"]) + x
                tracks[1, i] = int(M["m01"] / M["m00"])
        x += wseg
    for i in range(n_points):
        if tracks[0, i] > 0:
            M = cv2.moments(side[:, tracks[0, i]:tracks[0, i] + 1], True)
            if M["m00"] > 0:
                tracks[2, i] = int(M["m01"] / M["m00"])
    return tracks, mask


@pytest.mark.parametrize("conn", [4, 8])
def test_tail_tracks_match_cv2(oracle, conn):
    rng = _rng(77 + conn)
    for trial in range(40):
        hb, hs, w = 40, 28, int(rng.integers(20, 70))
        bb = np.zeros((hb, w), np.uint8)
        bs = np.zeros((hs, w), np.uint8)
        x0, x1 = sorted(rng.integers(0, w, 2))
        for x in range(x0, x1 + 1):   # a wavy line of random thickness + clutter
            yc = int(20 + 8 * np.sin(x / 5.0 + trial))
            bb[max(0, yc - 1): yc + int(rng.integers(1, 4)), x] = 1
            zc = int(14 + 5 * np.cos(x / 6.0))
            bs[max(0, zc - 1): zc + int(rng.integers(1, 3)), x] = 1
        bb |= (rng.random(bb.shape) < 0.03).astype(np.uint8)
        bs |= (rng.random(bs.shape) < 0.03).astype(np.uint8)
        ref_t, ref_m = _tail_with_cv2(bb, bs, conn, 15)
        got_t, got_m = oracle.tail_from_binary(bb, bs, conn, 15)
        assert np.array_equal(got_m, ref_m)
        assert np.array_equal(got_t, ref_t), f"trial {trial}\n{got_t}\n{ref_t}"


def test_tail_short_and_x0_cases(oracle):
    """Fewer columns than points -> zero-width segments stay -1; x == 0 gets no side z (class.cpp:2729)."""
    bb = np.zeros((10, 20), np.uint8)
    bs = np.ones((8, 20), np.uint8)
    bb[4:6, 0:4] = 1  # columns 0..3 -> first=0,last=3,width=3 -> segments 1,1,1,0...
    t, m = oracle.tail_from_binary(bb, bs, 8, 15)
    assert list(t[0, :4]) == [0, 1, 2, -1] and list(t[1, :3]) == [4, 4, 4]
    assert t[2, 0] == -1 and t[2, 1] == 3 and t[2, 2] == 3   # x == 0 skipped
    t0, m0 = oracle.tail_from_binary(np.zeros((10, 20), np.uint8), bs, 8, 15)
    assert (t0 == -1).all() and m0.max() == 0
    ref_t, _ = _tail_with_cv2(bb, bs, 8, 15)
    assert np.array_equal(t, ref_t)


# ---- pairing primitives ----------------------------------------------------------------------------
def test_normalize_boolD_quirk_q7():
    """normalize(boolD, 0, 1, NORM_MINMAX): an all-255 (or all-0) matrix becomes all zeros."""
    for m, expect in (([[255]], [[0]]), ([[255, 255], [255, 255]], [[0, 0], [0, 0]]),
                      ([[255, 0]], [[1, 0]]), ([[0, 0]], [[0, 0]])):
        a = np.array(m, np.uint8)
        out = cv2.normalize(a, None, 0, 1, cv2.NORM_MINMAX, -1)
        assert out.tolist() == expect


def test_velocity_window_primitives():
    """subtract saturates, threshold(S, 25, 1, BINARY) is strict > (class.cpp:1259-1263)."""
    a = np.array([[100, 10, 60, 26]], np.uint8)
    b = np.array([[74, 90, 35, 0]], np.uint8)
    s = cv2.subtract(a, b)
    _, t = cv2.threshold(s, 25, 1, cv2.THRESH_BINARY)
    assert s.tolist() == [[26, 0, 25, 26]] and t.tolist() == [[1, 0, 0, 1]]


def test_reference_shim_primitives_match_cv2():
    """oracle/ref_shim/opencv2/core.hpp supplies the few elementwise OpenCV primitives that the reference's pairing code
    (compiled from /root/reference into oracle/_ref/libref_nms.so) calls.  Each is checked here against the real OpenCV."""
    import ctypes as C

    from oracle import reference_nms as ref
    if not ref.available():
        pytest.skip("oracle/_ref/libref_nms.so not built (reference not mounted)")
    L = ref.lib()
    L.ref_shim_primitives.restype = None
    rng = np.random.Generator(np.random.PCG64(2718))
    for it in range(120):
        rows, cols, ovlp = int(rng.integers(1, 9)), int(rng.integers(1, 9)), int(rng.integers(1, 20))
        D = rng.integers(0, 30, (rows, cols)).astype(np.int32)
        if it % 10 == 0:
            D[:] = 0           # all true: normalize maps the constant matrix to zeros
        if it % 10 == 1:
            D[:] = 100         # all false
        r2, c2 = int(rng.integers(1, 20)), int(rng.integers(1, 20))
        a, b = rng.integers(0, 256, (r2, c2)).astype(np.uint8), rng.integers(0, 256, (r2, c2)).astype(np.uint8)
        raw, norm = np.zeros((rows, cols), np.uint8), np.zeros((rows, cols), np.uint8)
        cs, rs, w, cnt = np.zeros(cols, np.float32), np.zeros(rows, np.float32), np.zeros((rows, cols), np.float64), np.zeros(1, np.float64)
        p = lambda x: C.c_void_p(x.ctypes.data)
        L.ref_shim_primitives(p(D), rows, cols, ovlp, p(raw), p(norm), p(cs), p(rs), p(w), p(a), p(b), r2, c2, C.c_double(25.0), p(cnt))
        want_raw = cv2.compare(D, np.full_like(D, ovlp), cv2.CMP_LE)
        assert np.array_equal(raw, want_raw.reshape(rows, cols))
        want_norm = cv2.normalize(want_raw, None, 0, 1, cv2.NORM_MINMAX, -1).reshape(rows, cols)
        assert np.array_equal(norm, want_norm)
        assert np.array_equal(cs, cv2.reduce(want_norm, 0, cv2.REDUCE_SUM, dtype=cv2.CV_32F).ravel())
        assert np.array_equal(rs, cv2.reduce(want_norm, 1, cv2.REDUCE_SUM, dtype=cv2.CV_32F).ravel())
        # 1 - D/ovlp is ONE convertTo(alpha = -(1/ovlp), beta = 1) in OpenCV (MatOp_AddEx); cv2 exposes no f64 convertTo, so the
        # arithmetic form is stated here: one multiply and one add in double
        assert np.array_equal(w, D.astype(np.float64) * (-(1.0 / ovlp)) + 1.0)
        s = cv2.subtract(a, b)
        _, t = cv2.threshold(s, 25.0, 1, cv2.THRESH_BINARY)
        assert cnt[0] == cv2.sumElems(t)[0]


def test_reference_shim_primitives_batch2_match_cv2():
    """The elementwise shim primitives behind the compiled tail / candidate / readFrame code, each against the real OpenCV."""
    import ctypes as C

    from oracle import reference_nms as ref
    if not ref.available():
        pytest.skip("oracle/_ref/libref_nms.so not built (reference not mounted)")
    L = ref.lib()
    L.ref_shim_primitives2.restype = None
    rng = np.random.Generator(np.random.PCG64(31415))
    for it in range(60):
        rows, cols = int(rng.integers(1, 24)), int(rng.integers(1, 24))
        a = rng.integers(0, 256, (rows, cols)).astype(np.uint8)
        a[rng.random((rows, cols)) < 0.3] = rng.choice([0, 25, 26])
        b = (rng.random((rows, cols)) < 0.4).astype(np.uint8) * int(rng.choice([1, 255]))
        f = rng.normal(0, 1, (rows, cols)).astype(np.float32)
        f[rng.random((rows, cols)) < 0.2] = 0.0
        lab = rng.integers(0, 5, (rows, cols)).astype(np.uint16)
        val = int(rng.integers(0, 5))
        out8 = np.zeros((7, rows, cols), np.uint8)
        outf = np.zeros((rows, cols), np.float32)
        mom = np.zeros(3, np.float64)
        p = lambda x: C.c_void_p(x.ctypes.data)
        L.ref_shim_primitives2(p(a), p(b), p(f), p(lab), val, rows, cols, p(out8), p(outf), p(mom))
        _, t = cv2.threshold(f, 0, 1, cv2.THRESH_BINARY)
        assert np.array_equal(out8[0], t.astype(np.uint8))
        assert np.array_equal(out8[1], cv2.threshold(a, 25.5, 255, cv2.THRESH_BINARY_INV)[1])
        assert np.array_equal(out8[2], cv2.repeat(cv2.reduce(a, 0, cv2.REDUCE_MAX), rows, 1))
        assert np.array_equal(out8[3], cv2.bitwise_and(cv2.compare(f, np.zeros_like(f), cv2.CMP_GT), a))
        assert np.array_equal(out8[4], cv2.compare(lab, np.full_like(lab, val), cv2.CMP_EQ))
        assert np.array_equal(out8[5], cv2.subtract(a, b))
        want = a.copy()
        want[b != 0] = 7                      # Mat::setTo(value, mask)
        assert np.array_equal(out8[6], want)
        wf = f.copy()
        wf[b != 0] = 0
        assert np.array_equal(outf, wf)
        m = cv2.moments(a, True)
        assert (mom[0], mom[1], mom[2]) == (m["m00"], m["m10"], m["m01"])


def test_reference_shim_primitives_batch3_match_cv2():
    """calcHist / sum / colRange / rowRange / setTo of the shim (pass-1 code) against the real OpenCV."""
    import ctypes as C

    from oracle import reference_nms as ref
    if not ref.available():
        pytest.skip("oracle/_ref/libref_nms.so not built (reference not mounted)")
    L = ref.lib()
    L.ref_shim_primitives3.restype = None
    rng = np.random.Generator(np.random.PCG64(1618))
    for _ in range(20):
        rows, cols = int(rng.integers(2, 40)), int(rng.integers(4, 60))
        a = rng.integers(0, 256, (rows, cols)).astype(np.uint8)
        hist, total, band = np.zeros(256, np.float32), np.zeros(1, np.float64), np.zeros((rows, cols), np.uint8)
        p = lambda x: C.c_void_p(x.ctypes.data)
        L.ref_shim_primitives3(p(a), rows, cols, p(hist), p(total), p(band))
        want = cv2.calcHist([a], [0], None, [256], [0, 256]).ravel()
        assert np.array_equal(hist, want) and total[0] == cv2.sumElems(want)[0]
        w = a.copy()
        w[:, :cols // 4] = 0
        w[rows // 2:, :] = 0
        assert np.array_equal(band, w)
