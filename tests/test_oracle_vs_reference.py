"""Pins the oracle's nmsMax / peakClustering restatement (and the host C++ Candidate / P22D mirror's contract)
against the REFERENCE'S OWN CODE: LocoMouse_class.cpp:1610-1905 and Candidates/Candidates.cpp compiled from
/root/reference by `make -C oracle ref` (value-type shim only, oracle/ref_shim).  The committed golden file
tests/golden/reference_nms.npz was produced by that library (tests/golden/make_reference_golden.py), so the pin
also holds where the reference is not mounted; where the library exists, fresh random maps are compared too."""
import os

import numpy as np
import pytest

from oracle import reference_nms as ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_nms.npz")


def _same(a, b):
    """a: the oracle's list of (x, y, score); b: structured array from the reference.  Scores bit for bit."""
    a = np.array(a, dtype=ref.CAND) if len(a) else np.zeros(0, ref.CAND)
    b = np.asarray(b)
    return len(a) == len(b) and np.array_equal(a["x"], b["x"]) and np.array_equal(a["y"], b["y"]) and np.array_equal(
        np.ascontiguousarray(a["s"]).view(np.uint64), np.ascontiguousarray(b["s"]).view(np.uint64))


def _cases():
    z = np.load(GOLD)
    names = sorted({k.rsplit("_", 1)[0] for k in z.files})
    return z, names


def test_oracle_matches_reference_golden(oracle):
    z, names = _cases()
    assert len(names) >= 10
    total = 0
    for n in names:
        s, (bw, bh) = z[f"{n}_scores"], z[f"{n}_box"]
        got_n = oracle.nms_max(s, int(bw), int(bh))
        got_p = oracle.peak_clustering(s, int(bw), int(bh))
        assert _same(got_n, z[f"{n}_nmsmax"]), f"nmsMax differs from the reference on {n}"
        assert _same(got_p, z[f"{n}_peak"]), f"peakClustering differs from the reference on {n}"
        total += len(got_n) + len(got_p)
    assert total > 500
    # the chain case separates the two suppression rules
    assert len(z["chain_nmsmax"]) == 1 and len(z["chain_peak"]) == 2
    # the half-way case separates the two rounding rules: x = 4.5 -> 4 (half-even) vs 5 (half-away)
    assert z["halfway_nmsmax"]["x"].tolist() != z["halfway_peak"]["x"].tolist() or z["halfway_nmsmax"]["y"].tolist() != z[
        "halfway_peak"]["y"].tolist()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_nms.so not built (reference not mounted)")
def test_golden_file_is_what_the_reference_code_produces():
    z, names = _cases()
    for n in names:
        s, (bw, bh) = z[f"{n}_scores"], z[f"{n}_box"]
        assert _same(ref.nms_max(s, int(bw), int(bh)), z[f"{n}_nmsmax"])
        assert _same(ref.peak_clustering(s, int(bw), int(bh)), z[f"{n}_peak"])


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_nms.so not built (reference not mounted)")
def test_oracle_matches_reference_on_fresh_maps(oracle):
    rng = np.random.Generator(np.random.PCG64(20261018))
    for it in range(40):
        rows, cols = int(rng.integers(8, 90)), int(rng.integers(8, 130))
        bw, bh = int(rng.integers(2, 31)), int(rng.integers(2, 31))
        s = rng.normal(0, 1, (rows, cols)).astype(np.float32)
        yy, xx = np.mgrid[0:rows, 0:cols]
        for _ in range(int(rng.integers(0, 5))):
            cy, cx, sg = rng.uniform(0, rows), rng.uniform(0, cols), rng.uniform(1.5, 7)
            s += (rng.uniform(1, 4) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * sg * sg))).astype(np.float32)
        s = (s - np.quantile(s, rng.uniform(0.5, 0.99))).astype(np.float32)
        s += np.arange(s.size, dtype=np.float32).reshape(s.shape) * np.float32(1e-7)
        if len(np.unique(s[s > 0])) != int((s > 0).sum()):
            continue  # std::sort is unstable on ties; the oracle's total order is only defined without them
        assert _same(oracle.nms_max(s, bw, bh), ref.nms_max(s, bw, bh)), f"nmsMax, map {it}"
        assert _same(oracle.peak_clustering(s, bw, bh), ref.peak_clustering(s, bw, bh)), f"peakClustering, map {it}"


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_nms.so not built (reference not mounted)")
def test_reference_p22d_contract():
    """The record semantics the C ABI's match arrays encode (include/locomouse_b200.h) and the host mirror
    (locomouse_cpp_b200/host/Candidates.cpp, test_candidates.cpp) reproduce, observed on the reference's own class."""
    assert ref.p22d(10, 20, 0.9, []) == (0, [])                               # sentinel: no side match
    assert ref.p22d(10, 20, 0.9, [(33, 0.25)]) == (1, [(33, 0.25)])
    assert ref.p22d(10, 20, 0.9, [(33, 0.25), (44, 0.0)]) == (2, [(33, 0.25), (44, 0.0)])
    n, _ = ref.p22d(1, 1, 1.0, [(9, -0.1)])                                    # negative first score keeps it "empty"
    assert n == 0
    x, y, s = (np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.float64))
    code = ref.lib().ref_default_candidate(x.ctypes.data, y.ctypes.data, s.ctypes.data)
    assert (int(x[0]), int(y[0]), float(s[0])) == (-1, -1, -1.0) and code == -1


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_nms.so not built (reference not mounted)")
def test_oracle_helpers_match_reference_code(oracle):
    """vecmovingaverage, firstLastOverT and LocoMouse::imadjust: the oracle vs the reference's own compiled lines."""
    rng = np.random.Generator(np.random.PCG64(99))
    for _ in range(60):
        n, w = int(rng.integers(1, 80)), int(rng.choice([1, 3, 5, 7, 9, 11]))
        v = rng.uniform(0, 1800, n)                      # non-negative: (uint32_t) of a negative double is UB in the reference
        assert np.array_equal(oracle.vecmovingaverage(v, w), ref.vecmovingaverage(v, w))
    for _ in range(60):
        vals = rng.integers(0, 60, int(rng.integers(1, 300))).astype(np.float32)
        th = int(rng.integers(0, 65))
        assert oracle.first_last_over_t(vals, th) == ref.first_last_over_t(vals, th)
    assert oracle.first_last_over_t([0, 7, 0], 7) == ref.first_last_over_t([0, 7, 0], 7) == (1, 0)
    assert oracle.first_last_over_t([0, 0], 1) == ref.first_last_over_t([0, 0], 1) == (-1, -1)
    for lo, hi, lo_o, hi_o in [(0.0, 0.6, 0.0, 1.0), (0.1, 0.9, 0.0, 1.0), (0.0, 1.0, 0.2, 0.8), (0.25, 0.5, 0.1, 1.0)]:
        assert np.array_equal(oracle.imadjust_lut(lo, hi, lo_o, hi_o), ref.imadjust_lut(lo, hi, lo_o, hi_o))
    # LocoMouse_TM::readFrame's imadjust(I, I, 0, 0.6, 0, 1) (LocoMouse_TM.cpp:247): 0, 2, 3, 5, 7, 8, 10, 12, ...
    assert ref.imadjust_lut()[:8].tolist() == [0, 2, 3, 5, 7, 8, 10, 12] and ref.imadjust_lut()[153:].min() == 255


# ---- pairing stage: matchingWithVelocityConstraint / xDist / matchViews / checkVelCriterion (class.cpp:1023-1267) ---------
PAIR_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_pairing.npz")


def _pair_cases():
    z = np.load(PAIR_GOLD)
    for i in range(len([k for k in z.files if k.endswith("_par")])):
        k = f"c{i:02d}"
        par = z[f"{k}_par"].tolist()
        lst = lambda a: [(int(x), int(y), float(s)) for x, y, s in a]
        want, o = [], 0
        for n in z[f"{k}_n"].tolist():
            want.append(list(zip(z[f"{k}_y"][o:o + n].tolist(), z[f"{k}_s"][o:o + n].tolist())))
            o += n
        yield dict(tsz=tuple(par[0:4]), org=tuple(par[4:7]), vel=par[7], T=float(z[f"{k}_T"]), I=z[f"{k}_img"][0], Ip=z[f"{k}_img"][1],
                   cb=lst(z[f"{k}_cb"]), cs=lst(z[f"{k}_cs"]), want=want)


def _same_pairs(a, b):
    if len(a) != len(b):
        return False
    for p, q in zip(a, b):
        if [y for y, _ in p] != [y for y, _ in q]:
            return False
        if not np.array_equal(np.array([s for _, s in p], np.float64).view(np.uint64), np.array([s for _, s in q], np.float64).view(np.uint64)):
            return False
    return True


def test_oracle_pairing_matches_reference_golden(oracle):
    """The oracle's pairing stage equals the reference's own compiled matchViews code, bit for bit (y, score x weight),
    on the committed vectors; the vectors exercise the velocity criterion both ways and the all-true boolD quirk (Q7)."""
    oracle.coverage(reset=True)
    n = 0
    for c in _pair_cases():
        twb, thb, tws, ths = c["tsz"]
        x0, y0b, y0s = c["org"]
        got = oracle.match_views(c["cb"], c["cs"], c["vel"], (twb, thb), (tws, ths), c["T"], c["I"], c["Ip"], x0, y0b, y0s)
        assert _same_pairs(got, c["want"]), f"pairing differs from the reference on case {n}"
        n += 1
    cov = oracle.coverage()
    assert n >= 60 and cov["all_equal_boolD_zeroed"] >= 2 and cov["velocity_rejections"] >= 3 and cov["velocity_accepts"] >= 10
    assert cov["moving_windows"] >= 5 and cov["side_matches"] >= 100 and cov["bottom_without_match"] >= 20


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_nms.so not built (reference not mounted)")
def test_pairing_golden_is_what_the_reference_code_produces():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_reference_pairing_golden", os.path.join(os.path.dirname(PAIR_GOLD), "make_reference_pairing_golden.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    for c, want in zip(g.cases(), _pair_cases()):
        assert _same_pairs(g.run_reference(c), want["want"])


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_nms.so not built (reference not mounted)")
def test_oracle_pairing_matches_reference_on_fresh_inputs(oracle):
    rng = np.random.Generator(np.random.PCG64(31337))
    nr, nc, W, hb, hs = 120, 200, 80, 50, 40
    for it in range(200):
        twb, thb, tws, ths = (int(v) for v in rng.integers(4, 31, 4))
        x0, y0b, y0s = int(rng.integers(-5, nc - W + 5)), int(rng.integers(30, nr - hb + 5)), int(rng.integers(-3, 30))
        I = rng.integers(0, 256, (nr, nc)).astype(np.uint8)
        Ip = I.copy()
        for _ in range(int(rng.integers(0, 12))):
            y, x = int(rng.integers(0, nr - 20)), int(rng.integers(0, nc - 20))
            Ip[y:y + 20, x:x + 20] = rng.integers(0, 60, (20, 20))
            I[y:y + 20, x:x + 20] = rng.integers(70, 256, (20, 20))
        nb, ns = int(rng.integers(0, 9)), int(rng.integers(0, 9))
        cx = rng.integers(0, W, 3)
        mk = lambda n_, h: [(int(np.clip(cx[rng.integers(0, 3)] + rng.integers(-12, 13), 0, W - 1)), int(rng.integers(0, h)),
                             float(rng.uniform(0.01, 3))) for _ in range(n_)]
        cb, cs = mk(nb, hb), mk(ns, hs)
        T = float(rng.choice([0.7, 0.5, 0.9]))
        if int(twb * (1 - T)) == 0:
            T = 0.5
        vel = int(rng.integers(0, 4) > 0)
        a = oracle.match_views(cb, cs, vel, (twb, thb), (tws, ths), T, I, Ip, x0, y0b, y0s)
        b = ref.match_views(cb, cs, vel, (twb, thb), (tws, ths), T, I, Ip, x0, y0b, y0s, hb, hs, W, (15, 15), (15, 15))
        assert _same_pairs(a, b), f"pairing, input {it}"


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_nms.so not built (reference not mounted)")
def test_reference_pairing_throws_when_overlap_is_zero():
    """ovlp == 0 (T close to 1) divides by zero; the weights become NaN and Candidates.cpp's CV_Assert(S >= 0)
    fires for a second side candidate.  The path is undefined there; the ABI rejects such a configuration."""
    I = np.zeros((40, 60), np.uint8)
    with pytest.raises(RuntimeError):
        ref.match_views([(10, 5, 1.0), (30, 5, 1.0)], [(10, 4, 1.0), (10, 9, 1.0)], 0, (4, 4), (4, 4), 0.9, I, I, 0, 20, 0, 20, 20, 60,
                        (15, 15), (15, 15))


# ---- tail stage: detectTail / detectLineCandidates / selectLargestRegion (class.cpp:2541-2767) -----------------------------
TAIL_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_tail.npz")


def _tail_gen():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_reference_tail_golden", os.path.join(os.path.dirname(TAIL_GOLD), "make_reference_tail_golden.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    return g


def test_oracle_tail_matches_reference_golden(oracle):
    """The oracle's tail segmentation (largest region, column extent, 15 segments, integer centroids, side z) equals the
    reference's own compiled control flow running on the real OpenCV's connected components, on the committed vectors."""
    z = np.load(TAIL_GOLD)
    n = len([k for k in z.files if k.endswith("_conn")])
    assert n >= 40
    found = 0
    for i in range(n):
        sb, ss = z[f"c{i:02d}_sb"].astype(np.float32) / 4, z[f"c{i:02d}_ss"].astype(np.float32) / 4
        t, m = oracle.tail_from_binary((sb > 0).astype(np.uint8), (ss > 0).astype(np.uint8), int(z[f"c{i:02d}_conn"]), 15)
        assert np.array_equal(t, z[f"c{i:02d}_tracks"]), f"tail tracks differ from the reference on case {i}"
        assert np.array_equal(np.packbits(m > 0), z[f"c{i:02d}_mask"]), f"TAIL_MASK differs from the reference on case {i}"
        found += int((t[0] >= 0).sum())
    assert found > 200
    assert (z["c00_tracks"] == -1).all() and not z["c00_mask"].any()          # no foreground
    assert z["c01_tracks"][0, 0] == 0 and z["c01_tracks"][1, 0] == 4 and (z["c01_tracks"][2] == -1).all()   # x == 0: no side z (Q13)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_nms.so not built (reference not mounted)")
def test_tail_golden_is_what_the_reference_code_produces_and_fresh_maps_agree(oracle):
    pytest.importorskip("cv2")
    g = _tail_gen()
    z = np.load(TAIL_GOLD)
    for i, c in enumerate(g.cases()):
        t, m = ref.detect_tail(*g.maps(c), c["conn"], 15)
        assert np.array_equal(t, z[f"c{i:02d}_tracks"]) and np.array_equal(np.packbits(m > 0), z[f"c{i:02d}_mask"])
    for c in g.cases(seed=5, n=80):
        sb, ss = g.maps(c)
        t, m = ref.detect_tail(sb, ss, c["conn"], 15)
        to, mo = oracle.tail_from_binary((sb > 0).astype(np.uint8), (ss > 0).astype(np.uint8), c["conn"], 15)
        assert np.array_equal(t, to) and np.array_equal(m > 0, mo > 0)


# ---- frame-level glue: detectBottomCandidates / detectSideCandidates (class.cpp:771-870) ------------------------------------
FRAME_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_frame_candidates.npz")


def _lists_equal(got, want):
    want = np.asarray(want)
    a = np.array(got, dtype=ref.CAND) if len(got) else np.zeros(0, ref.CAND)
    return len(a) == len(want) and np.array_equal(a["x"], want["x"]) and np.array_equal(a["y"], want["y"]) and np.array_equal(
        np.ascontiguousarray(a["s"]).view(np.uint64), np.ascontiguousarray(want["s"]).view(np.uint64))


def test_oracle_frame_candidates_match_reference_golden(oracle):
    """oracle.detect's four candidate lists per frame (masks at <= 25, TAIL_MASK over the tail box, unpadded window, nmsMax /
    peakClustering, side detection skipped when the bottom list is empty) equal what the reference's own
    detectBottomCandidates / detectSideCandidates code produced for the same frames (committed vectors)."""
    import _frame_glue as G

    z = np.load(FRAME_GOLD)
    total = 0
    for ci, kw in enumerate(G.CASES):
        cfg, model, bkg, calib, frames, bx, bs, bb = G.case_problem(kw)
        res = oracle.detect(cfg, model, bkg, calib, frames, bx, bs, bb, n_threads=4)
        for f in range(len(frames)):
            got = [res.candidates_bottom(f, 0), res.candidates_bottom(f, 1), res.candidates_side(f, 0), res.candidates_side(f, 1)]
            for k in range(4):
                assert _lists_equal(got[k], z[f"c{ci}_f{f}_l{k}"]), f"case {kw}, frame {f}, list {k}"
                total += len(got[k])
            if kw.get("q6"):
                assert got[0] == [] and got[2] == [] and len(got[3]) > 0   # Q6: side paw skipped, side snout still runs
    assert total > 300


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_nms.so not built (reference not mounted)")
def test_frame_golden_is_what_the_reference_code_produces(oracle):
    pytest.importorskip("cv2")
    import _frame_glue as G

    z = np.load(FRAME_GOLD)
    for ci, kw in enumerate(G.CASES[:2] + G.CASES[3:]):
        ci = G.CASES.index(kw)
        cfg, model, bkg, calib, frames, bx, bs, bb = G.case_problem(kw, n=2)
        for f in range(2):
            lists = G.reference_frame(oracle, cfg, model, bkg, calib, frames[f], bx[f], bs[f], bb[f])
            for k in range(4):
                assert _lists_equal(lists[k], z[f"c{ci}_f{f}_l{k}"])


# ---- readFrame / correctImage (class.cpp:1273-1406) + LocoMouse_TM::readFrame's imadjust -------------------------------------
READ_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_readframe.npz")


def _read_gen():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_reference_readframe_golden", os.path.join(os.path.dirname(READ_GOLD), "make_reference_readframe_golden.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    return g


def test_oracle_preprocess_matches_reference_readframe_golden(oracle):
    """oracle.preprocess (background subtraction, per-frame min-max normalisation, calibration gather, mirror, imadjust) gives
    the image the reference's own readFrame / correctImage code gave on the real OpenCV's normalize / flip (checksums)."""
    import zlib

    g = _read_gen()
    z = np.load(READ_GOLD)
    for ci, kw in enumerate(g.CASES):
        cfg, bkg, calib, frames = g.frames_of(kw)
        for f, fr in enumerate(frames):
            img, _ = oracle.preprocess(cfg, bkg, calib, fr)
            crc, total, mx = z[f"c{ci}_f{f}"].tolist()
            assert (zlib.crc32(img.tobytes()), int(img.sum()), int(img.max())) == (crc, total, mx), f"case {kw}, frame {f}"
        assert z[f"c{ci}_f0"][1] > 100000 and z[f"c{ci}_f3"][1] == 0   # a real image; the constant frame maps to zeros


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_nms.so not built (reference not mounted)")
def test_readframe_golden_is_what_the_reference_code_produces(oracle):
    pytest.importorskip("cv2")
    import zlib

    g = _read_gen()
    z = np.load(READ_GOLD)
    for ci, kw in enumerate(g.CASES):
        cfg, bkg, calib, frames = g.frames_of(kw)
        for f in (0, 3):
            img = g.reference_image(cfg, bkg, calib, frames[f])
            assert zlib.crc32(img.tobytes()) == int(z[f"c{ci}_f{f}"][0])
            want, _ = oracle.preprocess(cfg, bkg, calib, frames[f])
            assert np.array_equal(img, want)


# ---- pass 1, LocoMouse_TM_DE: computeMouseBox_DE + imadjust_default (TM_DE.cpp:56-113, class.cpp:3244-3311) ------------------
PASS1_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_pass1.npz")


def _pass1_gen():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_reference_pass1_golden", os.path.join(os.path.dirname(PASS1_GOLD), "make_reference_pass1_golden.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    return g


def test_oracle_pass1_matches_reference_golden(oracle):
    """The oracle's pass 1 (per-frame bb_x of LocoMouse_TM_DE) equals what the reference's own readFrame -> imadjust_default ->
    computeMouseBox_DE code produced, as exact doubles, incl. degenerate frames (-1.1 = no qualifying column x WIDTH_MARGIN)."""
    from locomouse_cpp_b200.types import bb_de_params

    g = _pass1_gen()
    z = np.load(PASS1_GOLD)
    for ci, kw in enumerate(g.CASES):
        cfg, bkg, calib, frames = g.frames_of(kw)
        _, raw, _ = oracle.bounding_box_tm_de(cfg, bkg, calib, frames, bb_de_params(cfg, side_h=g.SIDE_H), window=5)
        assert np.array_equal(raw.view(np.uint64), z[f"c{ci}"].view(np.uint64)), f"case {kw}: {raw.tolist()} vs {z[f'c{ci}'].tolist()}"
    assert len(set(z["c0"].tolist())) >= 3 and z["c0"][-1] == -1.1


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_nms.so not built (reference not mounted)")
def test_pass1_golden_is_what_the_reference_code_produces():
    pytest.importorskip("cv2")
    g = _pass1_gen()
    z = np.load(PASS1_GOLD)
    cfg, bkg, calib, frames = g.frames_of(g.CASES[1])
    got = [ref.mouse_box_de(ref.read_frame(fr, bkg, calib, cfg.flip)[:g.SIDE_H]) for fr in frames[[0, 3, 8]]]
    assert got == z["c1"][[0, 3, 8]].tolist()
