"""Parity tests proper: the CUDA path, called through the C ABI (liblocomouse_b200.so via ctypes),
against the CPU oracle on identical inputs and against the committed golden fixtures.

Bar (BASELINE.json north_star): candidate coordinates, counts, pairings, tail tracks and flags
bit-exact; detector scores within 1e-5 relative.  In the default fused mode the scores are in fact
compared BIT-EXACTLY (score_rtol=0): kernel and oracle perform the same fp32 operation sequence per
output pixel.  All tests need a B200: run with  pytest -m gpu.
"""
import os

import numpy as np
import pytest

from _golden import DETECT_CASES, load_detect_case
from locomouse_cpp_b200 import synth
from locomouse_cpp_b200.types import FLAG_CAND_OVERFLOW, FLAG_DET_OVERFLOW, Config, Model, Results, diff_results

pytestmark = pytest.mark.gpu

SCORE_RTOL = 1e-5  # north_star tolerance for detector scores (only used across accumulation modes)


def _detector(cfg, model, bkg, calib):
    from locomouse_cpp_b200.api import Detector

    return Detector(cfg, model, bkg, calib, device=0)


def _both(oracle, spec, n, seed, first=0, **kw):
    total = first + n
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, total, seed=seed)
    frames = frames.numpy()
    prev = frames[first - 1] if first > 0 else None
    fr, bx, bs, bb = frames[first:], bx[first:], bs[first:], bb[first:]
    ref = oracle.detect(cfg, model, bkg, calib, fr, bx, bs, bb, prev_frame=prev, first_frame_index=first, n_threads=8)
    det = _detector(cfg, model, bkg, calib)
    got = det.detect_batch(fr, bx, bs, bb, prev_frame=prev, first_frame_index=first, **kw)
    return got, ref, det, (cfg, model, bkg, calib, fr, bx, bs, bb, prev)


# ---- golden fixtures ------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", DETECT_CASES)
def test_golden_case_bitexact(name):
    c = load_detect_case(name)
    det = _detector(c["cfg"], c["model"], c["bkg"], c["calib"])
    got = det.detect_batch(c["frames"], c["bb_x"], c["bb_y_side"], c["bb_y_bottom"], prev_frame=c["prev"],
                           first_frame_index=c["first"])
    assert diff_results(got, c["expected"]) == []
    assert got.checksum() == c["expected"].checksum()


# ---- config-1 geometry, all method variants -------------------------------------------------------------
@pytest.mark.parametrize("method", ["TM", "TM_DE", "base"])
def test_full_geometry_methods(oracle, method):
    got, ref, det, _ = _both(oracle, synth.SynthSpec(method=method), 6, seed=1000)
    assert diff_results(got, ref) == []
    assert int(ref.n_bottom.sum()) > 10 and int(ref.n_side.sum()) > 10 and int(ref.match_n.sum()) > 0
    assert (ref.tail[:, 0] >= 0).any()


@pytest.mark.parametrize("kw", [dict(flip=True), dict(warp=True, vid_pad=5), dict(conn=4), dict(flip=True, warp=True, conn=4)])
def test_full_geometry_variants(oracle, kw):
    got, ref, det, _ = _both(oracle, synth.SynthSpec(**kw), 4, seed=1003)
    assert diff_results(got, ref) == []
    assert int(ref.n_bottom.sum()) > 0


def test_mul_add_mode_is_bitexact_and_close_to_fused(oracle):
    """fma_mode=0 reproduces OpenCV's direct-path rounding (two roundings per tap) bit for bit; its scores
    agree with the fused mode within the north_star tolerance."""
    spec = synth.SynthSpec(fma_mode=False)
    got, ref, det, (cfg, model, bkg, calib, fr, bx, bs, bb, prev) = _both(oracle, spec, 4, seed=1000)
    assert diff_results(got, ref) == []
    cfg2 = synth.SynthSpec(fma_mode=True).config()
    fused = _detector(cfg2, model, bkg, calib).detect_batch(fr, bx, bs, bb)
    # same candidates whenever no score sits within rounding distance of a decision; compare scores loosely
    same = np.array_equal(fused.n_bottom, got.n_bottom) and np.array_equal(fused.n_side, got.n_side)
    if same:
        scale = float(np.abs(np.array(model.rho)).max())   # correlation magnitude: scores are sums of O(rho) minus rho
        assert diff_results(fused, got, score_rtol=SCORE_RTOL, score_atol=SCORE_RTOL * scale) == []


def test_mixed_template_shapes(oracle):
    """Non-square, odd/even, per-feature different template sizes (anchor = (cols/2, rows/2), halo = max over the
    view's templates, padded kernel widths 24/30/32/16)."""
    shapes = (((30, 30), (24, 28), (20, 16)), ((27, 30), (30, 22), (15, 17)))
    got, ref, det, _ = _both(oracle, synth.SynthSpec(tshapes=shapes), 4, seed=1001)
    assert diff_results(got, ref) == []
    assert int(ref.n_bottom.sum()) > 0


def test_config5_upsampled_60x60(oracle):
    """SURVEY config 5: 2x frames (800x3400), 60x60 templates, boxes 800 x 470 / 300."""
    spec = synth.SynthSpec(scale=2, det_cap=8192, cand_cap=128, match_cap=512)
    got, ref, det, _ = _both(oracle, spec, 2, seed=1000, allow_overflow=True)
    assert diff_results(got, ref) == []


# ---- temporal halo, sub-batches, memory spaces -------------------------------------------------------------
def test_prev_frame_halo_and_first_index(oracle):
    """Frames [3, 9) of a video with frame 2 as prev_frame equal the same frames of the whole-video run
    (velocity check reads the previous image; video frame 0 has none)."""
    spec = synth.SynthSpec()
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, 9, seed=1002)
    frames = frames.numpy()
    det = _detector(cfg, model, bkg, calib)
    whole = det.detect_batch(frames, bx, bs, bb)
    part = det.detect_batch(frames[3:], bx[3:], bs[3:], bb[3:], prev_frame=frames[2], first_frame_index=3)
    ref = oracle.detect(cfg, model, bkg, calib, frames, bx, bs, bb, n_threads=8)
    assert diff_results(whole, ref) == []
    for a in Results.ARRAYS:
        assert np.array_equal(getattr(part, a)[:6], getattr(whole, a)[3:9]), a
    with pytest.raises(ValueError):
        det.detect_batch(frames[3:], bx[3:], bs[3:], bb[3:], first_frame_index=3)  # prev_frame missing


def test_subbatch_split_and_device_frames_invariance(oracle, monkeypatch):
    """The same 11 frames through sub-batches of 4 (host frames, staged H2D) and as one device-resident
    batch give byte-identical results (checksum of checksums)."""
    import torch

    spec = synth.SynthSpec()
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, 11, seed=1004)
    fnp = frames.numpy()
    ref = oracle.detect(cfg, model, bkg, calib, fnp, bx, bs, bb, n_threads=8)
    monkeypatch.setenv("LM_SUBBATCH", "4")
    small = _detector(cfg, model, bkg, calib).detect_batch(fnp, bx, bs, bb)
    small_dev = _detector(cfg, model, bkg, calib).detect_batch(torch.from_numpy(fnp).cuda(), bx, bs, bb)
    monkeypatch.delenv("LM_SUBBATCH")
    big = _detector(cfg, model, bkg, calib).detect_batch(torch.from_numpy(fnp).cuda(), bx, bs, bb)
    pinned = _detector(cfg, model, bkg, calib).detect_batch(torch.from_numpy(fnp).pin_memory(), bx, bs, bb)
    assert diff_results(small, ref) == []
    assert small.checksum() == small_dev.checksum() == big.checksum() == pinned.checksum() == ref.checksum()


def test_scratch_grows_on_demand_and_short_host_videos_are_cut_in_two(oracle):
    """A short video in host memory runs as two sub-batches (second half's copy under the first half's kernels) in two small
    scratch sets; a later, longer resident call re-allocates the scratch once (all sets, bigger sub-batches) and a short call
    after that reuses it.  Results never change."""
    import torch

    spec = synth.SynthSpec()
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, 150, seed=1011)
    fnp = frames.numpy()
    ref = oracle.detect(cfg, model, bkg, calib, fnp, bx, bs, bb, n_threads=8)
    det = _detector(cfg, model, bkg, calib)
    det.set_option("subbatch", 256)
    a = det.detect_batch(fnp[:70], bx[:70], bs[:70], bb[:70])
    assert det.info("subbatch") == 64.0 and det.info("scratch_subbatch") == 64.0
    for name in Results.ARRAYS:
        assert np.array_equal(getattr(a, name)[:70], getattr(ref, name)[:70]), name
    b = det.detect_batch(frames.cuda(), bx, bs, bb)                       # 150 resident frames: sub-batches of 128, capacity grows
    assert det.info("subbatch") == 128.0 and det.info("scratch_subbatch") == 128.0
    assert diff_results(b, ref) == []
    c = det.detect_batch(fnp[:70], bx[:70], bs[:70], bb[:70])             # the small call again, in the big scratch
    assert det.info("subbatch") == 64.0 and det.info("scratch_subbatch") == 128.0
    for name in Results.ARRAYS:
        assert np.array_equal(getattr(c, name)[:70], getattr(a, name)[:70]), name
    d = det.detect_batch(fnp, bx, bs, bb)                                 # host frames: half the configured sub-batch
    assert det.info("subbatch") == 128.0
    assert diff_results(d, ref) == []
    det.close()
    torch.cuda.synchronize()


def test_tail_slow_path_equals_fast_path(oracle, monkeypatch):
    """k_tail labels runs in shared memory; frames with more runs than its capacity take the pixel-based
    global-memory kernel.  Force that path (capacity 4) and compare with the oracle and the fast path."""
    spec = synth.SynthSpec()
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, 5, seed=1009)
    fr = frames.numpy()
    ref = oracle.detect(cfg, model, bkg, calib, fr, bx, bs, bb, n_threads=5)
    fast = _detector(cfg, model, bkg, calib).detect_batch(fr, bx, bs, bb)
    monkeypatch.setenv("LM_TAIL_RUNCAP", "4")
    slow = _detector(cfg, model, bkg, calib).detect_batch(fr, bx, bs, bb)
    assert diff_results(fast, ref) == [] and diff_results(slow, ref) == []
    assert (ref.tail[:, 0] >= 0).any()


def test_repeatable(oracle):
    """Atomic-append order of detections must not leak into the results: 3 runs, identical bytes."""
    spec = synth.SynthSpec()
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, 5, seed=1005)
    det = _detector(cfg, model, bkg, calib)
    sums = {det.detect_batch(frames.numpy(), bx, bs, bb).checksum() for _ in range(3)}
    assert len(sums) == 1


# ---- edge cases --------------------------------------------------------------------------------------------
def test_empty_and_constant_frames(oracle):
    """No mouse (frame == background): nothing detected, tail all -1.  Constant frame: max == min ->
    normalisation scale 0 -> all zeros.  Saturated white frame."""
    spec = synth.SynthSpec()
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, 4, seed=1006)
    fr = frames.numpy().copy()
    fr[0] = bkg
    fr[1] = 0
    fr[2] = 255
    ref = oracle.detect(cfg, model, bkg, calib, fr, bx, bs, bb, n_threads=4)
    got = _detector(cfg, model, bkg, calib).detect_batch(fr, bx, bs, bb, allow_overflow=True)
    assert diff_results(got, ref) == []
    assert ref.n_bottom[1].sum() == 0 and (ref.tail[1] == -1).all()


def test_boxes_at_image_borders(oracle):
    """Boxes hanging over the left/top image edge (zero padded canvas) and touching the right edge."""
    spec = synth.SynthSpec()
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, 6, seed=1007)
    bx = np.array([14, 40, 399, 1000, 1690, 1699], np.uint32)   # 14 = leftmost valid corner for 30x30 templates
    bs = np.array([149, 164, 155, 164, 150, 164], np.uint32)
    fr = frames.numpy()
    ref = oracle.detect(cfg, model, bkg, calib, fr, bx, bs, bb, n_threads=6)
    got = _detector(cfg, model, bkg, calib).detect_batch(fr, bx, bs, bb, allow_overflow=True)
    assert diff_results(got, ref) == []


def test_roi_error_matches_reference_behaviour(oracle):
    """A corner left of spre-1 makes the reference's cv::Mat ROI throw; the library refuses the batch."""
    spec = synth.SynthSpec()
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, 2, seed=1008)
    det = _detector(cfg, model, bkg, calib)
    bad = bx.copy()
    bad[1] = 13
    assert oracle.check_roi(cfg, model, 13, bs[1], bb[1]) == -3
    with pytest.raises(RuntimeError, match="bounding box"):
        det.detect_batch(frames.numpy(), bad, bs, bb)
    bad[1] = 1700
    with pytest.raises(RuntimeError):
        det.detect_batch(frames.numpy(), bad, bs, bb)


def test_overflow_flags(oracle):
    """Tiny capacities: the per-frame flags say which list overflowed, as the oracle's do."""
    spec = synth.SynthSpec(cand_cap=2, match_cap=4)
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, 3, seed=1000)
    fr = frames.numpy()
    ref = oracle.detect(cfg, model, bkg, calib, fr, bx, bs, bb, n_threads=3)
    det = _detector(cfg, model, bkg, calib)
    from locomouse_cpp_b200.api import OverflowError_

    with pytest.raises(OverflowError_):
        det.detect_batch(fr, bx, bs, bb)
    got = det.detect_batch(fr, bx, bs, bb, allow_overflow=True)
    assert ((got.flags & FLAG_CAND_OVERFLOW) != 0).tolist() == ((ref.flags & FLAG_CAND_OVERFLOW) != 0).tolist()
    assert (got.flags & FLAG_CAND_OVERFLOW).any()
    # positive-pixel list overflow
    cfg2 = synth.SynthSpec(det_cap=16).config()
    got2 = _detector(cfg2, model, bkg, calib).detect_batch(fr, bx, bs, bb, allow_overflow=True)
    assert (got2.flags & FLAG_DET_OVERFLOW).any()


def test_zero_frames_and_bad_arguments():
    spec = synth.SynthSpec()
    cfg, model = spec.config(), synth.make_model(spec)
    bkg, calib = synth.make_background(spec), synth.make_calibration(spec)
    det = _detector(cfg, model, bkg, calib)
    empty = det.detect_batch(np.zeros((0, cfg.vid_rows, cfg.vid_cols), np.uint8), [], [], [])
    assert empty.n == 0
    with pytest.raises(ValueError):
        det.detect_batch(np.zeros((1, 10, 10), np.uint8), [1], [1], [1])
    with pytest.raises(RuntimeError):   # calibration index out of range: reference throws runtime_error (class.cpp:512-515)
        bad = calib.copy()
        bad[0, 0] = cfg.vid_rows * cfg.vid_cols
        det.set_calibration(bad)
    with pytest.raises(RuntimeError):   # background size mismatch (class.cpp:498-500)
        det.set_background(np.zeros((3, 3), np.uint8))


# ---- full BASELINE size through size-independent properties ---------------------------------------------------
def test_large_resident_batch_properties(oracle):
    """2048 device-rendered frames resident in HBM (the 10k-frame config is the same code path with more
    sub-batches; bench.py runs that size): (1) a random sample of frames, re-run on the CPU oracle with
    their true previous frames, matches bit for bit; (2) processing the batch in two halves with a halo
    frame gives the same bytes (shard invariance); (3) no overflow flags."""
    import torch

    spec = synth.SynthSpec()
    n = 2048
    cfg, model, bkg, calib, _, _, _, _ = synth.make_problem(spec, 8, seed=1000)
    frames, bx, bs, bb = synth.make_video(spec, n, 1000, "cuda", bkg)
    det = _detector(cfg, model, bkg, calib)
    whole = det.detect_batch(frames, bx, bs, bb)
    assert not whole.flags.any()
    h = n // 2 + 37
    a = det.detect_batch(frames[:h], bx[:h], bs[:h], bb[:h])
    b = det.detect_batch(frames[h:], bx[h:], bs[h:], bb[h:], prev_frame=frames[h - 1], first_frame_index=h)
    for name in Results.ARRAYS:
        assert np.array_equal(np.concatenate([getattr(a, name)[:h], getattr(b, name)[:n - h]]), getattr(whole, name)), name
    rng = np.random.Generator(np.random.PCG64(5))
    for f in sorted(rng.choice(np.arange(1, n), 24, replace=False).tolist()) + [0]:
        fr = frames[f:f + 1].cpu().numpy()
        prev = frames[f - 1].cpu().numpy() if f > 0 else None
        ref = oracle.detect(cfg, model, bkg, calib, fr, bx[f:f + 1], bs[f:f + 1], bb[f:f + 1], prev_frame=prev,
                            first_frame_index=f)
        for name in Results.ARRAYS:
            assert np.array_equal(getattr(ref, name)[0], getattr(whole, name)[f]), (f, name)


def test_pinned_and_pageable_result_buffers_agree():
    """Results in page-locked memory are filled by direct D2H copies, pageable ones through the library's staging buffers
    and a host memcpy; 2-, 3- and 4-slot pipelines and the screen's stream priority never change a byte."""
    from locomouse_cpp_b200.api import Detector
    from locomouse_cpp_b200.types import Results

    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(synth.SynthSpec(), 23, seed=1005)
    frames = frames.numpy()
    sums = set()
    for streams, prio, pinned in ((2, 1, True), (2, 1, False), (3, 0, True), (4, 1, False), (1, 1, True)):
        det = Detector(cfg, model, bkg, calib, device=0)
        det.set_option("subbatch", 4)
        det.set_option("streams", streams)
        det.set_option("screen_priority", prio)
        res = Results(len(frames), cfg.cand_cap, cfg.match_cap, cfg.n_tail_points, pinned=pinned)
        for _ in range(2):
            det.detect_batch(frames, bx, bs, bb, results=res)
        sums.add(res.checksum())
        det.close()
    assert len(sums) == 1


def test_results_in_lm_host_alloc_memory():
    """lm_host_alloc / lm_host_free: result arrays carved from the library's page-locked allocation receive the same bytes
    as ordinary numpy arrays."""
    import ctypes as C

    from locomouse_cpp_b200 import api
    from locomouse_cpp_b200.api import Detector

    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(synth.SynthSpec(), 5, seed=1006)
    frames = frames.numpy()
    L = api.load_library()
    nbytes = Results.raw_nbytes(len(frames), cfg.cand_cap, cfg.match_cap, cfg.n_tail_points)
    p = C.c_void_p()
    L.lm_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.lm_host_free.argtypes = [C.c_void_p]
    assert L.lm_host_alloc(C.byref(p), nbytes) == 0 and p.value
    try:
        buf = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_ubyte)), shape=(nbytes,))
        buf[:] = 0xAB
        pinned = Results(len(frames), cfg.cand_cap, cfg.match_cap, cfg.n_tail_points, buffer=buf)
        assert pinned.raw.ctypes.data == p.value          # adopted, not copied
        det = Detector(cfg, model, bkg, calib, device=0)
        det.detect_batch(frames, bx, bs, bb, results=pinned)
        plain = det.detect_batch(frames, bx, bs, bb, results=Results(len(frames), cfg.cand_cap, cfg.match_cap, cfg.n_tail_points))
        det.close()
        assert diff_results(pinned, plain) == []
        del pinned, buf
    finally:
        assert L.lm_host_free(p) == 0
    assert L.lm_host_free(None) == 0


# ---- API robustness (round-1 advisor findings) -------------------------------------------------------------------------
def test_large_candidate_capacity_runs_the_pairing_kernel(oracle):
    """cand_cap = 512 makes k_pair's per-warp arrays exceed the default 48 kB of dynamic shared memory: the launcher opts in."""
    spec = synth.SynthSpec(cand_cap=512, match_cap=2048)
    got, ref, det, _ = _both(oracle, spec, 3, seed=1000)
    assert diff_results(got, ref) == []
    assert int(ref.match_n.sum()) > 0


def test_reconfigure_keeps_model_and_rederives_geometry(oracle):
    """lm_configure with a new box size after lm_set_model: pads / canvas (lm_get_geometry) and the ROI check follow the new
    configuration, and detection with the re-sent static inputs equals the oracle."""
    import ctypes as C

    spec_a, spec_b = synth.SynthSpec(method="TM"), synth.SynthSpec(method="TM_DE")
    cfg_a, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec_a, 3, seed=1000)
    cfg_b, model_b, bkg_b, calib_b, frames_b, bx_b, bs_b, bb_b = synth.make_problem(spec_b, 3, seed=1000)
    det = _detector(cfg_a, model, bkg, calib)
    canvas = (C.c_int32 * 4)()
    pads = (C.c_int32 * 8)()
    assert det._L.lm_get_geometry(det._ctx, pads, canvas) == 0
    assert canvas[1] == max(cfg_a.bb_h_side, 15)
    c = cfg_b.to_c()
    det._check(det._L.lm_configure(det._ctx, C.byref(c)))          # model kept, geometry re-derived
    assert det._L.lm_get_geometry(det._ctx, pads, canvas) == 0
    assert canvas[1] == max(cfg_b.bb_h_side, 15) and cfg_a.bb_h_side != cfg_b.bb_h_side
    det.cfg = cfg_b
    det.set_model(model_b)
    det.set_background(bkg_b)
    det.set_calibration(calib_b)
    got = det.detect_batch(frames_b.numpy(), bx_b, bs_b, bb_b)
    ref = oracle.detect(cfg_b, model_b, bkg_b, calib_b, frames_b.numpy(), bx_b, bs_b, bb_b, n_threads=4)
    assert diff_results(got, ref) == []


def test_null_result_array_is_an_argument_error_and_device_is_restored():
    import ctypes as C

    import torch

    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(synth.SynthSpec(), 2, seed=1000)
    det = _detector(cfg, model, bkg, calib)
    res = Results(2, cfg.cand_cap, cfg.match_cap, cfg.n_tail_points)
    r = res.to_c()
    r.match_s = None
    fr = np.ascontiguousarray(frames.numpy())
    bx, bs, bb = (np.ascontiguousarray(a, np.uint32) for a in (bx, bs, bb))
    rc = det._L.lm_detect_batch(det._ctx, fr.ctypes.data, 0, None, 2, 0, bx.ctypes.data, bs.ctypes.data, bb.ctypes.data, C.byref(r))
    assert rc == -1 and b"lm_results" in det._L.lm_last_error(det._ctx)
    if torch.cuda.device_count() > 1:   # the library runs on its own device and leaves the caller's current device alone
        torch.cuda.set_device(1)
        det.detect_batch(fr, bx, bs, bb)
        assert torch.cuda.current_device() == 1
        torch.cuda.set_device(0)
