"""ctypes wrapper of the CPU oracle (oracle/liblm_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under locomouse_cpp_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from locomouse_cpp_b200.types import (CAND_DTYPE, Config, Model, Results, lm_config, lm_results, lm_template)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblm_oracle.so")
_lib = None

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_u32p = C.POINTER(C.c_uint32)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile if the .so is missing or older than its sources."""
    srcs = [os.path.join(_HERE, "lm_oracle.cpp"), os.path.join(_HERE, "lm_oracle.h"),
            os.path.join(_HERE, "..", "include", "locomouse_b200.h")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs if os.path.exists(s))
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liblm_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.lmo_detect.restype = C.c_int
        L.lmo_detect.argtypes = [C.POINTER(lm_config), C.c_void_p, _u8p, _i32p, _u8p, _u8p, C.c_int64, C.c_int64,
                                 _u32p, _u32p, _u32p, C.POINTER(lm_results), C.c_int, _f64p]
        L.lmo_geometry.restype = C.c_int
        L.lmo_geometry.argtypes = [C.POINTER(lm_config), C.c_void_p, _i32p, _i32p]
        L.lmo_check_roi.restype = C.c_int
        L.lmo_check_roi.argtypes = [C.POINTER(lm_config), C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.lmo_preprocess.restype = C.c_int
        L.lmo_preprocess.argtypes = [C.POINTER(lm_config), _u8p, _i32p, _u8p, _u8p, _i32p]
        L.lmo_imadjust_lut.restype = None
        L.lmo_imadjust_lut.argtypes = [C.c_double] * 4 + [_u8p]
        L.lmo_correlate.restype = None
        L.lmo_correlate.argtypes = [_u8p, C.c_int32, C.c_int32, C.POINTER(lm_template), C.c_int32, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_int32, _f32p]
        for fn in (L.lmo_nms_max, L.lmo_peak_clustering):
            fn.restype = C.c_int
            fn.argtypes = [_f32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]
        L.lmo_largest_region.restype = None
        L.lmo_largest_region.argtypes = [_u8p, C.c_int32, C.c_int32, C.c_int32, _u8p]
        L.lmo_tail_from_binary.restype = None
        L.lmo_tail_from_binary.argtypes = [_u8p, C.c_int32, _u8p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                           _i32p, _u8p]
        L.lmo_match_views.restype = C.c_int
        L.lmo_match_views.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_int32, C.c_double, _u8p, _u8p, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_int32, C.c_int32, _i32p, _i32p, _f64p,
                                      C.c_int32, _i32p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def detect(cfg: Config, model: Model, bkg, calib, frames, bb_x, bb_y_side, bb_y_bottom, prev_frame=None,
           first_frame_index: int = 0, n_threads: int = 1, stage_seconds=None) -> Results:
    """Oracle run of the whole path on host arrays; same argument meaning as Detector.detect_batch."""
    frames = _u8(frames)
    n = frames.shape[0]
    bkg = _u8(bkg)
    calib = np.ascontiguousarray(calib, dtype=np.int32)
    bb_x = np.ascontiguousarray(bb_x, dtype=np.uint32)
    bb_y_side = np.ascontiguousarray(bb_y_side, dtype=np.uint32)
    bb_y_bottom = np.ascontiguousarray(bb_y_bottom, dtype=np.uint32)
    prev = _u8(prev_frame) if prev_frame is not None else None
    res = Results(n, cfg.cand_cap, cfg.match_cap, cfg.n_tail_points)
    c = cfg.to_c()
    t = model.to_c()
    r = res.to_c()
    st = np.zeros(6, np.float64)
    rc = lib().lmo_detect(C.byref(c), C.cast(t, C.c_void_p), _p(bkg, _u8p), _p(calib, _i32p), _p(frames, _u8p),
                          _p(prev, _u8p), n, first_frame_index, _p(bb_x, _u32p), _p(bb_y_side, _u32p),
                          _p(bb_y_bottom, _u32p), C.byref(r), n_threads, _p(st, _f64p))
    if stage_seconds is not None:
        stage_seconds[:] = st
    res.rc = rc
    if rc not in (0, -4):
        raise RuntimeError(f"oracle lmo_detect failed: {rc}")
    return res


def geometry(cfg: Config, model: Model):
    pads = np.zeros(8, np.int32)
    canvas = np.zeros(4, np.int32)
    c, t = cfg.to_c(), model.to_c()
    rc = lib().lmo_geometry(C.byref(c), C.cast(t, C.c_void_p), _p(pads, _i32p), _p(canvas, _i32p))
    assert rc == 0
    return pads, canvas


def check_roi(cfg: Config, model: Model, bb_x, bb_y_side, bb_y_bottom) -> int:
    c, t = cfg.to_c(), model.to_c()
    return lib().lmo_check_roi(C.byref(c), C.cast(t, C.c_void_p), int(bb_x), int(bb_y_side), int(bb_y_bottom))


def preprocess(cfg: Config, bkg, calib, frame):
    bkg, frame = _u8(bkg), _u8(frame)
    calib = np.ascontiguousarray(calib, dtype=np.int32)
    out = np.zeros((cfg.n_rows, cfg.n_cols), np.uint8)
    mm = np.zeros(2, np.int32)
    c = cfg.to_c()
    lib().lmo_preprocess(C.byref(c), _p(bkg, _u8p), _p(calib, _i32p), _p(frame, _u8p), _p(out, _u8p), _p(mm, _i32p))
    return out, mm


def imadjust_lut(low_in=0.0, high_in=0.6, low_out=0.0, high_out=1.0):
    lut = np.zeros(256, np.uint8)
    lib().lmo_imadjust_lut(low_in, high_in, low_out, high_out, _p(lut, _u8p))
    return lut


def correlate(I, w, rho, x0, y0, width, height, fma_mode=True):
    I = _u8(I)
    w = np.ascontiguousarray(w, dtype=np.float32)
    t = lm_template(w.ctypes.data_as(_f32p), w.shape[0], w.shape[1], float(rho))
    out = np.zeros((height, width), np.float32)
    lib().lmo_correlate(_p(I, _u8p), I.shape[0], I.shape[1], C.byref(t), x0, y0, width, height, int(fma_mode),
                        _p(out, _f32p))
    return out


def _nms(fn, scores, box_w, box_h, cap):
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    out = np.zeros(cap, CAND_DTYPE)
    n = fn(_p(scores, _f32p), scores.shape[0], scores.shape[1], box_w, box_h, out.ctypes.data, cap)
    return [(int(c["x"]), int(c["y"]), float(c["s"])) for c in out[: min(n, cap)]], n


def nms_max(scores, box_w, box_h, cap=4096):
    return _nms(lib().lmo_nms_max, scores, box_w, box_h, cap)[0]


def peak_clustering(scores, box_w, box_h, cap=4096):
    return _nms(lib().lmo_peak_clustering, scores, box_w, box_h, cap)[0]


def largest_region(binary, conn=8):
    b = _u8(binary)
    out = np.zeros_like(b)
    lib().lmo_largest_region(_p(b, _u8p), b.shape[0], b.shape[1], conn, _p(out, _u8p))
    return out


def tail_from_binary(bin_bottom, bin_side, conn=8, n_points=15):
    bb, bs = _u8(bin_bottom), _u8(bin_side)
    assert bb.shape[1] == bs.shape[1]
    tracks = np.zeros((3, n_points), np.int32)
    mask = np.zeros_like(bb)
    lib().lmo_tail_from_binary(_p(bb, _u8p), bb.shape[0], _p(bs, _u8p), bs.shape[0], bb.shape[1], conn, n_points,
                               _p(tracks, _i32p), _p(mask, _u8p))
    return tracks, mask


def match_views(cb, cs, vel_check, tsize_b, tsize_s, T, I=None, Iprev=None, x0=0, y0b=0, y0s=0, match_cap=1024):
    """cb / cs: lists of (x, y, s); tsize_* = (cols, rows) of the feature's templates."""
    a = np.array([tuple(c) for c in cb], CAND_DTYPE) if len(cb) else np.zeros(0, CAND_DTYPE)
    b = np.array([tuple(c) for c in cs], CAND_DTYPE) if len(cs) else np.zeros(0, CAND_DTYPE)
    if I is None:
        I = np.zeros((1, 1), np.uint8)
    if Iprev is None:
        Iprev = np.zeros_like(I)
    I, Iprev = _u8(I), _u8(Iprev)
    mn = np.zeros(max(len(cb), 1), np.int32)
    my = np.full(match_cap, -1, np.int32)
    ms = np.full(match_cap, -1.0, np.float64)
    nm = np.zeros(1, np.int32)
    lib().lmo_match_views(a.ctypes.data, len(cb), b.ctypes.data, len(cs), int(vel_check), tsize_b[0], tsize_b[1],
                          tsize_s[0], tsize_s[1], float(T), _p(I, _u8p), _p(Iprev, _u8p), I.shape[0], I.shape[1],
                          x0, y0b, y0s, _p(mn, _i32p), _p(my, _i32p), _p(ms, _f64p), match_cap, _p(nm, _i32p))
    out, o = [], 0
    for i in range(len(cb)):
        m = int(mn[i])
        out.append([(int(my[o + j]), float(ms[o + j])) for j in range(m)])
        o += m
    return out


# ---- pass 1, LocoMouse_TM_DE ------------------------------------------------------------------------
def bounding_box_tm_de(cfg: Config, bkg, calib, frames, params, window: int = 5):
    """(BB_X_POS uint32[n], raw bb_x float64[n], lims int32[n, 2]) as LocoMouse_TM_DE::computeBoundingBox produces them."""
    L = lib()
    frames = _u8(frames)
    n = frames.shape[0]
    c = cfg.to_c()
    raw = np.zeros(n, np.float64)
    lims = np.zeros((n, 2), np.int32)
    L.lmo_bounding_box_tm_de.restype = C.c_int
    rc = L.lmo_bounding_box_tm_de(C.byref(c), _p(_u8(bkg), _u8p), _p(np.ascontiguousarray(calib, np.int32), _i32p), _p(frames, _u8p),
                                  C.c_int64(n), C.byref(params), raw.ctypes.data_as(_f64p), lims.ctypes.data_as(_i32p))
    if rc != 0:
        raise ValueError(f"lmo_bounding_box_tm_de failed ({rc})")
    out = np.zeros(n, np.uint32)
    L.lmo_vecmovingaverage(raw.ctypes.data_as(_f64p), C.c_int64(n), C.c_int32(window), out.ctypes.data_as(_u32p))
    return out, raw, lims


def bounding_box_base(cfg: Config, bkg, calib, frames, params):
    """Per-frame output of LocoMouse::computeMouseBox after the base readFrame (LocoMouse_class.cpp:579-631, 921-997):
    (box float64[n, 6], lims int32[n, 4, 2])."""
    L = lib()
    frames = _u8(frames)
    n = frames.shape[0]
    c = cfg.to_c()
    box = np.zeros((n, 6), np.float64)
    lims = np.zeros((n, 4, 2), np.int32)
    L.lmo_bounding_box_base.restype = C.c_int
    rc = L.lmo_bounding_box_base(C.byref(c), _p(_u8(bkg), _u8p), _p(np.ascontiguousarray(calib, np.int32), _i32p), _p(frames, _u8p),
                                 C.c_int64(n), C.byref(params), box.ctypes.data_as(_f64p), lims.ctypes.data_as(_i32p))
    if rc != 0:
        raise ValueError(f"lmo_bounding_box_base failed ({rc})")
    return box, lims


def bounding_box_tm(cfg: Config, bkg, calib, frames, params):
    """Per-frame output of LocoMouse_TM::computeMouseBox_DD after the base readFrame (LocoMouse_TM.cpp:115-269):
    (raw bb_x float64[n], lims int32[n, 2])."""
    L = lib()
    frames = _u8(frames)
    n = frames.shape[0]
    c = cfg.to_c()
    raw = np.zeros(n, np.float64)
    lims = np.zeros((n, 2), np.int32)
    L.lmo_bounding_box_tm.restype = C.c_int
    rc = L.lmo_bounding_box_tm(C.byref(c), _p(_u8(bkg), _u8p), _p(np.ascontiguousarray(calib, np.int32), _i32p), _p(frames, _u8p),
                               C.c_int64(n), C.byref(params), raw.ctypes.data_as(_f64p), lims.ctypes.data_as(_i32p))
    if rc != 0:
        raise ValueError(f"lmo_bounding_box_tm failed ({rc})")
    return raw, lims


def mouse_box_tm(side_view, disk, threshold=3, min_pixel_count=10, min_pixel_visible=1, conn=8, zero=(0, None, 0, None), sums_as_float=1):
    """computeMouseBox_DD on a calibrated side view (the whole image IS the side view) -> (bb_x, lims int32[2], stages) with
    stages = adjusted / binary / opened / filtered images and row_sums, as oracle/reference_nms.mouse_box_dd reports them."""
    from locomouse_cpp_b200.types import lm_bb_tm_params

    L = lib()
    I = _u8(side_view)
    rows, cols = I.shape
    dk = np.ascontiguousarray(disk, np.float32)
    z = [int(zero[0]), cols if zero[1] is None else int(zero[1]), int(zero[2]), rows if zero[3] is None else int(zero[3])]
    p = lm_bb_tm_params(side_x=0, side_y=0, side_w=cols, side_h=rows, side_threshold=int(threshold), min_pixel_count=int(min_pixel_count),
                        min_pixel_visible=int(min_pixel_visible), zero_col_pre=z[0], zero_col_post=z[1], zero_row_pre=z[2], zero_row_post=z[3],
                        sums_as_float=int(sums_as_float), disk_size=dk.shape[0], reserved=0)
    p.disk = dk.ctypes.data_as(C.POINTER(C.c_float))
    st = {k: np.zeros((rows, cols), np.uint8) for k in ("adjusted", "binary", "opened", "filtered")}
    st["row_sums"] = np.zeros(cols, np.int32)
    bbx = C.c_double(0.0)
    lims = np.zeros(2, np.int32)
    L.lmo_mouse_box_tm.restype = C.c_int
    rc = L.lmo_mouse_box_tm(_p(I, _u8p), rows, cols, int(conn), C.byref(p), C.byref(bbx), lims.ctypes.data_as(_i32p),
                            _p(st["adjusted"], _u8p), _p(st["binary"], _u8p), _p(st["opened"], _u8p), _p(st["filtered"], _u8p),
                            st["row_sums"].ctypes.data_as(_i32p))
    if rc != 0:
        raise ValueError(f"lmo_mouse_box_tm failed ({rc})")
    return float(bbx.value), lims, st


def filter2d_u8(image, kernel):
    """8-bit -> 8-bit filter2D, BORDER_REPLICATE, as the oracle's TM pass 1 evaluates it (the "filtered" stage with every other
    stage made transparent: threshold 0 on a 0 / 1 image, min_pixel_count 1 keeps everything)."""
    img = _u8(image)
    assert img.max() <= 1
    # imadjust_default would rescale the image: run the stages from "opened" on by feeding a binary image whose histogram maps
    # to itself is not possible in general, so the filter is evaluated through the dedicated entry point
    L = lib()
    dk = np.ascontiguousarray(kernel, np.float32)
    out = np.zeros_like(img)
    L.lmo_filter2d_u8.restype = None
    L.lmo_filter2d_u8(_p(img, _u8p), img.shape[0], img.shape[1], dk.ctypes.data_as(C.POINTER(C.c_float)), dk.shape[0], _p(out, _u8p))
    return out


def mouse_box_base(image, conn, params):
    """computeMouseBox on an already pre-processed calibrated image -> (box float64[6], lims int32[4, 2])."""
    L = lib()
    I = _u8(image)
    box = np.zeros(6, np.float64)
    lims = np.zeros((4, 2), np.int32)
    L.lmo_mouse_box_base.restype = C.c_int
    rc = L.lmo_mouse_box_base(_p(I, _u8p), I.shape[0], I.shape[1], int(conn), C.byref(params), box.ctypes.data_as(_f64p), lims.ctypes.data_as(_i32p))
    if rc != 0:
        raise ValueError(f"lmo_mouse_box_base failed ({rc})")
    return box, lims


def mouse_box_size(w, hb, hs):
    """computeMouseBoxSize (LocoMouse_class.cpp:1481-1506) -> (final_w, final_hb, final_hs)."""
    L = lib()
    a, b, c = (np.array(v, np.float64, copy=True) for v in (w, hb, hs))
    size = np.zeros(3, np.int32)
    L.lmo_mouse_box_size.restype = None
    L.lmo_mouse_box_size(a.ctypes.data_as(_f64p), b.ctypes.data_as(_f64p), c.ctypes.data_as(_f64p), C.c_int64(a.size), size.ctypes.data_as(_i32p))
    return tuple(int(v) for v in size)


def imadjust_default_lut(hist):
    L = lib()
    h = np.ascontiguousarray(hist, np.uint32)
    lut = np.zeros(256, np.uint8)
    mm = np.zeros(2, np.int32)
    L.lmo_imadjust_default_lut(h.ctypes.data_as(_u32p), lut.ctypes.data_as(_u8p), mm.ctypes.data_as(_i32p))
    return lut, (int(mm[0]), int(mm[1]))


def first_last_over_t(values, th: int):
    L = lib()
    v = np.ascontiguousarray(values, np.float32)
    fl = np.zeros(2, np.int32)
    L.lmo_first_last_over_t(v.ctypes.data_as(_f32p), C.c_uint32(v.size), C.c_int32(th), fl.ctypes.data_as(_i32p))
    return int(fl[0]), int(fl[1])


def vecmovingaverage(v, window: int):
    L = lib()
    a = np.ascontiguousarray(v, np.float64)
    out = np.zeros(a.size, np.uint32)
    L.lmo_vecmovingaverage(a.ctypes.data_as(_f64p), C.c_int64(a.size), C.c_int32(window), out.ctypes.data_as(_u32p))
    return out


# ---- cost builders -----------------------------------------------------------------------------------
def unary_cost_box(cands, bb_w, bb_h, priors):
    """cands: list of (x, y, s); priors: ctypes array of lm_location_prior -> MyMat as an (n, n_priors) array."""
    c = np.array([tuple(k) for k in cands], CAND_DTYPE) if len(cands) else np.zeros(0, CAND_DTYPE)
    out = np.zeros((len(priors), max(len(c), 1)), np.float64)
    L = lib()
    L.lmo_unary_cost_box.restype = None
    L.lmo_unary_cost_box(C.c_void_p(c.ctypes.data), len(c), int(bb_w), int(bb_h), priors, len(priors), C.c_void_p(out.ctypes.data))
    return out.reshape(-1)[:len(priors) * len(c)].reshape(len(priors), len(c)).T.copy()


def pairwise_potential(ci, cip1, params, cap=1 << 16):
    """-> (n_rows, n_cols, jc, ir, pr) of the MATSPARSE pairwisePotential returns."""
    a = np.array([tuple(k) for k in ci], CAND_DTYPE) if len(ci) else np.zeros(0, CAND_DTYPE)
    b = np.array([tuple(k) for k in cip1], CAND_DTYPE) if len(cip1) else np.zeros(0, CAND_DTYPE)
    nong = params.ong_w * params.ong_h
    jc = np.zeros(len(a) + nong + 1, np.int32)
    ir = np.zeros(cap, np.int32)
    pr = np.zeros(cap, np.float64)
    dims = np.zeros(3, np.int32)
    L = lib()
    L.lmo_pairwise_potential.restype = C.c_int
    rc = L.lmo_pairwise_potential(C.c_void_p(a.ctypes.data), len(a), C.c_void_p(b.ctypes.data), len(b), C.byref(params),
                                  C.c_void_p(jc.ctypes.data), C.c_void_p(ir.ctypes.data), C.c_void_p(pr.ctypes.data), C.c_int64(cap),
                                  C.c_void_p(dims.ctypes.data))
    assert rc == 0
    return int(dims[0]), int(dims[1]), jc, ir[:dims[2]].copy(), pr[:dims[2]].copy()


COVERAGE_NAMES = ("pairings", "all_equal_boolD_zeroed", "velocity_comparisons", "velocity_rejections", "velocity_accepts",
                  "moving_windows", "bottom_without_match", "side_matches")


def coverage(reset: bool = False) -> dict:
    """Branch-coverage counters of the oracle's pairing stage since the last reset."""
    L = lib()
    out = np.zeros(8, np.int64)
    L.lmo_coverage(out.ctypes.data_as(C.POINTER(C.c_int64)), int(reset))
    return dict(zip(COVERAGE_NAMES, map(int, out)))
