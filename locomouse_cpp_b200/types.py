"""ctypes mirrors of the POD types in include/locomouse_b200.h plus numpy result buffers.

Only plain data lives here (no compute): `Config` <-> `lm_config`, `Template` <-> `lm_template`,
`Results` <-> `lm_results` (struct-of-arrays, caller allocated).  The record layouts follow the
reference's value types: `lm_cand` = Candidate{Point_<int> p; double s} (Candidates/Candidates.hpp:16-34),
the match arrays = P22D's yt/st vectors (Candidates/Candidates.hpp:63-105).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

PAW, SNOUT, TAIL = 0, 1, 2
BOTTOM, SIDE = 0, 1

LM_OK = 0
LM_ERR_INVALID = -1
LM_ERR_RUNTIME = -2
LM_ERR_ROI = -3
LM_ERR_OVERFLOW = -4
LM_ERR_STATE = -5

FLAG_DET_OVERFLOW = 0x1
FLAG_CAND_OVERFLOW = 0x2
FLAG_MATCH_OVERFLOW = 0x4

CAND_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("s", "<f8")])
assert CAND_DTYPE.itemsize == 16


class lm_template(C.Structure):
    _fields_ = [("w", C.POINTER(C.c_float)), ("rows", C.c_int32), ("cols", C.c_int32), ("rho", C.c_double)]


class lm_config(C.Structure):
    _fields_ = [
        ("vid_rows", C.c_int32), ("vid_cols", C.c_int32),
        ("n_rows", C.c_int32), ("n_cols", C.c_int32),
        ("bb_w", C.c_int32), ("bb_h_bottom", C.c_int32), ("bb_h_side", C.c_int32),
        ("tail_w", C.c_int32), ("flip", C.c_int32), ("imadjust", C.c_int32),
        ("conn", C.c_int32), ("n_tail_points", C.c_int32),
        ("min_overlap", C.c_double),
        ("fma_mode", C.c_int32), ("cand_cap", C.c_int32), ("det_cap", C.c_int32), ("match_cap", C.c_int32),
    ]


class lm_results(C.Structure):
    _fields_ = [
        ("n_frames", C.c_int64),
        ("cand_cap", C.c_int32), ("match_cap", C.c_int32), ("n_tail_points", C.c_int32),
        ("n_bottom", C.c_void_p), ("n_side", C.c_void_p),
        ("bottom", C.c_void_p), ("side", C.c_void_p),
        ("match_n", C.c_void_p), ("match_y", C.c_void_p), ("match_s", C.c_void_p),
        ("tail", C.c_void_p), ("flags", C.c_void_p),
    ]


class lm_bb_de_params(C.Structure):
    _fields_ = [("side_x", C.c_int32), ("side_y", C.c_int32), ("side_w", C.c_int32), ("side_h", C.c_int32),
                ("zero_col_pre", C.c_int32), ("zero_col_post", C.c_int32), ("zero_row_pre", C.c_int32), ("zero_row_post", C.c_int32),
                ("threshold", C.c_double), ("min_count", C.c_int32), ("width_margin", C.c_double)]


def bb_de_params(cfg, side_h: int = 165, **kw) -> lm_bb_de_params:
    """LocoMouse_TM_DE defaults (LocoMouse_TM_DE.hpp:27-29, LocoMouse_TM_DE.cpp:68-71) for a side view that spans the
    upper `side_h` rows of the calibrated image."""
    d = dict(side_x=0, side_y=0, side_w=cfg.n_cols, side_h=side_h, zero_col_pre=46, zero_col_post=760, zero_row_pre=100,
             zero_row_post=149, threshold=255 * 0.05, min_count=10, width_margin=1.1)
    d.update(kw)
    return lm_bb_de_params(**d)


class lm_bb_base_params(C.Structure):
    _fields_ = [("side_x", C.c_int32), ("side_y", C.c_int32), ("side_w", C.c_int32), ("side_h", C.c_int32),
                ("bottom_x", C.c_int32), ("bottom_y", C.c_int32), ("bottom_w", C.c_int32), ("bottom_h", C.c_int32),
                ("median_filter_size", C.c_int32), ("min_pixel_visible", C.c_int32), ("sums_as_float", C.c_int32), ("reserved", C.c_int32)]


def bb_base_params(cfg, side_h: int = 165, **kw) -> lm_bb_base_params:
    """Base-class pass-1 defaults (LocoMouse_class.hpp:54-55) for views that split the calibrated image at row `side_h`;
    sums_as_float = 1 is the reference's behaviour (see include/locomouse_b200.h)."""
    d = dict(side_x=0, side_y=0, side_w=cfg.n_cols, side_h=side_h, bottom_x=0, bottom_y=side_h, bottom_w=cfg.n_cols,
             bottom_h=cfg.n_rows - side_h, median_filter_size=11, min_pixel_visible=1, sums_as_float=1, reserved=0)
    d.update(kw)
    return lm_bb_base_params(**d)


class lm_bb_tm_params(C.Structure):
    _fields_ = [("side_x", C.c_int32), ("side_y", C.c_int32), ("side_w", C.c_int32), ("side_h", C.c_int32),
                ("side_threshold", C.c_int32), ("min_pixel_count", C.c_int32), ("min_pixel_visible", C.c_int32),
                ("zero_col_pre", C.c_int32), ("zero_col_post", C.c_int32), ("zero_row_pre", C.c_int32), ("zero_row_post", C.c_int32),
                ("sums_as_float", C.c_int32), ("disk_size", C.c_int32), ("reserved", C.c_int32), ("disk", C.POINTER(C.c_float))]


def bb_tm_params(cfg, disk, side_h: int = 165, **kw) -> lm_bb_tm_params:
    """LocoMouse_TM pass-1 parameters (LocoMouse_TM.cpp:44-113 reads them from the configuration file; DISK_FILTER from
    diskfilter.yml) for a side view that spans the upper `side_h` rows.  `disk` (square float32 matrix) is kept alive on the
    returned structure.  sums_as_float = 1 is the reference's behaviour (see include/locomouse_b200.h)."""
    import numpy as np

    dk = np.ascontiguousarray(disk, np.float32)
    assert dk.ndim == 2 and dk.shape[0] == dk.shape[1]
    d = dict(side_x=0, side_y=0, side_w=cfg.n_cols, side_h=side_h, side_threshold=3, min_pixel_count=10, min_pixel_visible=1,
             zero_col_pre=0, zero_col_post=cfg.n_cols, zero_row_pre=0, zero_row_post=side_h, sums_as_float=1, disk_size=dk.shape[0], reserved=0)
    d.update(kw)
    p = lm_bb_tm_params(**d)
    p.disk = dk.ctypes.data_as(C.POINTER(C.c_float))
    p._keep = dk
    return p


class lm_location_prior(C.Structure):
    _fields_ = [("pos_x", C.c_double), ("pos_y", C.c_double), ("max_distance", C.c_double), ("area_x", C.c_double),
                ("area_y", C.c_double), ("area_w", C.c_double), ("area_h", C.c_double)]


class lm_pairwise_params(C.Structure):
    _fields_ = [("grid_x", C.c_double), ("grid_y", C.c_double), ("grid_spacing", C.c_double), ("ong_w", C.c_int32),
                ("ong_h", C.c_int32), ("max_displacement", C.c_double), ("alpha_vel", C.c_double), ("occluded_cost", C.c_double)]


def location_priors(rows):
    """rows: (x, y, max_distance, min_x, max_x, min_y, max_y) per prior -- the arguments of the reference's
    LocoMouse_LocationPrior constructor (LocoMouse_class.cpp:3196-3202) -> ctypes array of lm_location_prior."""
    arr = (lm_location_prior * len(rows))()
    for k, (x, y, md, x0, x1, y0, y1) in enumerate(rows):
        if not (x0 < x1 and y0 < y1):
            raise ValueError("location prior: min must be below max (CV_Assert in the reference)")
        arr[k] = lm_location_prior(x, y, md, x0, y0, x1 - x0, y1 - y0)
    return arr


def pairwise_params(bb_w: int, bb_h: int, spacing: int = 20, max_width: float = 0.75, max_displacement: float = 15,
                    alpha_vel: float = 1e-1, occluded_cost: float = 1e-2) -> lm_pairwise_params:
    """Occlusion grid and costs as LocoMouse::initializeFeatureLoop derives them (LocoMouse_class.cpp:723-733) from the
    reference's default parameters (LocoMouse_class.hpp:59-68): integer divisions as there."""
    ngrid_y = (bb_h - spacing) // spacing + 1
    ngrid_x = int((max_width * bb_w - spacing) / spacing + 1)
    return lm_pairwise_params(float(bb_w - 1 - spacing // 2), float(bb_h - 1 - spacing // 2), float(spacing), ngrid_x, ngrid_y,
                              float(max_displacement), float(alpha_vel), float(occluded_cost))


TemplateArray = (lm_template * 3) * 2  # t[view][feature]


@dataclass
class Config:
    """Scalars that reach the kernels (SURVEY.md §5 'Config / flags').  Defaults = reference defaults
    for LocoMouse_TM at the config-1 geometry (LocoMouse_TM.hpp:31-32, LocoMouse_class.hpp:53-86)."""

    vid_rows: int = 400
    vid_cols: int = 1700
    n_rows: int = 400
    n_cols: int = 1700
    bb_w: int = 400
    bb_h_bottom: int = 235
    bb_h_side: int = 150
    tail_sub_bounding_box: float = 0.6
    flip: bool = False
    imadjust: bool = True
    conn: int = 8
    n_tail_points: int = 15
    min_overlap: float = 0.7
    fma_mode: bool = True
    cand_cap: int = 64
    det_cap: int = 8192
    match_cap: int = 256

    @property
    def tail_w(self) -> int:
        # unsigned int tail_box_width = ((int)(double)(BB_BOTTOM_MOUSE.width) * tail_sub_bounding_box)
        # (LocoMouse_class.cpp:711): double product truncated
        return int(float(int(self.bb_w)) * self.tail_sub_bounding_box)

    def to_c(self) -> lm_config:
        return lm_config(
            self.vid_rows, self.vid_cols, self.n_rows, self.n_cols, self.bb_w, self.bb_h_bottom,
            self.bb_h_side, self.tail_w, int(self.flip), int(self.imadjust), self.conn, self.n_tail_points,
            float(self.min_overlap), int(self.fma_mode), self.cand_cap, self.det_cap, self.match_cap)


@dataclass
class Model:
    """Six detector templates + biases = LocoMouse_Model (LocoMouse_class.hpp:149-167).
    w[view][feature] float32 (rows, cols); rho[view][feature] float."""

    w: list  # [2][3] of np.ndarray float32 2-D
    rho: list  # [2][3] of float
    _keep: list = field(default_factory=list, repr=False)

    def to_c(self):
        arr = TemplateArray()
        self._keep = []
        for v in range(2):
            for k in range(3):
                a = np.ascontiguousarray(self.w[v][k], dtype=np.float32)
                if a.ndim != 2:
                    raise ValueError("template must be 2-D")
                self._keep.append(a)
                arr[v][k].w = a.ctypes.data_as(C.POINTER(C.c_float))
                arr[v][k].rows = a.shape[0]
                arr[v][k].cols = a.shape[1]
                arr[v][k].rho = float(self.rho[v][k])
        return arr


class Results:
    """Host result buffers for n frames (struct-of-arrays, see lm_results in the header)."""

    def __init__(self, n: int, cand_cap: int, match_cap: int, n_tail_points: int = 15, buffer: np.ndarray | None = None,
                 pinned: bool = False):
        """All arrays are views of ONE contiguous byte buffer (self.raw), so a whole result set moves with a single
        copy / collective.  `buffer` adopts an existing buffer of exactly raw_nbytes(...) bytes (e.g. a gathered one).
        pinned: allocate the buffer page-locked (needs a CUDA device), so lm_detect_batch copies results device -> here
        directly instead of through its staging buffers."""
        self.n = int(n)
        self.cand_cap = int(cand_cap)
        self.match_cap = int(match_cap)
        self.n_tail_points = int(n_tail_points)
        layout, total = self._layout(max(self.n, 1), self.cand_cap, self.match_cap, self.n_tail_points)
        self._pin = None
        if buffer is None and pinned:
            import torch

            self._pin = torch.zeros(total, dtype=torch.uint8, pin_memory=True)   # keeps the page-locked allocation alive
            buffer = self._pin.numpy()
        elif buffer is None:
            buffer = np.zeros(total, np.uint8)
        else:
            buffer = np.ascontiguousarray(buffer, dtype=np.uint8).reshape(-1)
            if buffer.size != total:
                raise ValueError(f"result buffer has {buffer.size} bytes, {total} expected")
        self.raw = buffer
        for name, (off, dtype, shape) in layout.items():
            nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
            setattr(self, name, buffer[off:off + nbytes].view(dtype).reshape(shape))

    @staticmethod
    def _layout(n, cand_cap, match_cap, n_tail_points):
        spec = (("n_bottom", np.int32, (n, 2)), ("n_side", np.int32, (n, 2)), ("bottom", CAND_DTYPE, (n, 2, cand_cap)),
                ("side", CAND_DTYPE, (n, 2, cand_cap)), ("match_n", np.int32, (n, 2, cand_cap)),
                ("match_y", np.int32, (n, 2, match_cap)), ("match_s", np.float64, (n, 2, match_cap)),
                ("tail", np.int32, (n, 3, n_tail_points)), ("flags", np.uint32, (n,)))
        layout, off = {}, 0
        for name, dtype, shape in spec:
            layout[name] = (off, dtype, shape)
            off += (int(np.prod(shape)) * np.dtype(dtype).itemsize + 15) & ~15
        return layout, off

    @staticmethod
    def raw_nbytes(n, cand_cap, match_cap, n_tail_points=15) -> int:
        return Results._layout(max(int(n), 1), cand_cap, match_cap, n_tail_points)[1]

    ARRAYS = ("n_bottom", "n_side", "bottom", "side", "match_n", "match_y", "match_s", "tail", "flags")

    def to_c(self) -> lm_results:
        r = lm_results()
        r.n_frames = self.n
        r.cand_cap = self.cand_cap
        r.match_cap = self.match_cap
        r.n_tail_points = self.n_tail_points
        for name in self.ARRAYS:
            setattr(r, name, getattr(self, name).ctypes.data)
        return r

    # ---- views in the reference's vocabulary -------------------------------------------------
    def candidates_bottom(self, f: int, feature: int):
        """CANDIDATES_BOTTOM_{PAW,SNOUT}[f] as a list of (x, y, score)."""
        k = int(self.n_bottom[f, feature])
        return [(int(c["x"]), int(c["y"]), float(c["s"])) for c in self.bottom[f, feature, :k]]

    def candidates_side(self, f: int, feature: int):
        k = int(self.n_side[f, feature])
        return [(int(c["x"]), int(c["y"]), float(c["s"])) for c in self.side[f, feature, :k]]

    def p22d(self, f: int, feature: int):
        """CANDIDATES_MATCHED_VIEWS_{PAW,SNOUT}[f]: list of ((x, yb, sb), [(yt, st), ...]); an empty
        side list stands for the reference's sentinel yt[0] = st[0] = -1."""
        out = []
        o = 0
        for i, cb in enumerate(self.candidates_bottom(f, feature)):
            m = int(self.match_n[f, feature, i])
            out.append((cb, [(int(self.match_y[f, feature, o + j]), float(self.match_s[f, feature, o + j]))
                             for j in range(m)]))
            o += m
        return out

    def checksum(self) -> int:
        """Order-sensitive 64-bit checksum of every result byte (used for size-independent parity
        properties: batch-split invariance, determinism)."""
        import zlib

        acc = 0
        for name in self.ARRAYS:
            a = getattr(self, name)[: self.n]
            acc = zlib.crc32(np.ascontiguousarray(a).view(np.uint8).reshape(-1), acc)
        return acc


def diff_results(a: Results, b: Results, score_rtol: float = 0.0, score_atol: float = 0.0) -> list:
    """Differences between two result sets.  Integer fields (counts, coordinates, pairings, tail,
    flags) must be identical; scores are compared bit-exactly when score_rtol == score_atol == 0, else
    within |x - y| <= score_atol + score_rtol * |y|  (score_atol carries the tolerance relative to the
    correlation magnitude: a detector score is a sum of O(rho) terms minus rho, so its error scales with
    |rho|, not with the possibly tiny score).  Returns human-readable mismatch strings (empty = equal)."""
    out = []
    if a.n != b.n:
        return [f"n {a.n} != {b.n}"]
    n = a.n
    for name in ("n_bottom", "n_side", "match_n", "match_y", "tail", "flags"):
        x, y = getattr(a, name)[:n], getattr(b, name)[:n]
        if not np.array_equal(x, y):
            idx = np.argwhere(x != y)[0]
            out.append(f"{name} differs first at {tuple(idx)}: {x[tuple(idx)]} vs {y[tuple(idx)]} "
                       f"({int((x != y).sum())} entries)")
    for name in ("bottom", "side"):
        x, y = getattr(a, name)[:n], getattr(b, name)[:n]
        for fld in ("x", "y"):
            if not np.array_equal(x[fld], y[fld]):
                idx = tuple(np.argwhere(x[fld] != y[fld])[0])
                out.append(f"{name}.{fld} differs first at {idx}: {x[fld][idx]} vs {y[fld][idx]}")
        if not _scores_equal(x["s"], y["s"], score_rtol, score_atol):
            out.append(f"{name}.s differs (max rel {_max_rel(x['s'], y['s']):.3e})")
    if not _scores_equal(a.match_s[:n], b.match_s[:n], score_rtol, score_atol):
        out.append(f"match_s differs (max rel {_max_rel(a.match_s[:n], b.match_s[:n]):.3e})")
    return out


def _scores_equal(x, y, rtol, atol=0.0):
    if rtol == 0.0 and atol == 0.0:
        return np.array_equal(x.view(np.uint64), y.view(np.uint64)) or np.array_equal(x, y, equal_nan=True)
    return bool(np.allclose(x, y, rtol=rtol, atol=atol, equal_nan=True))


def _max_rel(x, y):
    with np.errstate(divide="ignore", invalid="ignore"):
        d = np.abs(x - y) / np.maximum(np.abs(y), 1e-300)
    d = d[np.isfinite(d)]
    return float(d.max()) if d.size else 0.0
