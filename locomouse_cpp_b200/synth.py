"""Synthetic LocoMouse inputs in the reference's frame geometry.

The reference ships no video, model, background or calibration file, so every
benchmark / parity input is synthesised here: 8-bit grayscale raw frames with the side (mirror) view
in the upper rows and the bottom view in the lower rows, a smooth background, a mouse model (body,
four paws, snout, tail) walking left to right, random smooth detector templates of the reference's
shapes and per-frame bounding-box corners as pass 1 of the reference would hand them to the hot loop
(BB_X_POS / BB_Y_SIDE_POS / BB_Y_BOTTOM_POS, LocoMouse_TM.cpp:139-155, LocoMouse_TM_DE.cpp:32-51).

Frames are rendered with torch so the same code fills host arrays for the parity tests and HBM for
the benchmark.  torch is plumbing here: nothing in this file is on the measured path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

from .types import BOTTOM, PAW, SIDE, SNOUT, TAIL, Config, Model


@dataclass
class SynthSpec:
    """Geometry of one synthetic setup.  scale=1 is SURVEY config 1-4, scale=2 is config 5."""

    method: str = "TM"          # "TM" | "TM_DE" | "base"  (LocoMouse_Methods.cpp:9-22)
    scale: int = 1
    n_rows: int = 400           # calibrated image
    n_cols: int = 1700
    side_h: int = 165           # BB_SIDE_VIEW.height ; bottom view = the remaining rows
    bb_w: int = 400
    bb_h_side_tm: int = 150     # LocoMouse_TM.hpp:32
    tsize: int = 30             # all six templates tsize x tsize (assumed; no model file is shipped)
    vid_pad: int = 0            # raw video larger than the calibrated image by this many px per side
    flip: bool = False
    warp: bool = False          # seeded smooth calibration warp instead of identity
    fma_mode: bool = True
    conn: int = 8
    cand_cap: int = 64
    det_cap: int = 8192
    match_cap: int = 256
    mouse_scale: float = 1.0    # size of the rendered mouse relative to the config-1 animal
    tshapes: tuple | None = None  # optional ((rows, cols),)*3 per view: [[paw, snout, tail] bottom, [..] side]

    def scaled(self) -> "SynthSpec":
        s = self.scale
        if s == 1:
            return self
        import dataclasses

        return dataclasses.replace(self, scale=1, n_rows=self.n_rows * s, n_cols=self.n_cols * s,
                                   side_h=self.side_h * s, bb_w=self.bb_w * s, bb_h_side_tm=self.bb_h_side_tm * s,
                                   tsize=self.tsize * s, mouse_scale=self.mouse_scale * s)

    def template_shape(self, view: int, feat: int):
        if self.tshapes is not None:
            return tuple(self.tshapes[view][feat])
        t = self.scaled().tsize
        return (t, t)

    def max_tsize(self) -> int:
        return max(max(self.template_shape(v, k)) for v in range(2) for k in range(3))

    def config(self) -> Config:
        s = self.scaled()
        bottom_h = s.n_rows - s.side_h
        bb_h_side = s.bb_h_side_tm if s.method == "TM" else s.side_h
        return Config(vid_rows=s.n_rows + 2 * s.vid_pad, vid_cols=s.n_cols + 2 * s.vid_pad, n_rows=s.n_rows,
                      n_cols=s.n_cols, bb_w=s.bb_w, bb_h_bottom=bottom_h, bb_h_side=bb_h_side,
                      tail_sub_bounding_box=0.6, flip=s.flip, imadjust=(s.method != "base"), conn=s.conn,
                      n_tail_points=15, min_overlap=0.7, fma_mode=s.fma_mode, cand_cap=s.cand_cap,
                      det_cap=s.det_cap, match_cap=s.match_cap)


# ------------------------------------------------------------------------------------------------
# static inputs: background, calibration map, templates
# ------------------------------------------------------------------------------------------------
def make_background(spec: SynthSpec, seed: int = 1000) -> np.ndarray:
    """Smooth illumination gradient 20..60 plus N(0,2) pixel noise, raw-video sized u8."""
    cfg = spec.config()
    rng = np.random.Generator(np.random.PCG64(seed))
    y = np.linspace(0, 1, cfg.vid_rows)[:, None]
    x = np.linspace(0, 1, cfg.vid_cols)[None, :]
    g = 20 + 25 * x + 15 * y * (1 - x)
    g = g + rng.normal(0, 2, g.shape)
    return np.clip(np.rint(g), 0, 255).astype(np.uint8)


def make_calibration(spec: SynthSpec, seed: int = 2000) -> np.ndarray:
    """ind_warp_mapping (LocoMouse_class.cpp:428): int32 [n_rows][n_cols] of flat raw-frame indices.
    Identity (offset by vid_pad) or a seeded smooth warp of a few pixels."""
    cfg = spec.config()
    s = spec.scaled()
    r = np.arange(cfg.n_rows)[:, None] + s.vid_pad
    c = np.arange(cfg.n_cols)[None, :] + s.vid_pad
    rr = np.broadcast_to(r, (cfg.n_rows, cfg.n_cols)).astype(np.float64)
    cc = np.broadcast_to(c, (cfg.n_rows, cfg.n_cols)).astype(np.float64)
    if s.warp:
        rng = np.random.Generator(np.random.PCG64(seed))
        ph = rng.uniform(0, 2 * np.pi, 4)
        rr = rr + 2.0 * np.sin(cc / 190.0 + ph[0]) + 1.0 * np.sin(rr / 70.0 + ph[1])
        cc = cc + 3.0 * np.sin(rr / 110.0 + ph[2]) + 1.5 * np.sin(cc / 230.0 + ph[3])
    ri = np.clip(np.rint(rr), 0, cfg.vid_rows - 1).astype(np.int64)
    ci = np.clip(np.rint(cc), 0, cfg.vid_cols - 1).astype(np.int64)
    return (ri * cfg.vid_cols + ci).astype(np.int32)


def _smooth_template(rng, rows, cols, sigma):
    f = rng.normal(0, 1, (rows + 8 * int(sigma), cols + 8 * int(sigma)))
    k = np.arange(-4 * int(sigma), 4 * int(sigma) + 1)
    g = np.exp(-0.5 * (k / sigma) ** 2)
    g /= g.sum()
    f = np.apply_along_axis(lambda v: np.convolve(v, g, mode="valid"), 0, f)
    f = np.apply_along_axis(lambda v: np.convolve(v, g, mode="valid"), 1, f)
    f = f[:rows, :cols]
    hann = np.outer(np.hanning(rows + 2)[1:-1], np.hanning(cols + 2)[1:-1])
    f = f * hann
    f = f - f.mean()
    f = f / (np.sqrt((f ** 2).sum()) * 64.0)
    return f.astype(np.float32)


def make_model(spec: SynthSpec, seed: int = 7, rho=None) -> Model:
    """Six random-init smooth templates (Gaussian random field x Hann window, mean removed).  rho is
    filled by calibrate_rho() unless given."""
    s = spec.scaled()
    rng = np.random.Generator(np.random.PCG64(seed))
    w = [[None] * 3 for _ in range(2)]
    for v in (BOTTOM, SIDE):
        for k in (PAW, SNOUT, TAIL):
            rows, cols = spec.template_shape(v, k)
            sigma = 2.0 * min(rows, cols) / 30.0 if k != TAIL else 1.5 * min(rows, cols) / 30.0
            w[v][k] = _smooth_template(rng, rows, cols, max(1.0, sigma))
    if rho is None:
        rho = [[0.0] * 3 for _ in range(2)]
    return Model(w=w, rho=[list(map(float, r)) for r in rho])


# ------------------------------------------------------------------------------------------------
# video
# ------------------------------------------------------------------------------------------------
def _trajectory(spec: SynthSpec, n: int, seed: int, start_frame: int = 0, total: int | None = None):
    """Nose x position per frame and BB corners, for frames [start_frame, start_frame+n) of a video of
    `total` frames (default n): left to right across the corridor with treadmill-like jitter."""
    s = spec.scaled()
    total = total or (start_frame + n)
    rng = np.random.Generator(np.random.PCG64(seed + 17))
    jitter = np.cumsum(rng.normal(0, 0.6 * s.mouse_scale, total))
    jitter -= np.linspace(0, jitter[-1], total)
    t = np.arange(total) / max(total - 1, 1)
    x_lo, x_hi = 0.09 * s.n_cols, s.n_cols - 1 - 12 * s.mouse_scale
    nose = x_lo + (x_hi - x_lo) * t + jitter
    raw = nose + 10.0 * s.mouse_scale
    # moving average, window 5, as vecmovingaverage (LocoMouse_class.cpp:1559-1608)
    bbx = raw.copy()
    if total > 5:
        cs = np.convolve(raw, np.ones(5), mode="valid") / 5.0
        bbx[2:total - 2] = np.floor(cs)
    bbx = np.clip(np.floor(bbx), spec.max_tsize(), s.n_cols - 1).astype(np.uint32)
    sl = slice(start_frame, start_frame + n)
    return nose[sl], bbx[sl]


def make_video(spec: SynthSpec, n: int, seed: int = 1000, device="cpu", bkg: np.ndarray | None = None,
               start_frame: int = 0, total: int | None = None, chunk: int = 64, out: torch.Tensor | None = None):
    """Returns (frames u8 torch [n, vid_rows, vid_cols] on `device`, bb_x, bb_y_side, bb_y_bottom as
    uint32 numpy arrays).  Content: background + mouse model + N(0,3) noise, saturated to u8."""
    cfg = spec.config()
    s = spec.scaled()
    sc = float(s.mouse_scale)
    if bkg is None:
        bkg = make_background(spec, seed)
    dev = torch.device(device)
    H, W = cfg.vid_rows, cfg.vid_cols
    nose, bbx = _trajectory(spec, n, seed, start_frame, total)
    bb_y_bottom = np.full(n, cfg.n_rows - 1, np.uint32)
    bb_y_side = np.full(n, s.side_h - 1, np.uint32)  # 165-1 (LocoMouse_TM.cpp:143), BB_SIDE_VIEW.height-1 (TM_DE)
    frames = out if out is not None else torch.empty((n, H, W), dtype=torch.uint8, device=dev)
    bkg_t = torch.from_numpy(bkg.astype(np.float32)).to(dev)
    X = torch.arange(W, device=dev, dtype=torch.float32).view(1, 1, W) - s.vid_pad
    Y = torch.arange(H, device=dev, dtype=torch.float32).view(1, H, 1) - s.vid_pad
    gen = torch.Generator(device=dev)
    rngp = np.random.Generator(np.random.PCG64(seed + 99))
    tex = [(rngp.uniform(6, 40) * sc, rngp.uniform(0, 2 * np.pi), rngp.uniform(0, 2 * np.pi), rngp.uniform(6, 14))
           for _ in range(10)]
    side_h = float(s.side_h)
    bot_h = float(cfg.n_rows - s.side_h)
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        B = c1 - c0
        fidx = torch.arange(start_frame + c0, start_frame + c1, device=dev, dtype=torch.float32).view(B, 1, 1)
        nx = torch.from_numpy(nose[c0:c1].astype(np.float32)).to(dev).view(B, 1, 1)
        cx = nx - 150.0 * sc
        img = torch.zeros((B, H, W), dtype=torch.float32, device=dev)
        for view in (BOTTOM, SIDE):
            if view == BOTTOM:
                cy = side_h + 0.5 * bot_h
                ay, body = 42.0 * sc, 62.0
            else:
                cy = side_h - 58.0 * sc
                ay, body = 27.0 * sc, 55.0
            ax = 140.0 * sc
            u, v = X - cx, Y - cy
            inside = torch.sigmoid((1.0 - (u / ax) ** 2 - (v / ay) ** 2) * 6.0)
            t_acc = torch.zeros_like(img)
            for (lam, pu, pv, amp) in tex:
                t_acc = t_acc + amp * torch.sin(u * (2 * math.pi / lam) * math.cos(pu) +
                                                v * (2 * math.pi / lam) * math.sin(pu) + pv)
            layer = inside * (body + t_acc * 0.45)
            # paws: two front, two hind, gait-phased
            gait = fidx * 0.23
            for k, (ox, oy, ph) in enumerate(((95, 36, 0.0), (70, -36, math.pi), (-85, 38, math.pi), (-110, -38, 0.0))):
                px = cx + (ox + 22.0 * torch.sin(gait + ph)) * sc
                if view == BOTTOM:
                    py = cy + oy * sc
                else:
                    py = side_h - (14.0 + 5.0 * (k % 2)) * sc - 6.0 * sc * torch.clamp(torch.sin(gait + ph), min=0)
                layer = layer + 150.0 * torch.exp(-((X - px) ** 2 + (Y - py) ** 2) / (2 * (5.0 * sc) ** 2))
            # snout
            sx = nx - 8.0 * sc
            sy = cy if view == BOTTOM else side_h - 52.0 * sc
            layer = layer + 140.0 * torch.exp(-((X - sx) ** 2 + (Y - sy) ** 2) / (2 * (6.0 * sc) ** 2))
            # tail: thin wavy line behind the body
            tl0, tl1 = cx - 255.0 * sc, cx - 125.0 * sc
            ty = (cy if view == BOTTOM else side_h - 70.0 * sc) + (9.0 * sc) * torch.sin((X - cx) / (34.0 * sc) + fidx * 0.11)
            if view == SIDE:
                ty = ty - (cx - X).clamp(min=0) * 0.10
            win = torch.sigmoid((X - tl0) / (3.0 * sc)) * torch.sigmoid((tl1 - X) / (3.0 * sc))
            layer = layer + 95.0 * win * torch.exp(-((Y - ty) ** 2) / (2 * (2.2 * sc) ** 2))
            # confine each view to its rows
            if view == BOTTOM:
                layer = layer * (Y >= side_h).float()
            else:
                layer = layer * (Y < side_h).float()
            img = img + layer
        if s.flip:  # the raw video shows the mouse mirrored; readFrame flips it back (class.cpp:1323)
            img = img.flip(-1)
        gen.manual_seed(int(seed) * 1000003 + start_frame + c0)
        noise = torch.randn((B, H, W), generator=gen, device=dev, dtype=torch.float32) * 3.0
        img = img + bkg_t.view(1, H, W) + noise
        frames[c0:c1] = img.round_().clamp_(0, 255).to(torch.uint8)
    return frames, bbx, bb_y_side, bb_y_bottom


# ------------------------------------------------------------------------------------------------
# rho calibration (torch conv2d in float64 on CPU: only chooses numbers that then become config)
# ------------------------------------------------------------------------------------------------
def preprocess_reference_torch(cfg: Config, bkg: np.ndarray, calib: np.ndarray, frame: np.ndarray) -> np.ndarray:
    """Plain numpy restatement of readFrame + imadjust used ONLY to pick rho; not a parity oracle."""
    d = np.clip(frame.astype(np.int32) - bkg.astype(np.int32), 0, 255)
    lo, hi = int(d.min()), int(d.max())
    scale = 255.0 / (hi - lo) if hi > lo else 0.0
    nrm = np.clip(np.rint(d * scale - lo * scale), 0, 255).astype(np.uint8)
    I = nrm.reshape(-1)[calib.astype(np.int64)]
    if cfg.flip:
        I = I[:, ::-1]
    if cfg.imadjust:
        i = np.arange(256, dtype=np.float64)
        lut = np.where(i >= 153.0, 255.0, i * (255.0 / 153.0))
        lut = np.floor(lut + 0.5).astype(np.uint8)
        I = lut[I]
    return np.ascontiguousarray(I)


def calibrate_rho(spec: SynthSpec, model: Model, bkg, calib, frames: np.ndarray, bb_x, bb_y_side, bb_y_bottom,
                  target_frac=(0.012, 0.012, 0.02)) -> Model:
    """Choose rho per template so that `target_frac` of the unmasked crop pixels score > 0 on the given
    frames (paw, snout, tail).  This mimics an SVM margin and bounds the O(N^2) NMS of the CPU
    reference (SURVEY.md §7 hard part 5).  Deterministic for given inputs."""
    cfg = spec.config()
    rho = [[0.0] * 3 for _ in range(2)]
    crops = {BOTTOM: [], SIDE: []}
    for f in range(frames.shape[0]):
        I = preprocess_reference_torch(cfg, bkg, calib, frames[f]).astype(np.float64)
        x0 = int(bb_x[f]) - cfg.bb_w + 1
        for v, (y_pos, h) in ((BOTTOM, (bb_y_bottom[f], cfg.bb_h_bottom)), (SIDE, (bb_y_side[f], cfg.bb_h_side))):
            y0 = int(y_pos) - h + 1
            pad = spec.max_tsize()
            P = np.zeros((h + 2 * pad, cfg.bb_w + 2 * pad))
            ys, xs = np.arange(y0 - pad, y0 + h + pad), np.arange(x0 - pad, x0 + cfg.bb_w + pad)
            yv = (ys >= 0) & (ys < cfg.n_rows)
            xv = (xs >= 0) & (xs < cfg.n_cols)
            P[np.ix_(yv, xv)] = I[np.ix_(ys[yv], xs[xv])]
            crops[v].append(P)
    from scipy.signal import fftconvolve

    for v in (BOTTOM, SIDE):
        P = np.stack(crops[v])  # [n, h + 2 pad, w + 2 pad] float64
        h = cfg.bb_h_bottom if v == BOTTOM else cfg.bb_h_side
        pad = spec.max_tsize()
        centre = P[:, pad:pad + h, pad:pad + cfg.bb_w]
        for k in (PAW, SNOUT, TAIL):
            w = model.w[v][k].astype(np.float64)
            kh, kw = w.shape
            # valid cross-correlation via FFT (flip the kernel); only used to pick rho
            sc = np.stack([fftconvolve(P[i], w[::-1, ::-1], mode="valid") for i in range(P.shape[0])])
            oy, ox = pad - kh // 2, pad - kw // 2
            sc = sc[:, oy:oy + h, ox:ox + cfg.bb_w]
            if k == TAIL:
                vals = sc[:, :, :cfg.tail_w].reshape(-1)
            else:
                vals = sc[centre > 25]
            if vals.size == 0:
                rho[v][k] = 0.0
                continue
            q = np.quantile(vals, 1.0 - target_frac[k])
            rho[v][k] = float(np.float32(q))
    return Model(w=model.w, rho=rho)


def make_problem(spec: SynthSpec, n_frames: int, seed: int = 1000, device="cpu", calib_frames: int = 3, target_frac=None):
    """Everything one detect call needs: (cfg, model, bkg, calib, frames, bb_x, bb_y_side, bb_y_bottom).
    rho is calibrated on the first `calib_frames` CPU-rendered frames of video `seed` evenly spread over
    the sequence, so it is identical for host- and device-rendered videos."""
    cfg = spec.config()
    bkg = make_background(spec, seed)
    calib = make_calibration(spec, seed + 1000)
    total = max(n_frames, 8)
    pick = np.unique(np.linspace(0, total - 1, calib_frames).astype(int))
    cf, cbx, cbs, cbb = [], [], [], []
    for p in pick:
        fr, bx, bs, bb = make_video(spec, 1, seed, "cpu", bkg, start_frame=int(p), total=total)
        cf.append(fr[0].numpy())
        cbx.append(bx[0])
        cbs.append(bs[0])
        cbb.append(bb[0])
    model = (calibrate_rho(spec, make_model(spec), bkg, calib, np.stack(cf), cbx, cbs, cbb) if target_frac is None else
             calibrate_rho(spec, make_model(spec), bkg, calib, np.stack(cf), cbx, cbs, cbb, target_frac=tuple(target_frac)))
    frames, bb_x, bb_y_side, bb_y_bottom = make_video(spec, n_frames, seed, device, bkg, total=total)
    return cfg, model, bkg, calib, frames, bb_x, bb_y_side, bb_y_bottom
