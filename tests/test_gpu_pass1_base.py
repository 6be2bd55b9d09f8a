"""Pass 1 of the base class on the device: lm_bounding_box_base against the oracle (which is pinned to the
reference's own compiled computeMouseBox lines, tests/test_oracle_pass1_base.py) -- in the reference's mode (integer sums
read as floats) and with integer sums, for both connectivities, median sizes, a mirrored / warped calibration, views that do
not start at a word boundary, host and device frames, and the global-memory labelling path."""
import copy

import numpy as np
import pytest

from locomouse_cpp_b200 import synth
from locomouse_cpp_b200.types import bb_base_params

pytestmark = pytest.mark.gpu


def _problem(n, seed=1000, **kw):
    spec = synth.SynthSpec(method="base", **kw)
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, n, seed=seed)
    return spec, cfg, model, bkg, calib, frames.numpy()


def _detector(cfg, model, bkg, calib):
    from locomouse_cpp_b200.api import Detector

    return Detector(cfg, model, bkg, calib, device=0)


@pytest.mark.parametrize("kw,pk", [
    (dict(), dict()),
    (dict(), dict(sums_as_float=0)),
    (dict(flip=True, warp=True, vid_pad=5), dict(sums_as_float=0, median_filter_size=5)),
    (dict(conn=4), dict(sums_as_float=0, min_pixel_visible=400, median_filter_size=3)),
    (dict(), dict(sums_as_float=1, min_pixel_visible=0)),
])
def test_device_equals_oracle(oracle, kw, pk):
    spec, cfg, model, bkg, calib, frames = _problem(5, **kw)
    det = _detector(cfg, model, bkg, calib)
    P = bb_base_params(cfg, side_h=spec.side_h, **pk)
    want, wl = oracle.bounding_box_base(cfg, bkg, calib, frames, P)
    got, gl = det.bounding_box_base(frames, P)
    assert np.array_equal(gl, wl)
    assert np.array_equal(got, want)
    if not pk.get("sums_as_float", 1):
        assert (wl[:, :, 1] > 0).any()


def test_unaligned_views_noise_and_slow_path(oracle, monkeypatch):
    """Views inset from the image border (bit rows re-aligned), frames of pure noise with a 3 x 3 median (thousands of runs),
    and the same input forced through the global-memory labelling: all equal the oracle."""
    import torch

    spec, cfg, model, bkg, calib, frames = _problem(3)
    rng = np.random.Generator(np.random.PCG64(9))
    noise = rng.integers(0, 256, frames.shape, dtype=np.uint8)
    frames = np.concatenate([frames, noise[:2]])
    P = bb_base_params(cfg, side_h=spec.side_h, sums_as_float=0, median_filter_size=3)
    P.side_x, P.side_y, P.side_w, P.side_h = 37, 3, cfg.n_cols - 91, spec.side_h - 9
    P.bottom_x, P.bottom_w = 5, cfg.n_cols - 5
    want, wl = oracle.bounding_box_base(cfg, bkg, calib, frames, P)
    det = _detector(cfg, model, bkg, calib)
    got, gl = det.bounding_box_base(frames, P)
    assert np.array_equal(gl, wl) and np.array_equal(got, want)
    got_d, gl_d = det.bounding_box_base(torch.from_numpy(frames).cuda(), P)
    assert np.array_equal(gl_d, wl)
    monkeypatch.setenv("LM_BBOX_RUNCAP", "16")
    got_s, gl_s = det.bounding_box_base(frames, P)
    assert np.array_equal(gl_s, wl) and np.array_equal(got_s, want)


def test_box_size_and_argument_checks(oracle):
    spec, cfg, model, bkg, calib, frames = _problem(2)
    det = _detector(cfg, model, bkg, calib)
    rng = np.random.Generator(np.random.PCG64(1))
    for n in (1, 2, 7, 30):
        w, hb, hs = (np.floor(rng.uniform(0, 500, n)) for _ in range(3))
        assert det.mouse_box_size(w, hb, hs) == oracle.mouse_box_size(w, hb, hs)
    with pytest.raises(ValueError):
        det.bounding_box_base(frames, bb_base_params(cfg, median_filter_size=4))
    with pytest.raises(ValueError):
        det.bounding_box_base(frames, bb_base_params(cfg, side_h=cfg.n_rows + 1))
