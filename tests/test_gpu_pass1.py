"""Pass 1 of LocoMouse_TM_DE on the device (lm_bounding_box_tm_de, csrc/k_bbox.cu) against the oracle: per-frame raw box
positions, first/last columns and the smoothed BB_X_POS must be identical.  Needs a B200: pytest -m gpu."""
import numpy as np
import pytest

from locomouse_cpp_b200 import synth
from locomouse_cpp_b200.types import bb_de_params

pytestmark = pytest.mark.gpu


def _det(cfg, model, bkg, calib):
    from locomouse_cpp_b200.api import Detector

    return Detector(cfg, model, bkg, calib, device=0)


@pytest.mark.parametrize("kw", [dict(), dict(flip=True), dict(warp=True, vid_pad=4), dict(flip=True, warp=True, vid_pad=2)])
def test_pass1_matches_oracle(oracle, kw):
    spec = synth.SynthSpec(method="TM_DE", **kw)
    n = 12
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, n, seed=1000)
    frames = frames.numpy()
    p = bb_de_params(cfg, side_h=spec.side_h)
    want, want_raw, want_lims = oracle.bounding_box_tm_de(cfg, bkg, calib, frames, p)
    det = _det(cfg, model, bkg, calib)
    got, raw, lims = det.bounding_box_tm_de(frames, p)
    assert np.array_equal(lims, want_lims)
    assert np.array_equal(raw.view(np.uint64), want_raw.view(np.uint64))
    assert np.array_equal(got, want)
    assert (want_lims[:, 1] > 0).any()


def test_pass1_device_frames_many_and_parameters(oracle):
    """600 device-resident frames (three internal chunks) and non-default bands / thresholds on a subset."""
    import torch

    spec = synth.SynthSpec(method="TM_DE")
    cfg, model, bkg, calib, _, _, _, _ = synth.make_problem(spec, 4, seed=1000)
    frames, bx, bs, bb = synth.make_video(spec, 600, 77, "cuda", bkg)
    torch.cuda.synchronize()
    det = _det(cfg, model, bkg, calib)
    p = bb_de_params(cfg, side_h=spec.side_h)
    got, raw, lims = det.bounding_box_tm_de(frames, p)
    host = frames[:40].cpu().numpy()
    want, want_raw, want_lims = oracle.bounding_box_tm_de(cfg, bkg, calib, host, p)
    assert np.array_equal(lims[:40], want_lims) and np.array_equal(raw[:40], want_raw)
    # the moving average only looks 2 frames ahead: the first 38 smoothed values of the long run equal the short run's
    assert np.array_equal(got[:38], want[:38])
    assert np.unique(lims[:, 1]).size > 20          # the box follows the mouse
    for kwp in (dict(threshold=40.0, min_count=3), dict(zero_col_pre=0, zero_col_post=cfg.n_cols, zero_row_pre=0,
                                                         zero_row_post=spec.side_h, width_margin=1.0), dict(min_count=0),
                dict(threshold=300.0)):
        q = bb_de_params(cfg, side_h=spec.side_h, **kwp)
        a = det.bounding_box_tm_de(host, q)
        b = oracle.bounding_box_tm_de(cfg, bkg, calib, host, q)
        for x, y in zip(a, b):
            assert np.array_equal(x, y), kwp


def test_pass1_degenerate_frames_and_errors(oracle):
    spec = synth.SynthSpec(method="TM_DE")
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, 3, seed=1000)
    frames = frames.numpy().copy()
    frames[0] = bkg                    # nothing but background: constant zero difference, imin == imax
    frames[1] = 255                    # saturated
    frames[2, :, :] = bkg
    frames[2, 120:140, 500] = 255      # a single bright column: exactly one qualifying column -> lims = (x, 0)
    p = bb_de_params(cfg, side_h=spec.side_h)
    det = _det(cfg, model, bkg, calib)
    a = det.bounding_box_tm_de(frames, p)
    b = oracle.bounding_box_tm_de(cfg, bkg, calib, frames, p)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert b[2][2].tolist() == [500, 0]
    with pytest.raises(ValueError):
        det.bounding_box_tm_de(frames, bb_de_params(cfg, side_h=cfg.n_rows + 1))
