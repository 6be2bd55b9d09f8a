"""Pass 1 of LocoMouse_TM on the device: lm_bounding_box_tm against the oracle (pinned stage by stage to the
reference's own compiled computeMouseBox_DD / bwAreaOpen / imfill lines, tests/test_oracle_pass1_tm.py) -- in the reference's
mode (integer sums read as floats) and with integer sums, both connectivities, a mirrored / warped calibration, several disk
sizes and area limits, noise frames (thousands of components), host and device frames, and the global-memory run arrays."""
import numpy as np
import pytest

from locomouse_cpp_b200 import synth
from locomouse_cpp_b200.types import bb_tm_params

pytestmark = pytest.mark.gpu


def disk(r):
    y, x = np.mgrid[-r:r + 1, -r:r + 1]
    h = ((x * x + y * y) <= (r + 0.5) ** 2).astype(np.float64)
    return (h / h.sum()).astype(np.float32)


def _problem(n, seed=1000, **kw):
    spec = synth.SynthSpec(method="TM", **kw)
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, n, seed=seed)
    return spec, cfg, model, bkg, calib, frames.numpy()


def _detector(cfg, model, bkg, calib):
    from locomouse_cpp_b200.api import Detector

    return Detector(cfg, model, bkg, calib, device=0)


@pytest.mark.parametrize("kw,r,pk", [
    (dict(), 5, dict()),                                                     # the reference: every limit -1
    (dict(), 5, dict(sums_as_float=0)),
    (dict(), 3, dict(sums_as_float=1, min_pixel_visible=0)),                 # the reference with threshold 0: everything qualifies
    (dict(flip=True, warp=True, vid_pad=5), 4, dict(sums_as_float=0, min_pixel_count=40, side_threshold=8, zero_col_pre=30, zero_col_post=1650,
                                                     zero_row_pre=10, zero_row_post=150)),
    (dict(conn=4), 2, dict(sums_as_float=0, min_pixel_count=3, min_pixel_visible=600)),
    (dict(), 7, dict(sums_as_float=0, min_pixel_count=200)),                 # 15 x 15 kernel
])
def test_device_equals_oracle(oracle, kw, r, pk):
    spec, cfg, model, bkg, calib, frames = _problem(5, **kw)
    det = _detector(cfg, model, bkg, calib)
    P = bb_tm_params(cfg, disk(r), side_h=spec.side_h, **pk)
    want, wl = oracle.bounding_box_tm(cfg, bkg, calib, frames, P)
    got, gl = det.bounding_box_tm(frames, P)
    assert np.array_equal(gl, wl)
    assert np.array_equal(got, want)
    if not pk.get("sums_as_float", 1):
        assert (wl[:, 1] > 0).any()


def test_noise_device_frames_and_global_run_arrays(oracle, monkeypatch):
    """Frames of noise (tens of thousands of runs: beyond the shared-memory capacity), an image that is foreground at pixel
    (0, 0) (the flood fill starts on the object), device-resident frames, and everything again with the run capacity forced
    down so that every frame takes the global-memory instance."""
    import torch

    spec, cfg, model, bkg, calib, frames = _problem(3)
    rng = np.random.Generator(np.random.PCG64(9))
    noise = rng.integers(0, 256, frames.shape, dtype=np.uint8)
    bright = np.clip(frames[:1].astype(np.int32) + 0, 0, 255).astype(np.uint8)
    bright[0, :60, :300] = 255
    frames = np.concatenate([frames, noise[:2], bright])
    for pk in (dict(sums_as_float=0, min_pixel_count=4), dict(sums_as_float=0, min_pixel_count=1, side_threshold=0, min_pixel_visible=3000)):
        P = bb_tm_params(cfg, disk(2), side_h=spec.side_h, **pk)
        want, wl = oracle.bounding_box_tm(cfg, bkg, calib, frames, P)
        det = _detector(cfg, model, bkg, calib)
        got, gl = det.bounding_box_tm(frames, P)
        assert np.array_equal(gl, wl) and np.array_equal(got, want)
        got_d, gl_d = det.bounding_box_tm(torch.from_numpy(frames).cuda(), P)
        assert np.array_equal(gl_d, wl)
        monkeypatch.setenv("LM_BBOX_RUNCAP", "16")
        got_s, gl_s = det.bounding_box_tm(frames, P)
        assert np.array_equal(gl_s, wl) and np.array_equal(got_s, want)
        monkeypatch.delenv("LM_BBOX_RUNCAP")


def test_argument_checks(oracle):
    spec, cfg, model, bkg, calib, frames = _problem(2)
    det = _detector(cfg, model, bkg, calib)
    with pytest.raises(ValueError):   # colRange(ZERO_COL_POST, N_COLS) on a narrower side view throws in the reference
        det.bounding_box_tm(frames, bb_tm_params(cfg, disk(2), side_w=cfg.n_cols - 10))
    with pytest.raises(ValueError):
        det.bounding_box_tm(frames, bb_tm_params(cfg, disk(2), min_pixel_count=0))
    with pytest.raises(ValueError):   # not a smoothing kernel: filtered values above 1
        det.bounding_box_tm(frames, bb_tm_params(cfg, np.ones((3, 3), np.float32)))
    with pytest.raises(ValueError):
        det.bounding_box_tm(frames, bb_tm_params(cfg, disk(2), zero_row_post=spec.side_h + 1))
