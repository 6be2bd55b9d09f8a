"""Coverage-guided parity for the pairing stage (matchViews + checkVelCriterion, LocoMouse_class.cpp:999-1267): the
standard synthetic clips never reject a match on the velocity criterion and never hit the all-true boolD quirk
(SURVEY Q7/Q8), so these inputs are chosen -- and checked with the oracle's branch counters -- to exercise them.
Needs a B200: pytest -m gpu."""
import numpy as np
import pytest

from locomouse_cpp_b200 import synth
from locomouse_cpp_b200.types import Model, diff_results

pytestmark = pytest.mark.gpu


def _clip(total, start, n, rho_scale):
    spec = synth.SynthSpec()
    cfg, model, bkg, calib, _, _, _, _ = synth.make_problem(spec, 8, seed=1000)
    frames, bx, bs, bb = synth.make_video(spec, n, 1000, "cpu", bkg, start_frame=start, total=total)
    prev = synth.make_video(spec, 1, 1000, "cpu", bkg, start_frame=start - 1, total=total)[0][0].numpy()
    model = Model(w=model.w, rho=[[r * rho_scale for r in row] for row in model.rho])
    return cfg, model, bkg, calib, frames.numpy(), bx, bs, bb, prev


@pytest.mark.parametrize("total,rho_scale,need", [
    (600, 1.0, ("velocity_rejections", "velocity_accepts", "moving_windows")),     # mid-clip, ~2.7 px / frame
    (300, 1.0, ("velocity_rejections", "velocity_accepts")),
    (600, 1.45, ("all_equal_boolD_zeroed",)),                                       # sparse candidates: 1 x 1 pairings
])
def test_pairing_branches_are_exercised_and_bit_exact(oracle, total, rho_scale, need):
    from locomouse_cpp_b200.api import Detector

    cfg, model, bkg, calib, frames, bx, bs, bb, prev = _clip(total, total // 3, 40, rho_scale)
    oracle.coverage(reset=True)
    ref = oracle.detect(cfg, model, bkg, calib, frames, bx, bs, bb, prev_frame=prev, first_frame_index=total // 3, n_threads=8)
    cov = oracle.coverage()
    for k in need:
        assert cov[k] > 0, f"input does not exercise {k}: {cov}"
    for screen in (2, 0):
        det = Detector(cfg, model, bkg, calib, device=0)
        det.set_option("screen", screen)
        det.set_option("subbatch", 16)      # the previous-frame halo crosses sub-batch borders too
        got = det.detect_batch(frames, bx, bs, bb, prev_frame=prev, first_frame_index=total // 3)
        det.close()
        assert diff_results(got, ref) == []
