import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure; built on demand from oracle/)."""
    from oracle import oracle as o

    o.lib()
    return o


@pytest.fixture(scope="session")
def small_problem():
    """A few CPU-rendered frames at the config-1 geometry with calibrated rho (shared by tests)."""
    from locomouse_cpp_b200 import synth

    spec = synth.SynthSpec()
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, 6, seed=1000)
    return dict(spec=spec, cfg=cfg, model=model, bkg=bkg, calib=calib, frames=frames.numpy(), bb_x=bx,
                bb_y_side=bs, bb_y_bottom=bb)
