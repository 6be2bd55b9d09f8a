"""The NMS kernels (k_nms_bottom = nmsMax, k_nms_side = peakClustering) checked DIRECTLY against the reference's own
code: the golden vectors in tests/golden/reference_nms.npz were produced by LocoMouse_class.cpp:1610-1905 compiled from
/root/reference (oracle/Makefile `ref`, tests/golden/make_reference_golden.py).  Needs a B200: pytest -m gpu."""
import os

import numpy as np
import pytest

from locomouse_cpp_b200.types import Config, Model

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_nms.npz")


def _detector_for(rows, cols, bw, bh, cand_cap=256):
    from locomouse_cpp_b200.api import Detector

    cfg = Config(vid_rows=2 * rows, vid_cols=cols, n_rows=2 * rows, n_cols=cols, bb_w=cols, bb_h_bottom=rows, bb_h_side=rows,
                 cand_cap=cand_cap, det_cap=8192, match_cap=256)
    w = np.zeros((bh, bw), np.float32)
    model = Model(w=[[w, w, w], [w, w, w]], rho=[[1.0] * 3, [1.0] * 3])
    bkg = np.zeros((cfg.vid_rows, cfg.vid_cols), np.uint8)
    calib = np.arange(cfg.n_rows * cfg.n_cols, dtype=np.int32).reshape(cfg.n_rows, cfg.n_cols)
    return Detector(cfg, model, bkg, calib, device=0)


def _same(a, b):
    return len(a) == len(b) and np.array_equal(a["x"], b["x"]) and np.array_equal(a["y"], b["y"]) and np.array_equal(
        np.ascontiguousarray(a["s"]).view(np.uint64), np.ascontiguousarray(b["s"]).view(np.uint64))


def test_nms_kernels_equal_the_reference_code():
    z = np.load(GOLD)
    names = sorted({k.rsplit("_", 1)[0] for k in z.files})
    total = 0
    for n in names:
        s, (bw, bh) = z[f"{n}_scores"], z[f"{n}_box"]
        det = _detector_for(s.shape[0], s.shape[1], int(bw), int(bh))
        for feat in (0, 1):
            got_b = det.debug_nms(0, feat, s)
            got_s = det.debug_nms(1, feat, s)
            assert _same(got_b, z[f"{n}_nmsmax"]), f"k_nms_bottom != reference nmsMax on {n}"
            assert _same(got_s, z[f"{n}_peak"]), f"k_nms_side != reference peakClustering on {n}"
            total += len(got_b) + len(got_s)
        det.close()
    assert total > 1000


def test_nms_kernels_ties_follow_the_documented_total_order(oracle):
    """Equal scores (std::sort leaves their order unspecified in the reference): kernel == oracle's total order
    (score descending, then row-major pixel index ascending), including the large-list launch class (> 1024 positives)."""
    rng = np.random.Generator(np.random.PCG64(7))
    for rows, cols, bw, bh, levels, dens in [(40, 64, 9, 9, 5, 0.3), (64, 96, 12, 12, 3, 0.6), (48, 48, 30, 30, 8, 0.2)]:
        s = rng.integers(1, levels + 1, (rows, cols)).astype(np.float32)
        s[rng.random((rows, cols)) > dens] = -1.0
        det = _detector_for(rows, cols, bw, bh, cand_cap=1024)
        want_b = oracle.nms_max(s, bw, bh)
        want_s = oracle.peak_clustering(s, bw, bh)
        got_b, got_s = det.debug_nms(0, 0, s), det.debug_nms(1, 1, s)
        det.close()
        assert [(int(c["x"]), int(c["y"]), float(c["s"])) for c in got_b] == want_b
        assert [(int(c["x"]), int(c["y"]), float(c["s"])) for c in got_s] == want_s
        assert len(want_b) > 0
