"""The oracle against the committed golden vectors (CPU only): real-OpenCV outputs captured with
cv2 4.13 (cv2_primitives.npz) and frozen end-to-end cases (detect_*.npz)."""
import os

import numpy as np
import pytest

from _golden import DETECT_CASES, GOLDEN_DIR, load_detect_case
from locomouse_cpp_b200.types import Config, diff_results


@pytest.fixture(scope="module")
def prim():
    return np.load(os.path.join(GOLDEN_DIR, "cv2_primitives.npz"))


def test_preprocess_vs_opencv_vector(oracle, prim):
    bkg, frame, calib = prim["pre_bkg"], prim["pre_frame"], prim["pre_calib"]
    cfg = Config(vid_rows=bkg.shape[0], vid_cols=bkg.shape[1], n_rows=calib.shape[0], n_cols=calib.shape[1], bb_w=10,
                 bb_h_bottom=10, bb_h_side=10, flip=True, imadjust=False)
    got, _ = oracle.preprocess(cfg, bkg, calib, frame)
    assert np.array_equal(got, prim["pre_norm_gather_flip"])


def test_filter2d_vs_opencv_vector(oracle, prim):
    I = prim["f2d_image"]
    for i in range(3):
        k, ref = prim[f"f2d_k{i}"], prim[f"f2d_out{i}"]
        got = oracle.correlate(I, k, 0.25, -5, -5, ref.shape[1], ref.shape[0], fma_mode=False)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), f"kernel {i}"
        fused = oracle.correlate(I, k, 0.25, -5, -5, ref.shape[1], ref.shape[0], fma_mode=True)
        assert np.abs(fused - ref).max() <= 1e-5 * np.abs(ref).max()


def test_largest_region_vs_opencv_vector(oracle, prim):
    bins, larg = prim["cc_bin"], prim["cc_largest"]
    for i in range(bins.shape[0]):
        conn = 4 if i < 12 else 8
        assert np.array_equal(oracle.largest_region(bins[i], conn), larg[i]), f"case {i}"


@pytest.mark.parametrize("name", DETECT_CASES)
def test_oracle_reproduces_frozen_case(oracle, name):
    c = load_detect_case(name)
    got = oracle.detect(c["cfg"], c["model"], c["bkg"], c["calib"], c["frames"], c["bb_x"], c["bb_y_side"],
                        c["bb_y_bottom"], prev_frame=c["prev"], first_frame_index=c["first"], n_threads=2)
    assert diff_results(got, c["expected"]) == []
    assert got.checksum() == c["expected"].checksum()
    assert int(got.n_bottom.sum()) > 0 and int(got.n_side.sum()) > 0
