"""Host-side C++ mirror of the reference classes (locomouse_cpp_b200/host).

CPU part: the value types (Candidate / P22D) behave like Candidates/Candidates.cpp, the driver keeps
main.cpp's error convention.  GPU part: the reference's main.cpp call sequence, run through the C++
classes on files, yields exactly the oracle's candidates / P22D records / tail tracks.
"""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "locomouse_cpp_b200", "host")


@pytest.fixture(scope="module")
def host_build():
    p = subprocess.run(["make", "-C", HOST, "test_candidates"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    return HOST


def _build_driver():
    if not os.path.exists(os.path.join(ROOT, "locomouse_cpp_b200", "liblocomouse_b200.so")):
        pytest.skip("CUDA library not built")
    p = subprocess.run(["make", "-C", HOST], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    return os.path.join(HOST, "locomouse_b200")


def test_value_types(host_build):
    p = subprocess.run([os.path.join(host_build, "test_candidates")], capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout.strip().endswith("ok"), p.stdout + p.stderr


def test_driver_error_convention(tmp_path):
    """main.cpp:94-101: invalid inputs are reported and turned into EXIT_FAILURE, not a crash."""
    exe = _build_driver()
    p = subprocess.run([exe, "1", "only", "three"], capture_output=True, text=True)
    assert p.returncode == 1 and "Invalid inputs" in p.stdout and "Total Elapsed time" in p.stdout
    cfg = tmp_path / "config.yml"
    cfg.write_text("conn_comp_connectivity: 5\n")
    p = subprocess.run([exe, "1", str(cfg), "v", "b", "m", "c", "R", str(tmp_path)], capture_output=True, text=True)
    assert p.returncode == 1 and "conn_comp_connectivity must be either 4 or 8" in p.stdout
    cfg.write_text("conn_comp_connectivity: 8\n")
    p = subprocess.run([exe, "1", str(cfg), str(tmp_path / "missing.lmv"), "b", "m", "c", "R", str(tmp_path)],
                       capture_output=True, text=True)
    assert p.returncode == 1 and "Runtime Error" in p.stdout and "Cannot open file" in p.stdout


# ---------------------------------------------------------------------------------------------------
from locomouse_cpp_b200.lmfiles import write_problem_files  # noqa: E402


def read_output(path):
    """LMO1 (LocoMouse::exportResults) -> per frame dict."""
    buf = open(path, "rb").read()
    assert buf[:4] == b"LMO1"
    off = 4
    n, nt = struct.unpack_from("<ii", buf, off)
    off += 8
    frames = []
    for _ in range(n):
        tail = np.frombuffer(buf, "<i4", 3 * nt, off).reshape(3, nt)
        off += 12 * nt
        feats = []
        for _feat in range(2):
            lists = []
            for _l in range(2):
                (k,) = struct.unpack_from("<i", buf, off)
                off += 4
                c = []
                for _i in range(k):
                    x, y, s = struct.unpack_from("<iid", buf, off)
                    off += 16
                    c.append((x, y, s))
                lists.append(c)
            matches = []
            for _i in range(len(lists[0])):
                (m,) = struct.unpack_from("<i", buf, off)
                off += 4
                mm = []
                for _q in range(m):
                    y, s = struct.unpack_from("<id", buf, off)
                    off += 12
                    mm.append((y, s))
                matches.append(mm)
            feats.append((lists[0], lists[1], matches))
        frames.append((tail, feats))
    assert off == len(buf)
    return frames


@pytest.mark.gpu
@pytest.mark.parametrize("method,flip", [("TM", False), ("TM_DE", True)])
def test_main_sequence_matches_oracle(tmp_path, oracle, method, flip):
    from locomouse_cpp_b200 import synth

    exe = _build_driver()
    spec = synth.SynthSpec(method=method, flip=flip)
    n = 7
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, n, seed=1000)
    frames = frames.numpy()
    ref = oracle.detect(cfg, model, bkg, calib, frames, bx, bs, bb, n_threads=4)
    # batch_frames 3 -> chunks of 3, 3, 1 frames: exercises the previous-frame halo between chunks
    write_problem_files(tmp_path, cfg, model, bkg, calib, frames, bx, bs, bb, spec.side_h, extra_cfg="batch_frames: 3\n")
    meth = {"base": "0", "TM": "1", "TM_DE": "2"}[method]
    p = subprocess.run([exe, meth, str(tmp_path / "config.yml"), str(tmp_path / "video.lmv"), str(tmp_path / "bkg.lmi"),
                        str(tmp_path / "model.lmm"), str(tmp_path / "calib.lmc"), "L" if flip else "R", str(tmp_path)],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    out = read_output(tmp_path / "output_video.lmo")
    assert len(out) == n
    total = 0
    for f, (tail, feats) in enumerate(out):
        assert np.array_equal(tail, ref.tail[f])
        for feat in range(2):
            cb, cs, matches = feats[feat]
            assert cb == ref.candidates_bottom(f, feat)
            assert cs == ref.candidates_side(f, feat)
            want = ref.p22d(f, feat)
            assert [m for m in matches] == [w[1] for w in want]
            total += len(cb)
    assert total > 0


@pytest.mark.gpu
def test_tm_de_pass1_on_device_then_detection(tmp_path, oracle):
    """LocoMouse_TM_DE with no pass-1 file: getBoundingBox() runs computeMouseBox_DE on the device
    (lm_bounding_box_tm_de) + the host moving average; detection then uses those boxes.  Everything equals the oracle."""
    from locomouse_cpp_b200 import synth
    from locomouse_cpp_b200.types import bb_de_params

    exe = _build_driver()
    spec = synth.SynthSpec(method="TM_DE")
    n = 8
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, n, seed=1000)
    frames = frames.numpy()
    want_bx, _, _ = oracle.bounding_box_tm_de(cfg, bkg, calib, frames, bb_de_params(cfg, side_h=spec.side_h), window=5)
    ref = oracle.detect(cfg, model, bkg, calib, frames, want_bx, bs, bb, n_threads=4)
    assert ref.rc == 0
    write_problem_files(tmp_path, cfg, model, bkg, calib, frames, bx, bs, bb, spec.side_h, extra_cfg="batch_frames: 5\n",
                        with_boxes=False)
    p = subprocess.run([exe, "2", str(tmp_path / "config.yml"), str(tmp_path / "video.lmv"), str(tmp_path / "bkg.lmi"),
                        str(tmp_path / "model.lmm"), str(tmp_path / "calib.lmc"), "R", str(tmp_path)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    out = read_output(tmp_path / "output_video.lmo")
    assert len(out) == n
    for f, (tail, feats) in enumerate(out):
        assert np.array_equal(tail, ref.tail[f])
        for feat in range(2):
            cb, cs, matches = feats[feat]
            assert cb == ref.candidates_bottom(f, feat) and cs == ref.candidates_side(f, feat)
            assert matches == [w[1] for w in ref.p22d(f, feat)]


@pytest.mark.gpu
def test_main_sequence_builds_tracker_costs_on_device(tmp_path, oracle):
    """With a location_prior in the configuration the main.cpp sequence's computeUnaryCostsBottom / computePairwiseCostsBottom
    (LocoMouse_class.cpp:873-919) hand out, per frame, the MyMat / MATSPARSE that the device built for the whole chunk;
    chunks of 3 frames exercise the one-frame halo of the pairwise transition.  Everything equals the oracle bit for bit."""
    from locomouse_cpp_b200 import synth
    from locomouse_cpp_b200.types import location_priors, pairwise_params

    exe = _build_driver()
    spec = synth.SynthSpec(method="TM")
    n = 8
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, n, seed=1000)
    frames = frames.numpy()
    ref = oracle.detect(cfg, model, bkg, calib, frames, bx, bs, bb, n_threads=4)
    rows = [(0.8, 0.25, 0.5, 0.4, 1.0, 0.0, 0.5), (0.8, 0.75, 0.5, 0.4, 1.0, 0.5, 1.0), (0.3, 0.25, 0.4, 0.0, 0.6, 0.0, 0.5),
            (0.3, 0.75, 0.35, 0.0, 0.6, 0.5, 1.0), (0.95, 0.5, 0.6, 0.5, 1.0, 0.0, 1.0)]
    flat = ", ".join(repr(float(v)) for r in rows for v in r)
    write_problem_files(tmp_path, cfg, model, bkg, calib, frames, bx, bs, bb, spec.side_h,
                        extra_cfg=f"batch_frames: 3\nlocation_prior: [{flat}]\nmax_displacement_bottom: 40\n")
    p = subprocess.run([exe, "1", str(tmp_path / "config.yml"), str(tmp_path / "video.lmv"), str(tmp_path / "bkg.lmi"),
                        str(tmp_path / "model.lmm"), str(tmp_path / "calib.lmc"), "R", str(tmp_path)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    buf = (tmp_path / "costs_video.lmo").read_bytes()
    assert buf[:4] == b"LMC1"
    (nf,) = struct.unpack_from("<i", buf, 4)
    assert nf == n
    off = 8
    pri = [location_priors(rows[:4]), location_priors(rows[4:])]
    P = pairwise_params(cfg.bb_w, cfg.bb_h_bottom, max_displacement=40)
    nnz = 0
    for f in range(n):
        for feat in range(2):
            nr, nc = struct.unpack_from("<ii", buf, off)
            off += 8
            U = np.frombuffer(buf, np.float64, nr * nc, off).reshape(nc, nr).T
            off += 8 * nr * nc
            want = oracle.unary_cost_box(ref.candidates_bottom(f, feat), cfg.bb_w, cfg.bb_h_bottom, pri[feat])
            assert U.shape == want.shape and np.array_equal(np.ascontiguousarray(U).view(np.uint64), want.view(np.uint64)), (f, feat)
            if f == 0:
                continue
            sr, sc, nz = struct.unpack_from("<iii", buf, off)
            off += 12
            jc = np.frombuffer(buf, np.int32, sc + 1, off)
            off += 4 * (sc + 1)
            ir = np.frombuffer(buf, np.int32, nz, off)
            off += 4 * nz
            pr = np.frombuffer(buf, np.float64, nz, off)
            off += 8 * nz
            wr, wc, wjc, wir, wpr = oracle.pairwise_potential(ref.candidates_bottom(f - 1, feat), ref.candidates_bottom(f, feat), P)
            assert (sr, sc) == (wr, wc) and np.array_equal(jc, wjc) and np.array_equal(ir, wir), (f, feat)
            assert np.array_equal(pr.view(np.uint64), wpr.view(np.uint64)), (f, feat)
            nnz += nz
    assert off == len(buf) and nnz > n * P.ong_w * P.ong_h


# ---- the reference's own on-disk formats (OpenCV YAML) ------------------------------------------------------
def _write_opencv_yaml(path, items):
    """Written by the REAL OpenCV (cv2.FileStorage), as the reference's users' files are."""
    cv2 = pytest.importorskip("cv2")
    fs = cv2.FileStorage(str(path), cv2.FILE_STORAGE_WRITE)
    for k, v in items:
        fs.write(k, v)
    fs.release()


def test_opencv_yaml_reader_reads_what_opencv_writes(tmp_path):
    """host/cv_yaml.hpp against files written by cv2.FileStorage: matrices of every depth the reference uses (f64 templates,
    i32 calibration map and view boxes, f32, u8), scalars, strings, infinities; missing nodes read as 0 / empty."""
    p = subprocess.run(["make", "-C", HOST, "test_cv_yaml"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    rng = np.random.default_rng(1)
    A = rng.normal(0, 1e-3, (30, 30))
    A[0, :3] = [np.inf, -np.inf, 1.0]
    B = rng.integers(0, 680000, (13, 17)).astype(np.int32)
    Cm = rng.normal(0, 1, (2, 3)).astype(np.float32)
    U = rng.integers(0, 256, (2, 5)).astype(np.uint8)
    f = tmp_path / "m.yml"
    _write_opencv_yaml(f, [("modelPaw_side", A), ("ind_warp_mapping", B), ("f", Cm), ("u", U), ("biasPaw_side", -1.2345678901234567),
                           ("n", 7), ("name", "abc def")])
    out = subprocess.run([os.path.join(HOST, "test_cv_yaml"), str(f), "modelPaw_side", "ind_warp_mapping", "f", "u", "biasPaw_side", "n",
                          "name", "nothing"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-300:]
    lines = out.stdout.strip().split("\n")
    vals = lambda l: np.array([float(x) for x in l.split()[5:]])
    assert lines[0].split()[1:5] == ["matrix", "30", "30", "d"] and np.array_equal(vals(lines[0]).reshape(30, 30), A)
    assert lines[1].split()[1:5] == ["matrix", "13", "17", "i"] and np.array_equal(vals(lines[1]).reshape(13, 17), B)
    assert lines[2].split()[4] == "f" and np.array_equal(vals(lines[2]).reshape(2, 3).astype(np.float32), Cm)
    assert lines[3].split()[4] == "u" and np.array_equal(vals(lines[3]).reshape(2, 5), U)
    assert lines[4].split()[:3] == ["biasPaw_side", "scalar", repr(-1.2345678901234567)]
    assert lines[5].split()[:3] == ["n", "scalar", "7"] and lines[6].split()[1] == "string" and lines[6].endswith("[abc def]")
    assert lines[7].split()[:3] == ["nothing", "missing", "0"]
    out = subprocess.run([os.path.join(HOST, "test_cv_yaml"), str(tmp_path / "absent.yml"), "x"], capture_output=True, text=True)
    assert out.returncode == 1 and "not-opened" in out.stdout


@pytest.mark.gpu
def test_main_sequence_with_opencv_yaml_model_calibration_and_config(tmp_path, oracle):
    """The class mirror driven with the reference's own file formats: the video as an AVI and the background as a PNG, both
    written by the real OpenCV (cv2.VideoWriter / cv2.imwrite), the model (six f64 matrices + biases), the calibration
    (ind_warp_mapping, view_boxes) and the configuration (location_prior as a 5 x 7 matrix node) written by cv2.FileStorage:
    detection results and cost matrices equal the oracle."""
    cv2 = pytest.importorskip("cv2")
    from locomouse_cpp_b200 import synth
    from locomouse_cpp_b200.types import location_priors

    exe = _build_driver()
    spec = synth.SynthSpec(method="TM", warp=True)
    n = 5
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, n, seed=1000)
    frames = frames.numpy()
    ref = oracle.detect(cfg, model, bkg, calib, frames, bx, bs, bb, n_threads=4)
    write_problem_files(tmp_path, cfg, model, bkg, calib, frames, bx, bs, bb, spec.side_h, extra_cfg="batch_frames: 4\n")
    names = (("Paw", 0), ("Snout", 1), ("Tail", 2))
    items = []
    for view, vn in ((1, "side"), (0, "bottom")):
        for fn, k in names:
            items.append((f"model{fn}_{vn}", np.asarray(model.w[view][k], np.float64)))     # templates stored as doubles, as MATLAB exports them
            items.append((f"bias{fn}_{vn}", float(model.rho[view][k])))
    _write_opencv_yaml(tmp_path / "model.yml", items)
    boxes = np.array([[0, 0, cfg.n_cols, spec.side_h], [0, spec.side_h, cfg.n_cols, cfg.n_rows - spec.side_h]], np.int32)
    _write_opencv_yaml(tmp_path / "calibration.yml", [("ind_warp_mapping", np.ascontiguousarray(calib, np.int32)), ("view_boxes", boxes)])
    rows = [(0.8, 0.25, 0.5, 0.4, 1.0, 0.0, 0.5), (0.8, 0.75, 0.5, 0.4, 1.0, 0.5, 1.0), (0.3, 0.25, 0.4, 0.0, 0.6, 0.0, 0.5),
            (0.3, 0.75, 0.35, 0.0, 0.6, 0.5, 1.0), (0.95, 0.5, 0.6, 0.5, 1.0, 0.0, 1.0)]
    # append the matrix node to the scalar configuration, as cv::FileStorage lays it out
    _write_opencv_yaml(tmp_path / "prior.yml", [("location_prior", np.array(rows, np.float64))])
    node = (tmp_path / "prior.yml").read_text().split("---\n", 1)[1]
    (tmp_path / "config.yml").write_text((tmp_path / "config.yml").read_text() + node)
    vw = cv2.VideoWriter(str(tmp_path / "video.avi"), 0, 30.0, (cfg.vid_cols, cfg.vid_rows), False)  # uncompressed 8-bit grey
    assert vw.isOpened()
    for f in frames:
        vw.write(f)
    vw.release()
    assert cv2.imwrite(str(tmp_path / "bkg.png"), bkg)
    p = subprocess.run([exe, "1", str(tmp_path / "config.yml"), str(tmp_path / "video.avi"), str(tmp_path / "bkg.png"),
                        str(tmp_path / "model.yml"), str(tmp_path / "calibration.yml"), "R", str(tmp_path)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    out = read_output(tmp_path / "output_video.lmo")
    assert len(out) == n
    for f, (tail, feats) in enumerate(out):
        assert np.array_equal(tail, ref.tail[f])
        for feat in range(2):
            cb, cs, matches = feats[feat]
            assert cb == ref.candidates_bottom(f, feat) and cs == ref.candidates_side(f, feat)
            assert [m for m in matches] == [w[1] for w in ref.p22d(f, feat)]
    assert (tmp_path / "output_video.yml").exists()   # tracks: the reference's output file
    # the priors reached the cost builders: first frame's paw unary matrix
    buf = (tmp_path / "costs_video.lmo").read_bytes()
    nr, nc = struct.unpack_from("<ii", buf, 8)
    U = np.frombuffer(buf, np.float64, nr * nc, 16).reshape(nc, nr).T
    want = oracle.unary_cost_box(ref.candidates_bottom(0, 0), cfg.bb_w, cfg.bb_h_bottom, location_priors(rows[:4]))
    assert np.array_equal(np.ascontiguousarray(U).view(np.uint64), want.view(np.uint64))


@pytest.mark.gpu
def test_base_class_pass1_on_device_then_detection(tmp_path, oracle):
    """Method 0 with no pass-1 file: LocoMouse::computeBoundingBox runs computeMouseBox on the device (lm_bounding_box_base),
    computeMouseBoxSize and the moving averages on the host; detection then uses those boxes.  With the reference's own
    reading of the integer sums the box is empty (and the driver says so); with pass1_integer_sums everything equals the
    oracle's pass 1 + detection."""
    import copy

    from locomouse_cpp_b200 import synth
    from locomouse_cpp_b200.types import bb_base_params

    exe = _build_driver()
    spec = synth.SynthSpec(method="base")
    n = 9
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, n, seed=1000)
    frames = frames.numpy()
    args = lambda d: [exe, "0", str(d / "config.yml"), str(d / "video.lmv"), str(d / "bkg.lmi"), str(d / "model.lmm"), str(d / "calib.lmc"), "R", str(d)]
    write_problem_files(tmp_path, cfg, model, bkg, calib, frames, bx, bs, bb, spec.side_h, extra_cfg="batch_frames: 4\n", with_boxes=False)
    p = subprocess.run(args(tmp_path), capture_output=True, text=True)
    assert p.returncode == 1 and "the mouse box is empty" in p.stdout, p.stdout
    d2 = tmp_path / "int"
    d2.mkdir()
    write_problem_files(d2, cfg, model, bkg, calib, frames, bx, bs, bb, spec.side_h, extra_cfg="batch_frames: 4\npass1_integer_sums: 1\n", with_boxes=False)
    p = subprocess.run(args(d2), capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    box, _ = oracle.bounding_box_base(cfg, bkg, calib, frames, bb_base_params(cfg, side_h=spec.side_h, sums_as_float=0))
    w, hb, hs = oracle.mouse_box_size(box[:, 3], box[:, 4], box[:, 5])
    c2 = copy.copy(cfg)
    c2.bb_w, c2.bb_h_bottom, c2.bb_h_side = w, hb, hs   # tail_w follows (property)
    wx, wyb, wys = (oracle.vecmovingaverage(box[:, k], 5) for k in (0, 1, 2))
    ref = oracle.detect(c2, model, bkg, calib, frames, wx, wys, wyb, n_threads=4)
    assert ref.rc == 0
    out = read_output(d2 / "output_video.lmo")
    assert len(out) == n
    for f, (tail, feats) in enumerate(out):
        assert np.array_equal(tail, ref.tail[f])
        for feat in range(2):
            cb, cs, matches = feats[feat]
            assert cb == ref.candidates_bottom(f, feat) and cs == ref.candidates_side(f, feat)
            assert matches == [m[1] for m in ref.p22d(f, feat)]


@pytest.mark.gpu
def test_tm_pass1_on_device_then_detection(tmp_path, oracle):
    """Method 1 with no pass-1 file: LocoMouse_TM::computeBoundingBox runs computeMouseBox_DD on the device
    (lm_bounding_box_tm: imadjust_default, bands, threshold, bwAreaOpen, disk filter from a diskfilter.yml written by the real
    OpenCV, imfill, column sums) + the host moving average; detection then uses those boxes.  With the reference's own reading
    of the integer sums every bb_x is -1 and the crop fails, as it does in the reference; with pass1_integer_sums everything
    equals the oracle's pass 1 + detection."""
    cv2 = pytest.importorskip("cv2")
    from locomouse_cpp_b200 import synth
    from locomouse_cpp_b200.types import bb_tm_params

    exe = _build_driver()
    spec = synth.SynthSpec(method="TM")
    n = 9
    cfg, model, bkg, calib, frames, bx, bs, bb = synth.make_problem(spec, n, seed=1000)
    frames = frames.numpy()
    y, x = np.mgrid[-5:6, -5:6]
    disk = ((x * x + y * y) <= 5.5 ** 2).astype(np.float64)
    disk = (disk / disk.sum()).astype(np.float32)
    fs = cv2.FileStorage(str(tmp_path / "diskfilter.yml"), cv2.FILE_STORAGE_WRITE)
    fs.write("H", disk)
    fs.release()
    tm_cfg = (f"bw_threshold_bottom: 40\nbw_threshold_side: 40\nmin_pixel_count: 25\nzero_col_pre: 0\nzero_col_post: {cfg.n_cols}\n"
              f"zero_row_pre: 0\nzero_row_post: {spec.side_h}\ndisk_filter_file: {tmp_path / 'diskfilter.yml'}\n")
    args = lambda d: [exe, "1", str(d / "config.yml"), str(d / "video.lmv"), str(d / "bkg.lmi"), str(d / "model.lmm"), str(d / "calib.lmc"), "R", str(d)]
    d1 = tmp_path / "ref"
    d1.mkdir()
    write_problem_files(d1, cfg, model, bkg, calib, frames, bx, bs, bb, spec.side_h, extra_cfg="batch_frames: 4\n" + tm_cfg, with_boxes=False)
    p = subprocess.run(args(d1), capture_output=True, text=True)
    assert p.returncode == 1, p.stdout + p.stderr   # bb_x = -1 for every frame: the crop leaves the image (ROI exception in the reference)
    d2 = tmp_path / "int"
    d2.mkdir()
    write_problem_files(d2, cfg, model, bkg, calib, frames, bx, bs, bb, spec.side_h, extra_cfg="batch_frames: 4\npass1_integer_sums: 1\n" + tm_cfg,
                        with_boxes=False)
    p = subprocess.run(args(d2), capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    P = bb_tm_params(cfg, disk, side_h=spec.side_h, side_threshold=40, min_pixel_count=25, sums_as_float=0)
    raw, _ = oracle.bounding_box_tm(cfg, bkg, calib, frames, P)
    want_bx = oracle.vecmovingaverage(raw, 5)
    assert (raw > 0).all()
    ys = np.full(n, 164, np.uint32)
    yb = np.full(n, cfg.n_rows - 1, np.uint32)
    ref = oracle.detect(cfg, model, bkg, calib, frames, want_bx, ys, yb, n_threads=4)
    assert ref.rc == 0
    out = read_output(d2 / "output_video.lmo")
    assert len(out) == n
    total = 0
    for f, (tail, feats) in enumerate(out):
        assert np.array_equal(tail, ref.tail[f])
        for feat in range(2):
            cb, cs, matches = feats[feat]
            assert cb == ref.candidates_bottom(f, feat) and cs == ref.candidates_side(f, feat)
            assert matches == [m[1] for m in ref.p22d(f, feat)]
            total += len(cb)
    assert total > 0
