"""Writers for the raw input containers the C++ driver reads (locomouse_cpp_b200/host/lm_files.hpp): used by the tests and by
bench.py's measurement of the driver (the reference's main() sequence).  `d` is a pathlib.Path."""
import struct

import numpy as np


def write_problem_files(d, cfg, model, bkg, calib, frames, bx, bs, bb, side_h, extra_cfg="", with_boxes=True):
    """The raw containers of locomouse_cpp_b200/host/lm_files.hpp."""
    n, rows, cols = frames.shape
    with open(d / "video.lmv", "wb") as f:
        f.write(b"LMV1" + struct.pack("<iii", n, rows, cols))
        f.write(np.ascontiguousarray(frames, np.uint8).tobytes())
    with open(d / "bkg.lmi", "wb") as f:
        f.write(b"LMI1" + struct.pack("<ii", *bkg.shape))
        f.write(np.ascontiguousarray(bkg, np.uint8).tobytes())
    with open(d / "model.lmm", "wb") as f:
        f.write(b"LMM1")
        for v in range(2):
            for k in range(3):
                w = np.ascontiguousarray(model.w[v][k], np.float32)
                f.write(struct.pack("<iid", w.shape[0], w.shape[1], float(model.rho[v][k])))
                f.write(w.tobytes())
    with open(d / "calib.lmc", "wb") as f:
        f.write(b"LMC1" + struct.pack("<ii", *calib.shape))
        f.write(struct.pack("<8i", 0, 0, cfg.n_cols, side_h, 0, side_h, cfg.n_cols, cfg.n_rows - side_h))
        f.write(np.ascontiguousarray(calib, np.int32).tobytes())
    with open(d / "boxes.lmb", "wb") as f:
        f.write(b"LMB1" + struct.pack("<i", n))
        for a in (bx, bs, bb):
            f.write(np.ascontiguousarray(a, np.uint32).tobytes())
    (d / "config.yml").write_text(
        "%YAML:1.0\n"
        f"conn_comp_connectivity: {cfg.conn}\n"
        f"side_bottom_min_overlap: {cfg.min_overlap!r}\n"
        f"tail_sub_bounding_box: {cfg.tail_sub_bounding_box!r}\n"
        f"bb_width: {cfg.bb_w}\nbb_height_side: {cfg.bb_h_side}\n"
        + (f"bounding_box_file: {d / 'boxes.lmb'}\n" if with_boxes else "")
        + 
        f"fma_mode: {int(cfg.fma_mode)}\ncand_cap: {cfg.cand_cap}\ndet_cap: {cfg.det_cap}\nmatch_cap: {cfg.match_cap}\n"
        + extra_cfg)
