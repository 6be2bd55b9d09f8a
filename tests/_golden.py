"""Loader for the committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import json
import os

import numpy as np

from locomouse_cpp_b200 import synth
from locomouse_cpp_b200.types import Model, Results

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DETECT_CASES = ("detect_small_tm", "detect_small_tmde_flip_warp", "detect_small_base_muladd")


def load_detect_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    spec_kw = json.loads(str(z["spec"]))
    if spec_kw.get("tshapes") is not None:
        spec_kw["tshapes"] = tuple(tuple(tuple(s) for s in v) for v in spec_kw["tshapes"])
    spec = synth.SynthSpec(**spec_kw)
    cfg = spec.config()
    model = Model(w=[[z[f"w_{v}_{k}"] for k in range(3)] for v in range(2)], rho=z["rho"].tolist())
    frames = z["frames"]
    exp = Results(frames.shape[0], cfg.cand_cap, cfg.match_cap, cfg.n_tail_points)
    for a in Results.ARRAYS:
        getattr(exp, a)[...] = z["exp_" + a]
    prev = z["prev"] if z["prev"].size else None
    return dict(spec=spec, cfg=cfg, model=model, bkg=z["bkg"], calib=z["calib"], frames=frames, bb_x=z["bb_x"],
                bb_y_side=z["bb_y_side"], bb_y_bottom=z["bb_y_bottom"], prev=prev, first=int(z["first"]), expected=exp)
