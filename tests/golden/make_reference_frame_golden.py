"""Golden candidate lists produced by the REFERENCE's own detectBottomCandidates / detectSideCandidates code (compiled from
/root/reference, oracle/ref_glue.cpp::ref_detect_candidates) on seeded synthetic frames: see tests/_frame_glue.py for how
the frames reach it.  tests/test_oracle_vs_reference.py requires oracle.detect to reproduce them bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _frame_glue as G  # noqa: E402
from oracle import oracle  # noqa: E402
from oracle.reference_nms import CAND  # noqa: E402


def main():
    out = {}
    total = 0
    for ci, kw in enumerate(G.CASES):
        cfg, model, bkg, calib, frames, bx, bs, bb = G.case_problem(kw)
        for f in range(len(frames)):
            lists = G.reference_frame(oracle, cfg, model, bkg, calib, frames[f], bx[f], bs[f], bb[f])
            for k, lst in enumerate(lists):
                out[f"c{ci}_f{f}_l{k}"] = np.array(lst, CAND) if lst else np.zeros(0, CAND)
                total += len(lst)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_frame_candidates.npz")
    np.savez_compressed(path, **out)
    print("lists", len(out), "candidates", total, "bytes", os.path.getsize(path))


if __name__ == "__main__":
    main()
