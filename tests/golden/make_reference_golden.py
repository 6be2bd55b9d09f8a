"""Golden vectors produced by the REFERENCE's own nmsMax / peakClustering code (oracle/_ref/libref_nms.so, built by
`make -C oracle ref` where /root/reference is mounted): random and adversarial score maps -> candidate lists.
The committed file lets the pin travel to machines without the reference (tests/test_oracle_vs_reference.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_nms as ref  # noqa: E402


def score_maps():
    rng = np.random.Generator(np.random.PCG64(424242))
    cases = []
    for i, (rows, cols, bw, bh, dens) in enumerate([(40, 60, 9, 9, 0.08), (64, 96, 30, 30, 0.05), (50, 50, 12, 8, 0.2),
                                                    (30, 80, 7, 15, 0.5), (90, 120, 30, 30, 0.02), (25, 25, 4, 4, 0.9),
                                                    (60, 60, 16, 16, 0.01), (48, 64, 30, 22, 0.1)]):
        s = rng.normal(0, 1, (rows, cols)).astype(np.float32)
        # smooth blobs + noise so that clusters form; distinct scores (no sort ties)
        yy, xx = np.mgrid[0:rows, 0:cols]
        for _ in range(4):
            cy, cx = rng.uniform(0, rows), rng.uniform(0, cols)
            s += (3.0 * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * rng.uniform(2, 6) ** 2))).astype(np.float32)
        thr = np.quantile(s, 1 - dens)
        s = (s - thr).astype(np.float32)
        s += (np.arange(s.size, dtype=np.float32).reshape(s.shape) * np.float32(1e-7))  # break exact ties
        cases.append((f"case{i}", s, bw, bh))
    # half-way cases for the two rounding rules: two equal-weight detections one pixel apart
    s = np.full((12, 12), -1, np.float32)
    s[3, 4] = 1.0
    s[3, 5] = 1.0  # equal weights: the mean is exactly 4.5; the tie cannot change the result (symmetric sums)
    s[8, 2] = 2.0
    s[9, 2] = 2.0
    cases.append(("halfway", s, 6, 6))
    # 3-detection chain that separates nmsMax (chain suppression through discarded detections) from peakClustering
    s = np.full((10, 40), -1, np.float32)
    s[5, 5], s[5, 8], s[5, 11] = 3.0, 2.0, 1.0
    cases.append(("chain", s, 10, 10))
    cases.append(("empty", np.full((8, 8), -1, np.float32), 4, 4))
    return cases


def main():
    out = {}
    for name, s, bw, bh in score_maps():
        out[f"{name}_scores"] = s
        out[f"{name}_box"] = np.array([bw, bh], np.int32)
        out[f"{name}_nmsmax"] = ref.nms_max(s, bw, bh, 0.5)
        out[f"{name}_peak"] = ref.peak_clustering(s, bw, bh, 0.5)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_nms.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.endswith(("nmsmax", "peak"))})


if __name__ == "__main__":
    main()
