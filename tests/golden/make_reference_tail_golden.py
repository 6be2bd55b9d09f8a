"""Golden vectors produced by the REFERENCE's own tail code -- detectTail / detectLineCandidates / selectLargestRegion,
LocoMouse_class.cpp:2541-2767, compiled from /root/reference into oracle/_ref/libref_nms.so (oracle/ref_glue.cpp::
ref_detect_tail) -- with connectedComponentsWithStats executed by the real OpenCV (cv2).  Inputs are tail score maps
(quantised to quarters so that the file stays small); outputs the 3 x 15 tail tracks and TAIL_MASK."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_nms as ref  # noqa: E402


def cases(seed=2024, n=40):
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for it in range(n):
        hb, hs, tw = int(rng.integers(8, 48)), int(rng.integers(8, 40)), int(rng.integers(4, 64))
        yy, xx = np.mgrid[0:hb, 0:tw]
        y2, x2 = np.mgrid[0:hs, 0:tw]
        sb = rng.normal(-1.0, 1, (hb, tw))
        ss = rng.normal(-1.0, 1, (hs, tw))
        sl = rng.uniform(-0.3, 0.3)
        sb += 4 * np.exp(-((yy - (rng.uniform(0, hb) + sl * xx)) ** 2) / (2 * rng.uniform(0.8, 3) ** 2)) * (xx > rng.integers(0, max(1, tw // 2)))
        ss += 4 * np.exp(-((y2 - (rng.uniform(0, hs) + sl * x2)) ** 2) / (2 * rng.uniform(0.8, 3) ** 2))
        if it == 0:
            sb[:] = -1                # no foreground at all: all -1, zero mask
        if it == 1:
            sb[:] = -1                # foreground in columns 0 and 1: segment 0 = column 0, centroid x == 0 -> no side z (Q13)
            sb[3:6, 0:2] = 1
            ss[:] = 1
        if it == 2:
            sb[:, 10:] = -1           # fewer columns than tail points: zero-width segments stay -1
        q = lambda a: np.clip(np.rint(a * 4), -127, 127).astype(np.int8)
        out.append(dict(sb=q(sb), ss=q(ss), conn=int(rng.choice([4, 8]))))
    return out


def maps(c):
    return c["sb"].astype(np.float32) / 4, c["ss"].astype(np.float32) / 4


def main():
    out = {}
    for i, c in enumerate(cases()):
        sb, ss = maps(c)
        t, m = ref.detect_tail(sb, ss, c["conn"], 15)
        out[f"c{i:02d}_sb"], out[f"c{i:02d}_ss"], out[f"c{i:02d}_conn"] = c["sb"], c["ss"], np.int32(c["conn"])
        out[f"c{i:02d}_tracks"], out[f"c{i:02d}_mask"] = t, np.packbits(m > 0)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_tail.npz")
    np.savez_compressed(path, **out)
    print("cases", len(out) // 5, "bytes", os.path.getsize(path))


if __name__ == "__main__":
    main()
