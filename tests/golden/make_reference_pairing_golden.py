"""Golden vectors produced by the REFERENCE's own pairing code -- matchingWithVelocityConstraint / xDist / matchViews /
checkVelCriterion, LocoMouse_class.cpp:1023-1267, compiled from /root/reference into oracle/_ref/libref_nms.so by
`make -C oracle ref` (oracle/ref_glue.cpp::ref_match_views).  Random candidate lists over two views of a small image pair
(current / previous) with moved patches, so that the velocity criterion accepts and rejects; plus the all-true boolD case
(SURVEY Q7).  The committed file lets the pin travel to machines without the reference
(tests/test_oracle_vs_reference.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_nms as ref  # noqa: E402

NR, NC, W, HB, HS = 72, 112, 64, 36, 28
SPRE = (15, 15)


def cases(seed=777, n=64):
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for it in range(n):
        twb, thb, tws, ths = (int(v) for v in rng.integers(4, 31, 4))
        x0, y0b, y0s = int(rng.integers(-4, NC - W + 4)), int(rng.integers(30, NR - HB + 4)), int(rng.integers(-3, 8))
        # few grey levels in 8x8 blocks: compresses well, still gives window counts on both sides of the criterion
        I = np.kron(rng.choice(np.array([0, 10, 40, 90], np.uint8), (NR // 8, NC // 8)), np.ones((8, 8), np.uint8))
        Ip = I.copy()
        for _ in range(int(rng.integers(0, 10))):
            y, x, h, w = int(rng.integers(0, NR - 8)), int(rng.integers(0, NC - 8)), int(rng.integers(2, 24)), int(rng.integers(2, 24))
            I[y:y + h, x:x + w] = 200
            Ip[y:y + h, x:x + w] = int(rng.choice([0, 180, 174, 175]))   # 200-175 = 25 is NOT > 25; 200-174 is
        nb, ns = int(rng.integers(0, 8)), int(rng.integers(0, 8))
        cx = rng.integers(0, W, 3)
        mk = lambda n_, h: [(int(np.clip(cx[rng.integers(0, 3)] + rng.integers(-12, 13), 0, W - 1)), int(rng.integers(0, h)),
                             float(np.float32(rng.uniform(0.01, 3)))) for _ in range(n_)]
        cb, cs = mk(nb, HB), mk(ns, HS)
        for (x, y, _), yo in [(c, y0b) for c in cb] + [(c, y0s) for c in cs]:   # movement under about half of the candidates
            if rng.integers(0, 2):
                h, w, yy, xx = int(rng.integers(2, 12)), int(rng.integers(2, 12)), max(0, yo + y - 4), max(0, x0 + x - 4)
                I[yy:yy + h, xx:xx + w] = 200
                Ip[yy:yy + h, xx:xx + w] = int(rng.choice([0, 174, 175]))
        if it == 0:      # Q7: one bottom and one side candidate within the overlap -> boolD all true -> no pairing at all
            cb, cs = [(20, 5, 1.5)], [(22, 7, 0.75)]
        if it == 1:      # same x for everything: all-true matrix with several rows / columns
            cb, cs = [(30, 3, 1.0), (30, 9, 2.0)], [(30, 4, 0.5), (30, 8, 0.25), (30, 12, 0.125)]
        T = float(rng.choice([0.7, 0.5, 0.9]))
        if int(twb * (1 - T)) == 0:
            T = 0.5          # ovlp == 0 divides by zero in the reference: NaN weights, CV_Assert(S >= 0) throws
        out.append(dict(tsz=(twb, thb, tws, ths), org=(x0, y0b, y0s), I=I, Ip=Ip, cb=cb, cs=cs, T=T, vel=int(it % 4 != 3)))
    return out


def run_reference(c):
    twb, thb, tws, ths = c["tsz"]
    x0, y0b, y0s = c["org"]
    return ref.match_views(c["cb"], c["cs"], c["vel"], (twb, thb), (tws, ths), c["T"], c["I"], c["Ip"], x0, y0b, y0s, HB, HS, W,
                           SPRE, SPRE)


def main():
    out = {}
    n_match = 0
    for i, c in enumerate(cases()):
        r = run_reference(c)
        out[f"c{i:02d}_img"] = np.stack([c["I"], c["Ip"]])
        out[f"c{i:02d}_par"] = np.array(list(c["tsz"]) + list(c["org"]) + [c["vel"]], np.int32)
        out[f"c{i:02d}_T"] = np.float64(c["T"])
        out[f"c{i:02d}_cb"] = np.array(c["cb"], np.float64).reshape(-1, 3)
        out[f"c{i:02d}_cs"] = np.array(c["cs"], np.float64).reshape(-1, 3)
        out[f"c{i:02d}_n"] = np.array([len(p) for p in r], np.int32)
        out[f"c{i:02d}_y"] = np.array([y for p in r for y, _ in p], np.int32)
        out[f"c{i:02d}_s"] = np.array([s for p in r for _, s in p], np.float64)
        n_match += sum(len(p) for p in r)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_pairing.npz")
    np.savez_compressed(path, **out)
    print("cases", len(out) // 8, "side matches", n_match, "bytes", os.path.getsize(path))


if __name__ == "__main__":
    main()
