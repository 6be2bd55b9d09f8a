"""Cost builders of the host tracker: LocoMouse::unaryCostBox / pairwisePotential + MATSPARSE
(LocoMouse_class.cpp:1909-2070, MyMat.cpp:141-178).
CPU: the oracle's restatement equals the REFERENCE'S OWN CODE (compiled from /root/reference with its own MyMat.cpp by
`make -C oracle ref`) bit for bit on committed vectors and on fresh random inputs.
GPU (-m gpu): the kernels behind lm_unary_costs / lm_pairwise_costs equal the oracle bit for bit on every frame."""
import os

import numpy as np
import pytest

from locomouse_cpp_b200.types import Results, location_priors, pairwise_params
from oracle import reference_nms as ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_costs.npz")
PRIOR_ROWS = [(0.8, 0.25, 0.5, 0.4, 1.0, 0.0, 0.5), (0.8, 0.75, 0.5, 0.4, 1.0, 0.5, 1.0), (0.3, 0.25, 0.4, 0.0, 0.6, 0.0, 0.5),
              (0.3, 0.75, 0.35, 0.0, 0.6, 0.5, 1.0)]
BW, BH = 400, 235


def _cands(rng, n, cx=None, cy=None):
    out = []
    for _ in range(n):
        if cx is not None and rng.integers(0, 2):
            x, y = int(np.clip(cx + rng.integers(-25, 26), 0, BW - 1)), int(np.clip(cy + rng.integers(-25, 26), 0, BH - 1))
        else:
            x, y = int(rng.integers(0, BW)), int(rng.integers(0, BH))
        out.append((x, y, float(np.float32(rng.uniform(0.01, 3)))))
    return out


def golden_cases(seed=4242, n=48):
    rng = np.random.Generator(np.random.PCG64(seed))
    cases = []
    for it in range(n):
        ni, nj = int(rng.integers(0, 10)), int(rng.integers(0, 10))
        if it == 0:
            ni, nj = 0, 5       # empty C_i: the "ONG -> X(i+1)" entries are never written (quirk)
        if it == 1:
            ni, nj = 4, 0
        cx, cy = int(rng.integers(0, BW)), int(rng.integers(0, BH))
        alpha = float(rng.choice([0.1, 0.1, 0.0, 100.0]))
        occ = float(rng.choice([1e-2, 1e-2, 0.0]))
        cases.append(dict(ci=_cands(rng, ni, cx, cy), cj=_cands(rng, nj, cx, cy), alpha=alpha, occ=occ))
    return cases


def _same_sparse(a, b):
    return (a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
            and np.array_equal(np.ascontiguousarray(a[4]).view(np.uint64), np.ascontiguousarray(b[4]).view(np.uint64)))


def _ref_pairwise(c, P):
    return ref.pairwise_potential(c["ci"], c["cj"], P.grid_x, P.grid_y, P.grid_spacing, P.ong_w, P.ong_h, P.max_displacement, P.alpha_vel,
                                  P.occluded_cost)


def make_golden():
    out = {}
    for i, c in enumerate(golden_cases()):
        P = pairwise_params(BW, BH, alpha_vel=c["alpha"], occluded_cost=c["occ"])
        nr, nc, jc, ir, pr = _ref_pairwise(c, P)
        out[f"c{i:02d}_dims"] = np.array([nr, nc], np.int32)
        out[f"c{i:02d}_jc"], out[f"c{i:02d}_ir"], out[f"c{i:02d}_pr"] = jc, ir, pr
        out[f"c{i:02d}_unary"] = ref.unary_cost_box(c["cj"], BW, BH, PRIOR_ROWS)
    np.savez_compressed(GOLD, **out)
    return len(out) // 5


def test_oracle_cost_builders_match_reference_golden(oracle):
    z = np.load(GOLD)
    pri = location_priors(PRIOR_ROWS)
    nnz = nz_unary = 0
    for i, c in enumerate(golden_cases()):
        P = pairwise_params(BW, BH, alpha_vel=c["alpha"], occluded_cost=c["occ"])
        got = oracle.pairwise_potential(c["ci"], c["cj"], P)
        want = (int(z[f"c{i:02d}_dims"][0]), int(z[f"c{i:02d}_dims"][1]), z[f"c{i:02d}_jc"], z[f"c{i:02d}_ir"], z[f"c{i:02d}_pr"])
        assert _same_sparse(got, want), f"pairwisePotential differs from the reference on case {i}"
        u = oracle.unary_cost_box(c["cj"], BW, BH, pri)
        assert u.shape == z[f"c{i:02d}_unary"].shape and np.array_equal(u.view(np.uint64), z[f"c{i:02d}_unary"].view(np.uint64)), f"unaryCostBox, case {i}"
        nnz += len(got[3])
        nz_unary += int((u != 0).sum())
    assert nnz > 3000 and nz_unary > 100
    # quirk: with no candidates in frame i the ONG -> X(i+1) block stays empty: only the Nong diagonal entries remain
    assert len(z["c00_ir"]) == pairwise_params(BW, BH).ong_w * pairwise_params(BW, BH).ong_h


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref_nms.so not built (reference not mounted)")
def test_cost_golden_is_what_the_reference_code_produces_and_fresh_inputs_agree(oracle):
    z = np.load(GOLD)
    for i, c in enumerate(golden_cases()):
        P = pairwise_params(BW, BH, alpha_vel=c["alpha"], occluded_cost=c["occ"])
        want = (int(z[f"c{i:02d}_dims"][0]), int(z[f"c{i:02d}_dims"][1]), z[f"c{i:02d}_jc"], z[f"c{i:02d}_ir"], z[f"c{i:02d}_pr"])
        assert _same_sparse(_ref_pairwise(c, P), want)
        assert np.array_equal(ref.unary_cost_box(c["cj"], BW, BH, PRIOR_ROWS), z[f"c{i:02d}_unary"])
    pri = location_priors(PRIOR_ROWS)
    for c in golden_cases(seed=99, n=150):
        P = pairwise_params(BW, BH, alpha_vel=c["alpha"], occluded_cost=c["occ"])
        assert _same_sparse(oracle.pairwise_potential(c["ci"], c["cj"], P), _ref_pairwise(c, P))
        a, b = oracle.unary_cost_box(c["ci"], BW, BH, pri), ref.unary_cost_box(c["ci"], BW, BH, PRIOR_ROWS)
        assert a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def _random_results(rng, n, cap=16):
    res = Results(n, cap, 4 * cap)
    for f in range(n):
        cx, cy = int(rng.integers(0, BW)), int(rng.integers(0, BH))
        for feat in range(2):
            k = int(rng.integers(0, cap + 1)) if f % 7 else 0
            res.n_bottom[f, feat] = k
            for i, (x, y, s) in enumerate(_cands(rng, k, cx, cy)):
                res.bottom[f, feat, i] = (x, y, s)
    return res


@pytest.mark.gpu
@pytest.mark.parametrize("alpha,occ", [(0.1, 1e-2), (0.0, 1e-2), (100.0, 0.0)])
def test_gpu_cost_builders_equal_oracle(oracle, alpha, occ):
    from locomouse_cpp_b200 import synth
    from locomouse_cpp_b200.api import Detector

    rng = np.random.Generator(np.random.PCG64(77))
    n = 300
    res = _random_results(rng, n)
    cfg, model, bkg, calib, *_ = synth.make_problem(synth.SynthSpec(), 1, seed=1000)
    det = Detector(cfg, model, bkg, calib, device=0)
    pri = location_priors(PRIOR_ROWS)
    P = pairwise_params(BW, BH, alpha_vel=alpha, occluded_cost=occ)
    for feat in range(2):
        U = det.unary_costs(res, feat, BW, BH, pri)
        offs, jc, ir, pr = det.pairwise_costs(res, feat, P, cap=16)   # too small on purpose: the binding retries with the reported size
        assert offs[0] == 0 and offs[1] == 0 and not jc[0].any()
        for f in range(n):
            k = int(res.n_bottom[f, feat])
            cands = [tuple(c) for c in res.bottom[f, feat, :k].tolist()]
            want = oracle.unary_cost_box(cands, BW, BH, pri)                 # (k, n_priors)
            assert np.array_equal(U[f, :, :k].T.copy().view(np.uint64), want.view(np.uint64)) and not U[f, :, k:].any(), (feat, f)
            if f == 0:
                continue
            kp = int(res.n_bottom[f - 1, feat])
            prev = [tuple(c) for c in res.bottom[f - 1, feat, :kp].tolist()]
            nr, nc, wjc, wir, wpr = oracle.pairwise_potential(prev, cands, P)
            assert np.array_equal(jc[f, :nc + 1], wjc) and (jc[f, nc + 1:] == wjc[-1]).all(), (feat, f)
            assert offs[f + 1] - offs[f] == len(wir)
            assert np.array_equal(ir[offs[f]:offs[f + 1]], wir), (feat, f)
            assert np.array_equal(pr[offs[f]:offs[f + 1]].view(np.uint64), wpr.view(np.uint64)), (feat, f)
    det.close()


if __name__ == "__main__":
    print("cases", make_golden())
