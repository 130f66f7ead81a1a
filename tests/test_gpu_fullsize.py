"""Full-size runs of the BASELINE.json configurations the oracle cannot finish in seconds: checked
through size-independent properties of the algorithm (DESC.m:148-261, GCW.m) instead of element-wise
against the oracle (that is done at reduced sizes in test_gpu_parity.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import desc_b200                                   # noqa: E402
from desc_b200 import synth                        # noqa: E402
from oracle import desc_oracle as O                # noqa: E402  (metric only)


def _solve(mo, n, iters, rule, n_sample, want_w=True, gcw=True):
    Ind_d, R_d = mo["Ind"].reshape(-1).contiguous(), mo["RijMat"].reshape(-1).contiguous()
    out = {}
    with desc_b200.Solver(Ind_d, R_d, n=n) as s:
        out["info"] = s.build_incidence(n_sample=n_sample, seed=1)
        out["codeg"] = s.codeg()
        s.cycle_inconsistency()
        out["S"], out["hist"], out["iters_run"] = s.pgd(iters, rule)
        if want_w:
            out["rowptr"], _ = s.incidence()
            out["w"] = s.w()
            out["S0"] = s.S0()
        if gcw:
            out["R"] = s.gcw()
        out["timings"] = s.timings()
    return out


def _check_properties(out, mo, n_sample_rule, rot_tol_deg):
    info, S, hist = out["info"], out["S"], out["hist"]
    ns = info["n_sample"]
    assert ns == n_sample_rule
    # CSR budget (DESC.m:43-51): every edge keeps min(codeg, n_sample) cycles
    assert info["m_cycle"] == int(np.minimum(out["codeg"].astype(np.int64), ns).sum())
    assert info["m_pos"] == int((out["codeg"] > 0).sum())
    # simplex constraints of every edge (DESC.m:213-224) and S = w . S0 (DESC.m:229)
    rp, w, S0 = out["rowptr"], out["w"], out["S0"]
    assert w.min() >= 0.0
    nz = np.flatnonzero(np.diff(rp) > 0)
    sums = np.add.reduceat(w, rp[nz])
    assert np.max(np.abs(sums - 1.0)) <= 1e-12
    sw = np.add.reduceat(w * S0, rp[nz])
    assert np.max(np.abs(sw - S[nz])) <= 1e-12
    assert np.all(S[np.diff(rp) == 0] == 1.0)                       # DESC.m:148
    assert S.min() >= 0.0 and S.max() <= 1.0 + 1e-12
    # the objective decreases monotonically (projected gradient on a quadratic, small step)
    obj = hist[:out["iters_run"], 1]
    assert np.all(np.diff(obj) <= 1e-9 * np.abs(obj[:-1]))
    # corrupted edges are separated from clean ones, and the rotations are recovered
    corr = mo["corrupted"].cpu().numpy()
    assert S[corr].mean() > 3.0 * S[~corr].mean()
    Ro = mo["R_orig"].cpu().numpy().transpose(2, 1, 0)
    assert O.rotation_alignment(out["R"], Ro)[2] <= rot_tol_deg


def test_config4_n10000_p01_properties_and_reproducibility():
    """configs[3]: Uniform_Topology n=10000 p=0.1 (5.0e6 edges, 1.5e8 sampled 3-cycles), 12 iterations"""
    mo = synth.uniform_topology(10000, 0.1, 0.2, 0.1, "uniform", seed=0, device="cuda")
    a = _solve(mo, 10000, 12, desc_b200.ConstantStepSize(0.01), 0)
    assert a["info"]["m"] == mo["m"] and a["info"]["m_cycle"] > 1.4e8
    _check_properties(a, mo, 30, 1.0)
    del a["w"], a["S0"]
    b = _solve(mo, 10000, 12, desc_b200.ConstantStepSize(0.01), 0, want_w=False)
    np.testing.assert_array_equal(a["S"], b["S"])                   # no atomics: bit-reproducible
    np.testing.assert_array_equal(a["hist"], b["hist"])


def test_config5_sfm_shaped_n50000_deg100_50_cycles_per_edge():
    """configs[4]: locally clustered graph, n=50000, mean degree 100, 50 sampled cycles per edge, the
    reference's large-scale settings ConstantStepSize(1), iters=30 (compare_algorithms.m:2-5)"""
    mo = synth.ring_topology(50000, 100, 75, 0.2, 0.05, seed=2, device="cuda")
    assert 2.3e6 < mo["m"] < 2.7e6
    a = _solve(mo, 50000, 30, desc_b200.ConstantStepSize(1.0), 50, gcw=False)
    assert a["info"]["max_slots_per_edge"] == 50 and a["info"]["m_cycle"] > 0.8e8
    info, S = a["info"], a["S"]
    rp, w, S0 = a["rowptr"], a["w"], a["S0"]
    nz = np.flatnonzero(np.diff(rp) > 0)
    assert w.min() >= 0.0 and np.max(np.abs(np.add.reduceat(w, rp[nz]) - 1.0)) <= 1e-12
    assert np.max(np.abs(np.add.reduceat(w * S0, rp[nz]) - S[nz])) <= 1e-12
    corr = mo["corrupted"].cpu().numpy()
    assert S[corr].mean() > 3.0 * S[~corr].mean()
    assert info["m_cycle"] == int(np.minimum(a["codeg"].astype(np.int64), 50).sum())


@pytest.mark.parametrize("case", [("configs[1]: n=1000 p=0.5 q=0.3 sigma=0.1", 1000, 0.5, 0.3, 0.1, 63, 30),
                                  ("n=2000 p=0.5 (the size / n_sample regime of configs[2])", 2000, 0.5, 0.25, 0.1, 125, 12)])
def test_dense_graphs_full_size_properties(case):
    """longer slot lists (n_sample 63 and 125: 8 / 16 lanes per edge in the streamed kernel) at full size"""
    _, n, p, q, sigma, ns_rule, iters = case
    mo = synth.uniform_topology(n, p, q, sigma, "uniform", seed=3, device="cuda")
    a = _solve(mo, n, iters, desc_b200.ConstantStepSize(0.01), 0)
    assert abs(a["info"]["n_sample"] - ns_rule) <= 1
    _check_properties(a, mo, a["info"]["n_sample"], 2.0)
