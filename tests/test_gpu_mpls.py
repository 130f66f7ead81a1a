"""GPU parity tests of MPLS on the DESC incidence (Algorithms/MPLS.m:152-256) through the C ABI:
the CEMP+MST initialisation (Boruvka spanning tree + propagation along the tree) and the MPLS reweighting loop
(Weighted_LAA + cycle reweighting of the residuals), against the CPU oracle and the golden fixtures.

Tolerances: rotations 1e-6 deg mean angular error (north_star); `score` values 1e-8; identical iteration counts."""
import numpy as np
import pytest

import desc_b200
from conftest import cemp_golden_names, load_golden
from gpu_util import ROT_TOL_DEG
from oracle import desc_oracle as O

pytestmark = pytest.mark.gpu

DEMO_MPLS = dict(stop_threshold=1e-3, max_iter=100, reweighting=[32.0], thresholding=[0.95, 0.9, 0.85, 0.8],
                 cycle_info_ratio=1.0 / (np.arange(1, 101) + 1))             # compare_algorithms.m:36-40
DEMO_CEMP = dict(max_iter=6, reweighting=2.0 ** np.arange(6), nsample=50)    # compare_algorithms.m:29-32


@pytest.mark.parametrize("name", cemp_golden_names())
def test_mpls_golden_fixture(name):
    g = load_golden(name)
    MP = dict(stop_threshold=float(g["mpls_stop_threshold"]), max_iter=int(g["mpls_max_iter"]),
              reweighting=g["mpls_reweighting"], thresholding=g["mpls_thresholding"],
              cycle_info_ratio=g["mpls_cycle_info_ratio"])
    with desc_b200.Solver(g["Ind"], g["RijMat"]) as s:
        s.build_incidence(n_sample=int(g["nsample"]), cycles=(g["cyc_ptr"], g["cyc_apex"]))
        s.cycle_inconsistency()
        # (a) each stage from the fixture's own inputs
        R_mst = s.mst_init(g["SVec"])
        np.testing.assert_allclose(R_mst, g["R_mst"], atol=1e-11, rtol=0)   # same tree, same products
        R, scores = s.mpls_refine(MP, SVec=g["SVec"], R_init=g["R_mst"])
        assert scores.size == g["mpls_scores"].size
        np.testing.assert_allclose(scores, g["mpls_scores"], rtol=1e-8, atol=1e-12)
        assert O.aligned_angle_deg(R, g["R_mpls"]).mean() <= ROT_TOL_DEG
        # (b) chained on the device: cemp -> mst_init -> mpls_refine
        s.cemp(int(g["max_iter"]), g["reweighting"])
        R_mst2 = s.mst_init()
        np.testing.assert_allclose(R_mst2, g["R_mst"], atol=1e-10, rtol=0)
        R2, scores2 = s.mpls_refine(MP)
        assert scores2.size == g["mpls_scores"].size
        assert O.aligned_angle_deg(R2, g["R_mpls"]).mean() <= ROT_TOL_DEG
        tm = s.timings()
        assert tm["mst_ms"] > 0.0 and tm["laa_iters"] == scores2.size


@pytest.mark.parametrize("case", [(100, 0.5, 0.2, 0.1, 1), (200, 0.3, 0.3, 0.05, 2), (150, 0.6, 0.1, 0.2, 3)])
def test_mpls_entry_point_matches_oracle(case):
    """`[R_est, R_init] = MPLS(Ind, RijMat, CEMP_parameters, MPLS_parameters)` with the demo's parameters"""
    n, p, q, sigma, seed = case
    mo = O.uniform_topology(n, p, q, sigma, "uniform", rng=200 + seed)
    oR, oR0, info = O.MPLS(mo["Ind"], mo["RijMat"], DEMO_CEMP, DEMO_MPLS, seed=seed, return_info=True)
    R, R0 = desc_b200.MPLS(mo["Ind"], mo["RijMat"], dict(DEMO_CEMP, seed=seed), DEMO_MPLS)
    assert R.shape == (3, 3, n) and R0.shape == (3, 3, n)
    np.testing.assert_allclose(R0, oR0, atol=1e-10, rtol=0)
    assert O.aligned_angle_deg(R, oR).mean() <= ROT_TOL_DEG
    # the refinement beats its initialisation (what compare_algorithms.m:77-78 tabulates)
    _, _, e0, _ = desc_b200.Rotation_Alignment(R0, mo["R_orig"])
    _, _, e1, _ = desc_b200.Rotation_Alignment(R, mo["R_orig"])
    assert e1 < e0
    with desc_b200.Solver(mo["Ind"], mo["RijMat"]) as s:
        s.build_incidence(n_sample=50, seed=seed)
        s.cycle_inconsistency()
        s.cemp(6, DEMO_CEMP["reweighting"])
        s.mst_init()
        _, scores = s.mpls_refine(DEMO_MPLS)
    assert scores.size == info["iterations"]
    np.testing.assert_allclose(scores, info["scores"], rtol=1e-7, atol=1e-11)


def test_mpls_schedules_and_stop_rules():
    """non-default schedules (compare_algorithms.m:47-48 variant): growing beta, cycle information weighted more
    over time; max_iter caps the loop at max_iter-1 iterations (MPLS.m:219)"""
    mo = O.uniform_topology(120, 0.5, 0.3, 0.1, "uniform", rng=301)
    CP = dict(max_iter=4, reweighting=[1.0, 4.0], nsample=20)
    for MP in (dict(stop_threshold=1e-3, max_iter=100, reweighting=0.1 * 1.5 ** np.arange(15), thresholding=[0.9],
                    cycle_info_ratio=1.0 - 1.0 / (np.arange(1, 101) + 1)),
               dict(stop_threshold=0.0, max_iter=4, reweighting=[8.0], thresholding=[0.95, 0.8], cycle_info_ratio=[0.5])):
        oR, oR0, info = O.MPLS(mo["Ind"], mo["RijMat"], CP, MP, seed=5, return_info=True)
        with desc_b200.Solver(mo["Ind"], mo["RijMat"]) as s:
            s.build_incidence(n_sample=20, seed=5)
            s.cycle_inconsistency()
            s.cemp(4, CP["reweighting"])
            R0 = s.mst_init()
            R, scores = s.mpls_refine(MP)
        assert scores.size == info["iterations"]
        np.testing.assert_allclose(scores, info["scores"], rtol=1e-7, atol=1e-11)
        assert O.aligned_angle_deg(R, oR).mean() <= ROT_TOL_DEG
    assert scores.size == 3    # max_iter = 4 with an unreachable threshold: 3 iterations


def test_mst_init_edge_cases():
    # clean graph: every tree gives the exact rotations; edges without cycles (SVec = 1) tie and are broken by index
    mo = O.uniform_topology(80, 0.15, 0.0, 0.0, rng=302)
    with desc_b200.Solver(mo["Ind"], mo["RijMat"]) as s:
        s.build_incidence(n_sample=10, seed=0)
        s.cycle_inconsistency()
        SVec = s.cemp(3, [1.0])
        R0 = s.mst_init()
        oR0, tree = O.mst_init(mo["Ind"], mo["RijMat"], SVec)
        np.testing.assert_allclose(R0, oR0, atol=1e-11, rtol=0)
        assert O.aligned_angle_deg(R0, mo["R_orig"]).max() < 1e-5
        np.testing.assert_allclose(R0[:, :, 0], np.eye(3), atol=0)          # MPLS.m:166
        # all weights equal: pure index tie-breaking
        R1 = s.mst_init(np.ones(mo["Ind"].shape[0]))
        oR1, _ = O.mst_init(mo["Ind"], mo["RijMat"], np.ones(mo["Ind"].shape[0]))
        np.testing.assert_allclose(R1, oR1, atol=1e-11, rtol=0)
        with pytest.raises(desc_b200.DescError):
            s.mpls_refine(dict(DEMO_MPLS, max_iter=0))
    # a path graph: the tree is the graph (depth n-1 exercises the level loop)
    n = 300
    Ind = np.stack([np.arange(1, n), np.arange(2, n + 1)], axis=1).astype(np.float64)
    Rn = O._rand_rot(n, np.random.default_rng(3))
    Rij = O.to_matlab(Rn[:-1] @ Rn[1:].transpose(0, 2, 1))
    with desc_b200.Solver(Ind, Rij) as s:
        R0 = s.mst_init(np.random.default_rng(4).random(n - 1))
    assert O.aligned_angle_deg(R0, O.to_matlab(Rn)).max() < 1e-5
    # a disconnected graph is rejected (the reference's loop MPLS.m:171 would never end)
    Ind = np.array([[1.0, 2.0], [3.0, 4.0]])
    with desc_b200.Solver(Ind, O.to_matlab(O._rand_rot(2, np.random.default_rng(0)))) as s:
        with pytest.raises(desc_b200.DescError) as e:
            s.mst_init(np.ones(2))
        assert e.value.code == desc_b200._lib.ERR_ARG
        with pytest.raises(desc_b200.DescError):
            s.mst_init()          # no cemp on this handle


@pytest.mark.parametrize("case", [(100, 0.5, 0.1, 0.1, 1), (250, 0.2, 0.3, 0.05, 2)])
def test_spectral_matches_oracle(case):
    """`R_est = Spectral(Ind, RijMat)` (Algorithms/Spectral.m): GCW's solver on the plain block matrix"""
    n, p, q, sigma, seed = case
    mo = O.uniform_topology(n, p, q, sigma, "uniform", rng=400 + seed)
    R = desc_b200.Spectral(mo["Ind"], mo["RijMat"])
    assert O.aligned_angle_deg(R, O.spectral(mo["Ind"], mo["RijMat"])).mean() <= ROT_TOL_DEG
    # a GCW call on the same handle afterwards is unaffected by the weight rule
    with desc_b200.Solver(mo["Ind"], mo["RijMat"]) as s:
        s.spectral()
        Rg = s.gcw(mo["ErrVec"])
    assert O.aligned_angle_deg(Rg, O.gcw(mo["Ind"], mo["RijMat"], mo["ErrVec"])).mean() <= ROT_TOL_DEG


def test_demo_compare_algorithms_table():
    """demo/compare_algorithms.py = Demo/compare_algorithms.m on the device: every row is produced, the robust
    methods beat the plain spectral one on a corrupted graph and the refinement does not hurt"""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("compare_algorithms", os.path.join(root, "demo", "compare_algorithms.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rows, extra = mod.run(n=100, p=0.5, q=0.2, sigma=0.1, seed=0)
    err = {name: mean for name, mean, _ in rows}
    assert list(err) == ["Spectral", "CEMP+MST", "CEMP+GCW", "MPLS", "DESC_init", "DESC"]
    assert all(np.isfinite(v) and 0.0 <= v < 30.0 for v in err.values())
    assert err["MPLS"] < err["CEMP+MST"] and err["DESC"] < err["Spectral"] and err["CEMP+GCW"] < err["Spectral"]
    assert err["DESC"] <= err["DESC_init"] + 0.5
    assert float(np.mean(np.abs(extra["S_vec"].ravel() - extra["ErrVec"].ravel()))) < 0.05
