"""The C/OpenMP restatement (oracle/desc_full.c + oracle/desc_pgd.c) against the numpy oracle.

The C port is what the full-size GPU parity tests and bench.py's CPU arm use (the numpy oracle cannot hold 1.5e8 slots);
it is pinned here, stage by stage, on random graphs and on the committed golden fixtures: integer work bit-exact,
d_ijk to 2 ulp (libm vs numpy acos), S_vec / w to 1e-13, rotations to 1e-9 deg.  CPU only."""
import numpy as np
import pytest
from conftest import golden_names, golden_rule, load_golden

from oracle import desc_oracle as O
from oracle import desc_oracle_c as OC

INT_FIELDS = ("codeg", "pos_edges", "rowptr", "e_ij", "e_jk", "e_ki", "k", "IKJ", "JKI")


def _same_incidence(a, b):
    assert (a.n, a.m, a.n_sample, a.m_pos, a.m_cycle) == (b.n, b.m, b.n_sample, b.m_pos, b.m_cycle)
    for f in INT_FIELDS:
        np.testing.assert_array_equal(np.asarray(getattr(a, f)), np.asarray(getattr(b, f)), err_msg=f)


@pytest.mark.parametrize("case", [(60, 0.5, None, 0), (120, 0.5, None, 3), (200, 0.3, 40, 5), (90, 0.6, -1, 2),
                                  (150, 0.1, None, 7), (80, 0.12, None, 9)])
def test_c_port_matches_numpy_oracle_stage_by_stage(case):
    n, p, ns, seed = case
    mo = O.uniform_topology(n, p, 0.2, 0.1, rng=seed)
    a = O.build_incidence(mo["Ind"], n_sample=ns, seed=seed)
    b = OC.build_incidence(mo["Ind"], n_sample=ns, seed=seed)
    _same_incidence(a, b)
    Sa, Sb = O.cycle_inconsistency(a, mo["RijMat"]), OC.cycle_inconsistency(b, mo["RijMat"])
    assert float(np.max(np.abs(Sa - Sb), initial=0.0)) <= 4.5e-16
    for rule in (lambda: O.ConstantStepSize(0.01), lambda: O.PiecewiseStepSize(0.05, 7)):
        Pa = O.pgd(a, Sa, 40, rule(), return_w=True)
        Pb = OC.pgd(b, Sa, 40, rule(), return_w=True)          # same S0: isolates the loop
        assert Pa[2] == Pb[2]
        np.testing.assert_allclose(Pb[0], Pa[0], rtol=0, atol=1e-13)
        np.testing.assert_allclose(Pb[3], Pa[3], rtol=0, atol=1e-13)
        np.testing.assert_allclose(Pb[1], Pa[1], rtol=1e-12, atol=1e-15)
    Ra, Rb = O.gcw(mo["Ind"], mo["RijMat"], Pa[0]), OC.gcw(mo["Ind"], mo["RijMat"], Pa[0])
    assert O.aligned_angle_deg(Ra, Rb).mean() <= 1e-9


def test_c_port_gcw_arpack_branch_matches_dense():
    """graphs above the dense limit go through ARPACK on the threaded block operator (what GCW.m:27 does)"""
    mo = O.uniform_topology(520, 0.06, 0.2, 0.1, rng=4)
    S = mo["ErrVec"]
    Ra, Rb = O.gcw(mo["Ind"], mo["RijMat"], S), OC.gcw(mo["Ind"], mo["RijMat"], S)
    assert O.aligned_angle_deg(Ra, Rb).mean() <= 1e-8
    for power in (1, None):
        Ra, Rb = O.gcw(mo["Ind"], mo["RijMat"], S, power=power), OC.gcw(mo["Ind"], mo["RijMat"], S, power=power)
        assert O.aligned_angle_deg(Ra, Rb).mean() <= 1e-8


@pytest.mark.parametrize("name", golden_names())
def test_c_port_reproduces_the_golden_fixtures(name):
    """the committed DESC fixtures (literal dense restatement): incidence bit-exact, S0 to 2 ulp, PGD, GCW"""
    g = load_golden(name)
    ns = int(g["n_sample_arg"])
    inc = OC.build_incidence(g["Ind"], n_sample=None if ns < 0 else ns, seed=int(g["sampler_seed"]))
    assert inc.n_sample == int(g["n_sample"])
    np.testing.assert_array_equal(inc.pos_edges + 1, g["CoDeg_pos_ind"])
    np.testing.assert_array_equal(inc.rowptr, g["cum_ind"])
    np.testing.assert_array_equal(inc.k + 1, g["IJK"])
    np.testing.assert_array_equal(inc.e_jk + 1, g["Ind_jk"])
    np.testing.assert_array_equal(inc.e_ki + 1, g["Ind_ki"])
    np.testing.assert_array_equal(inc.IKJ + 1, g["IKJ"])
    np.testing.assert_array_equal(inc.JKI + 1, g["JKI"])
    S0 = OC.cycle_inconsistency(inc, g["RijMat"])
    assert float(np.max(np.abs(S0 - g["S0_long"]), initial=0.0)) <= 4.5e-16
    if str(g["rule_kind"]) in ("const", "piecewise"):
        S_vec, hist, iters_run, w = OC.pgd(inc, S0, int(g["iters"]), golden_rule(g, O), return_w=True)
        assert iters_run == int(g["iters_run"])
        np.testing.assert_allclose(S_vec, g["S_vec"], rtol=1e-11, atol=1e-14)
        np.testing.assert_allclose(w, g["wijk"], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(hist, g["hist"], rtol=1e-10, atol=1e-13)
    R = OC.gcw(g["Ind"], g["RijMat"], g["S_vec"])
    assert O.aligned_angle_deg(R, g["R_est"]).mean() < 1e-6


def test_full_solve_and_stage_timings():
    mo = O.uniform_topology(300, 0.4, 0.2, 0.1, rng=11)
    r = OC.DESC_init(mo["Ind"], mo["RijMat"], dict(iters=25, Gradient=O.ConstantStepSize(0.01)), seed=2, full=True)
    oR, oS = O.DESC_init(mo["Ind"], mo["RijMat"], dict(iters=25, Gradient=O.ConstantStepSize(0.01)), seed=2)
    np.testing.assert_allclose(r["S_vec"], oS, rtol=0, atol=1e-13)
    assert O.aligned_angle_deg(r["R"], oR).mean() <= 1e-9
    assert set(r["timings"]) == {"graph_s", "build_s", "cycle_s", "pgd_s", "gcw_s"}
    assert OC.host_threads() >= 1
