"""GPU parity tests of SURVEY 8(f) #3/#4 through the C ABI: CEMP / CEMP+GCW on the DESC incidence
(Algorithms/CEMP.m, CEMP_GCW.m), the alignment metric (Utils/Rotation_Alignment.m) and the make_plots
diagnostics branch of the PGD loop (DESC.m:235-239), each against the CPU oracle / golden fixtures.

Tolerances: corruption estimates 1e-10 relative (north_star), rotations 1e-6 deg mean angular error."""
import numpy as np
import pytest

import desc_b200
from conftest import cemp_golden_names, load_golden
from gpu_util import rel_err, RTOL, ROT_TOL_DEG
from oracle import desc_oracle as O

pytestmark = pytest.mark.gpu


def _cemp_cuda(Ind, RijMat, P, seed=0, cycles=None, gcw=True):
    with desc_b200.Solver(Ind, RijMat) as s:
        info = s.build_incidence(n_sample=int(P["nsample"]), seed=seed, cycles=cycles)
        s.cycle_inconsistency()
        S0 = s.S0()
        SVec = s.cemp(P["max_iter"], P["reweighting"])
        R = s.cemp_gcw() if gcw else None
        tm = s.timings()
    return dict(info=info, S0=S0, SVec=SVec, R=R, timings=tm)


@pytest.mark.parametrize("name", cemp_golden_names())
def test_cemp_golden_fixture_with_replacement_draw(name):
    """the fixture's CoIndMat (CEMP.m:63 draws WITH replacement: repeated apices, unsorted) as explicit lists"""
    g = load_golden(name)
    P = dict(max_iter=int(g["max_iter"]), reweighting=g["reweighting"], nsample=int(g["nsample"]))
    c = _cemp_cuda(g["Ind"], g["RijMat"], P, cycles=(g["cyc_ptr"], g["cyc_apex"]))
    ns = int(g["nsample"])
    pos = np.diff(g["cyc_ptr"]) > 0
    assert rel_err(c["S0"].reshape(-1, ns).T, g["S0Mat"][:, pos], floor=1e-30) <= RTOL
    assert rel_err(c["SVec"], g["SVec"], floor=1e-12) <= RTOL
    assert (c["SVec"][~pos] == 1.0).all()
    assert O.aligned_angle_deg(c["R"], g["R_est"]).mean() <= ROT_TOL_DEG
    # every intermediate SVec_t of the fixture: run t reweightings
    for t in (0, 1, int(g["max_iter"]) - 1):
        Pt = dict(P, max_iter=t)
        ct = _cemp_cuda(g["Ind"], g["RijMat"], Pt, cycles=(g["cyc_ptr"], g["cyc_apex"]), gcw=False)
        assert rel_err(ct["SVec"], g["SVec_hist"][t], floor=1e-12) <= RTOL
    # reference-style entry points
    Pe = dict(P, cycles=(g["cyc_ptr"], g["cyc_apex"]))
    SVec = desc_b200.CEMP(g["Ind"], g["RijMat"], Pe)
    assert SVec.shape == (1, g["Ind"].shape[0]) and rel_err(SVec.ravel(), g["SVec"], floor=1e-12) <= RTOL
    R = desc_b200.CEMP_GCW(g["Ind"], g["RijMat"], Pe)
    assert R.shape == g["R_est"].shape and O.aligned_angle_deg(R, g["R_est"]).mean() <= ROT_TOL_DEG
    # the metric the demo prints (compare_algorithms.m:77-84)
    _, _, mean_err, med_err = desc_b200.Rotation_Alignment(R, g["R_orig"])
    assert abs(mean_err - float(g["mean_error"])) <= 1e-5 and abs(med_err - float(g["median_error"])) <= 1e-5


@pytest.mark.parametrize("case", [(200, 0.5, 0.2, 0.0, 50, 1), (300, 0.3, 0.3, 0.1, 30, 2), (150, 0.6, 0.25, 0.05, 100, 3)])
def test_cemp_device_sampler_matches_oracle(case):
    """demo parameters (compare_algorithms.m:29-33) on the shared counter-based sampler"""
    n, p, q, sigma, nsample, seed = case
    mo = O.uniform_topology(n, p, q, sigma, "uniform", rng=100 + seed)
    P = dict(max_iter=6, reweighting=2.0 ** np.arange(6), nsample=nsample)
    c = _cemp_cuda(mo["Ind"], mo["RijMat"], P, seed=seed)
    inc = O.cemp_incidence(mo["Ind"], nsample, seed=seed)
    S0 = O.cycle_inconsistency(inc, mo["RijMat"])
    assert c["info"]["m_cycle"] == inc.m_cycle
    assert rel_err(c["S0"], S0, floor=1e-30) <= RTOL
    SVec = O.cemp(inc, S0, 6, P["reweighting"])
    assert rel_err(c["SVec"], SVec, floor=1e-12) <= RTOL
    R = O.gcw(mo["Ind"], mo["RijMat"], SVec, power=1.0)
    assert O.aligned_angle_deg(c["R"], R).mean() <= ROT_TOL_DEG
    assert c["timings"]["cemp_iters"] == 6 and c["timings"]["cemp_ms"] > 0.0


def test_cemp_edge_cases():
    # (a) edges without a 3-cycle keep 1; (b) max_iter = 0 returns the plain mean; (c) short beta vector is padded
    mo = O.uniform_topology(60, 0.12, 0.2, 0.05, "uniform", rng=31)
    inc = O.cemp_incidence(mo["Ind"], 5, seed=4)
    assert inc.m_pos < inc.m
    S0 = O.cycle_inconsistency(inc, mo["RijMat"])
    for T, beta in ((0, [1.0]), (4, [1.0, 5.0]), (2, [1.0, 2.0, 4.0, 8.0])):
        c = _cemp_cuda(mo["Ind"], mo["RijMat"], dict(max_iter=T, reweighting=beta, nsample=5), seed=4, gcw=False)
        o = O.cemp(inc, S0, T, beta)
        assert rel_err(c["SVec"], o, floor=1e-12) <= RTOL
        has = np.zeros(inc.m, bool)
        has[inc.pos_edges] = True
        assert (c["SVec"][~has] == 1.0).all()
    # (d) triangle-free graph (a path): everything stays 1
    n = 12
    Ind = np.stack([np.arange(1, n), np.arange(2, n + 1)], axis=1).astype(np.float64)
    Rij = O.to_matlab(O._rand_rot(n - 1, np.random.default_rng(1)))
    c = _cemp_cuda(Ind, Rij, dict(max_iter=3, reweighting=[1.0], nsample=5), gcw=False)
    assert (c["SVec"] == 1.0).all()
    # (e) stage order is enforced
    with desc_b200.Solver(Ind, Rij) as s:
        with pytest.raises(desc_b200.DescError):
            s.cemp(1, [1.0])
        s.build_incidence(n_sample=5)
        s.cycle_inconsistency()
        with pytest.raises(desc_b200.DescError):
            s.cemp_gcw()          # no cemp yet
        with pytest.raises(ValueError):
            s.cemp(2, [])


def test_cycle_reweight_is_the_mpls_hvec_step():
    """MPLS.m:219-233: the CEMP reweighting applied to an arbitrary edge vector (the LAA residuals)"""
    mo = O.uniform_topology(120, 0.4, 0.2, 0.1, "uniform", rng=41)
    inc = O.cemp_incidence(mo["Ind"], 20, seed=7)
    S0 = O.cycle_inconsistency(inc, mo["RijMat"])
    x = np.random.default_rng(2).random(inc.m) * 0.3
    with desc_b200.Solver(mo["Ind"], mo["RijMat"]) as s:
        s.build_incidence(n_sample=20, seed=7)
        s.cycle_inconsistency()
        for beta, ev in ((32.0, 2.0 / 3.0), (0.0, 1.0), (5.0, 0.0)):
            got = s.cycle_reweight(x, beta, ev)
            assert rel_err(got, O.cemp_reweight(inc, S0, x, beta, empty_value=ev), floor=1e-12) <= RTOL


@pytest.mark.parametrize("n", [3, 7, 64, 1001])
def test_rotation_alignment_matches_oracle(n):
    rng = np.random.default_rng(n)
    R_gt = O.to_matlab(O._rand_rot(n, rng))
    Q = O._rand_rot(1, rng)[0]
    noisy = O.proj_so3(O.to_internal(R_gt) @ Q + 0.05 * rng.standard_normal((n, 3, 3)))
    R_est = O.to_matlab(noisy)
    R_out, R_align, mean_err, med_err = desc_b200.Rotation_Alignment(R_est, R_gt)
    oR_out, oR_align, o_mean, o_med = O.rotation_alignment(R_est, R_gt)
    np.testing.assert_allclose(R_align, oR_align, atol=1e-11)
    np.testing.assert_allclose(R_out, oR_out, atol=1e-11)
    assert abs(mean_err - o_mean) <= 1e-9 * max(1.0, o_mean) and abs(med_err - o_med) <= 1e-9 * max(1.0, o_med)
    # exact gauge: zero error up to the sqrt(eps) floor of acos at 1
    R_out, R_align, mean_err, med_err = desc_b200.Rotation_Alignment(O.to_matlab(O.to_internal(R_gt) @ Q), R_gt)
    np.testing.assert_allclose(R_align, Q.T, atol=1e-12)
    assert mean_err < 1e-5 and med_err < 1e-5


def test_make_plots_diagnostics_match_oracle(monkeypatch, tmp_path):
    """DESC.m:235-239: svec_errors, MSE_means, MSE_medians per iteration; the iterates are untouched"""
    mo = O.uniform_topology(150, 0.5, 0.25, 0.1, "uniform", rng=51)
    iters, lr, ns, seed = 12, 0.05, 20, 3
    inc = O.build_incidence(mo["Ind"], n_sample=ns, seed=seed)
    S0 = O.cycle_inconsistency(inc, mo["RijMat"])
    S_hist = []
    oS, ohist, orun = O.pgd(inc, S0, iters, O.ConstantStepSize(lr), S_hist=S_hist)
    od = O.pgd_diagnostics(mo["Ind"], mo["RijMat"], S_hist, mo["ErrVec"], mo["R_orig"])
    with desc_b200.Solver(mo["Ind"], mo["RijMat"]) as s:
        s.build_incidence(n_sample=ns, seed=seed)
        s.cycle_inconsistency()
        S_plain, hist_plain, run_plain = s.pgd(iters, desc_b200.ConstantStepSize(lr))
        S, hist, run, diag = s.pgd_diag(iters, desc_b200.ConstantStepSize(lr), mo["ErrVec"], mo["R_orig"])
    assert run == orun == run_plain and diag.shape == (run, 3)
    np.testing.assert_array_equal(S, S_plain)                       # diagnostics do not perturb the solve
    np.testing.assert_array_equal(hist, hist_plain)
    assert rel_err(S, oS, floor=1e-12) <= RTOL
    assert rel_err(diag[:, 0], od[:, 0], floor=1e-12) <= RTOL
    # alignment errors of two GCW solutions that agree to 1e-6 deg
    assert np.max(np.abs(diag[:, 1] - od[:, 1])) <= 1e-5 and np.max(np.abs(diag[:, 2] - od[:, 2])) <= 1e-5
    # through the reference-style entry point
    import desc_b200.solver as SV
    params = dict(iters=iters, Gradient=desc_b200.ConstantStepSize(lr), make_plots=True, ErrVec=mo["ErrVec"],
                  R_orig=mo["R_orig"], n_sample=ns, seed=seed)
    monkeypatch.chdir(tmp_path)       # DESC_init.m:261-262 appends the curves to CSV files in the working directory
    R, S2 = desc_b200.DESC_init(mo["Ind"], mo["RijMat"], params)
    np.testing.assert_array_equal(S2.ravel(), S)
    rot = np.loadtxt(tmp_path / "linear_convergence_rotation_error.csv", delimiter=",", ndmin=1)
    sv = np.loadtxt(tmp_path / "linear_convergence_svec_error.csv", delimiter=",", ndmin=1)
    assert rot.shape == (run,) and sv.shape == (run,)
    np.testing.assert_allclose(sv, diag[:, 0], rtol=1e-4)
    d = SV.last_diagnostics
    np.testing.assert_array_equal(d["svec_errors"], diag[:, 0])
    np.testing.assert_array_equal(d["obj_vals"], hist[:, 1])
    assert d["MSE_means"].shape == (run,) and d["MSE_medians"].shape == (run,)


def test_make_plots_with_early_stop():
    """sigma = 0: the objective stalls and the patience rule (DESC.m:243-256) stops the loop; the diagnostics
    have exactly iters_run rows"""
    mo = O.uniform_topology(60, 0.5, 0.0, 0.0, "uniform", rng=61)
    with desc_b200.Solver(mo["Ind"], mo["RijMat"]) as s:
        s.build_incidence(n_sample=10, seed=0)
        s.cycle_inconsistency()
        S_plain, hist_plain, run_plain = s.pgd(80, desc_b200.ConstantStepSize(0.01))
        S, hist, run, diag = s.pgd_diag(80, desc_b200.ConstantStepSize(0.01), mo["ErrVec"], mo["R_orig"])
    assert run == run_plain < 80 and diag.shape == (run, 3)
    np.testing.assert_array_equal(S, S_plain)
    assert np.all(diag[:, 0] < 1e-6) and np.all(diag[:, 1] < 1e-3)
