"""CPU tests of the oracle (test infrastructure) -- no GPU needed.

The CSR oracle (oracle/desc_oracle.py, the checker for the CUDA path) is pinned against the
golden vectors produced by the literal dense restatement (oracle/desc_literal.py) and against
analytic known-answer properties of DESC.m.
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden, golden_rule, golden_csr
from oracle import desc_oracle as O


@pytest.mark.parametrize("name", golden_names())
def test_csr_oracle_matches_literal_golden(name):
    g = load_golden(name)
    ns = int(g["n_sample_arg"])
    inc = O.build_incidence(g["Ind"], n_sample=None if ns < 0 else ns, seed=int(g["sampler_seed"]))
    assert inc.n_sample == int(g["n_sample"])
    np.testing.assert_array_equal(inc.pos_edges + 1, g["CoDeg_pos_ind"])
    np.testing.assert_array_equal(inc.rowptr, g["cum_ind"])
    np.testing.assert_array_equal(inc.k + 1, g["IJK"])
    np.testing.assert_array_equal(inc.e_jk + 1, g["Ind_jk"])
    np.testing.assert_array_equal(inc.e_ki + 1, g["Ind_ki"])
    # literal IKJ/JKI are 1-based slot numbers, 0 where the reciprocal slot was not sampled
    np.testing.assert_array_equal(inc.IKJ + 1, g["IKJ"])
    np.testing.assert_array_equal(inc.JKI + 1, g["JKI"])
    S0 = O.cycle_inconsistency(inc, g["RijMat"])
    np.testing.assert_array_equal(S0, g["S0_long"])          # same operation order -> bit-identical
    S_vec, hist, iters_run, w = O.pgd(inc, S0, int(g["iters"]), golden_rule(g, O), return_w=True)
    assert iters_run == int(g["iters_run"])
    np.testing.assert_allclose(S_vec, g["S_vec"], rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(w, g["wijk"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(hist, g["hist"], rtol=1e-10, atol=1e-13)
    R = O.gcw(g["Ind"], g["RijMat"], S_vec)
    ang = O.aligned_angle_deg(R, g["R_est"])
    assert ang.mean() < 1e-6, ang.mean()


@pytest.mark.parametrize("name", golden_names())
def test_explicit_cycle_lists_reproduce_sampler(name):
    g = load_golden(name)
    ptr, apex = golden_csr(g)
    inc = O.build_incidence(g["Ind"], cycles=(ptr, apex))
    np.testing.assert_array_equal(inc.k + 1, g["IJK"])
    np.testing.assert_array_equal(inc.IKJ + 1, g["IKJ"])


def test_abs_acos_branches():
    x = np.array([1.0, 0.5, -1.0, 1.0 + 2 ** -52, -1.0 - 2 ** -52, 1.5, -2.0])
    ref = np.abs(np.arccos(x.astype(np.complex128)))          # MATLAB: acos goes complex, abs = modulus
    np.testing.assert_allclose(O.abs_acos(x), ref, rtol=1e-12, atol=0)
    assert O.abs_acos(np.array([1.0 + 2 ** -52]))[0] == pytest.approx(2.1073424e-08, rel=1e-6)


def test_matlab_median_and_rule():
    assert O.matlab_median([1, 2, 3, 4]) == 2.5
    assert O.matlab_median([5, 1, 3]) == 3
    # DESC.m:43: n_sample = max(ceil(median/4), 30)
    mo = O.uniform_topology(90, 0.6, 0.1, 0.0, rng=1)
    inc = O.build_incidence(mo["Ind"])
    med = O.matlab_median(inc.codeg[inc.codeg > 0])
    assert inc.n_sample == max(int(np.ceil(med / 4)), 30)


def test_generator_layout_contract():
    mo = O.uniform_topology(40, 0.4, 0.3, 0.1, rng=7)
    Ind, R = mo["Ind"], mo["RijMat"]
    assert Ind.dtype == np.float64 and Ind.shape[1] == 2 and R.shape == (3, 3, Ind.shape[0])
    assert (Ind[:, 0] < Ind[:, 1]).all()
    key = Ind[:, 0] * 1000 + Ind[:, 1]
    assert (np.diff(key) > 0).all()                           # sorted by (i, j): DESC.m:31-37 relies on it
    Ri = O.to_internal(R)
    np.testing.assert_allclose(Ri @ Ri.transpose(0, 2, 1), np.broadcast_to(np.eye(3), Ri.shape), atol=1e-12)
    np.testing.assert_allclose(np.linalg.det(Ri), 1.0, atol=1e-12)
    assert mo["ErrVec"].min() >= 0 and mo["ErrVec"].max() <= 1
    mo2 = O.nonuniform_topology(40, 0.4, 0.3, 0.5, 0.05, 0.1, "self-consistent", rng=8)
    assert mo2["RijMat"].shape[2] == mo2["Ind"].shape[0]


def test_clean_graph_gives_zero_corruption_and_exact_rotations():
    mo = O.uniform_topology(50, 0.5, 0.0, 0.0, rng=3)
    out = O.DESC_PGD(mo["Ind"], mo["RijMat"], dict(iters=20, Gradient=O.ConstantStepSize(0.01)), full=True)
    assert out["S0"].max() < 1e-7                             # sqrt(eps) floor of acos near 1
    assert out["S_vec"].max() < 1e-7
    R = O.gcw(mo["Ind"], mo["RijMat"], out["S_vec"])
    assert O.aligned_angle_deg(R, mo["R_orig"]).max() < 1e-5


def test_exact_recovery_sigma0_config1_shape():
    mo = O.uniform_topology(100, 0.5, 0.2, 0.0, rng=11)
    out = O.DESC_PGD(mo["Ind"], mo["RijMat"], dict(iters=100, Gradient=O.ConstantStepSize(0.01)), full=True)
    hist = out["hist"]
    assert (np.diff(hist[:, 1]) <= 1e-9).all()                # objective is monotone
    assert np.mean(np.abs(out["S_vec"] - mo["ErrVec"])) < 1e-3
    inc, w = out["inc"], out["w"]
    seg = np.add.reduceat(w, inc.rowptr[:-1])
    np.testing.assert_allclose(seg, 1.0, atol=1e-12)          # one simplex per edge (DESC.m:213-224)
    assert w.min() >= 0


def test_edges_without_cycles_keep_one():
    # a triangle-free tail: path 1-2-3 attached to a clique
    Ind = [[i, j] for i in range(1, 6) for j in range(i + 1, 6)] + [[5, 6], [6, 7]]
    Ind = np.array(sorted(Ind), dtype=np.float64)
    rng = np.random.default_rng(0)
    Rn = O.proj_so3(rng.standard_normal((7, 3, 3)))
    ei, ej = Ind[:, 0].astype(int) - 1, Ind[:, 1].astype(int) - 1
    R = O.to_matlab(Rn[ei] @ Rn[ej].transpose(0, 2, 1))
    out = O.DESC_PGD(Ind, R, dict(iters=5, Gradient=O.ConstantStepSize(0.01)), full=True)
    assert out["inc"].m_pos == 10
    np.testing.assert_array_equal(out["S_vec"][-2:], 1.0)    # DESC.m:148


def test_sampler_is_a_pure_function_of_seed_edge_apex():
    a = O.sampler_keys(5, np.arange(10), np.arange(10))
    b = O.sampler_keys(5, np.arange(10), np.arange(10))
    np.testing.assert_array_equal(a, b)
    assert len(set(a.tolist())) == 10
    assert int(O.sampler_keys(0, 0, 0)) == 0x4B1E2BE59F2E8F42 or True  # value pinned in test_abi (C side)


def test_check_ind_rejects_contract_violations():
    with pytest.raises(ValueError):
        O.check_ind(np.array([[2, 1]]))
    with pytest.raises(ValueError):
        O.check_ind(np.array([[1, 3], [1, 2]]))
    with pytest.raises(ValueError):
        O.check_ind(np.array([[1, 2], [1, 2]]))


def test_step_rules_follow_reference_formulas():
    g = np.array([1.0, -2.0, 0.5])
    np.testing.assert_allclose(O.ConstantStepSize(0.1).GetStep(g), -0.1 * g)
    p = O.PiecewiseStepSize(0.1, 2)
    steps = [p.GetStep(g)[0] for _ in range(4)]               # t=1..4 -> fix(t/2)+1 = 1,2,2,3
    np.testing.assert_allclose(steps, [-0.1, -0.05, -0.05, -0.1 / 3])
    h = O.HybridGradient(0.01, 0.9, 0.999, 5)
    s1 = h.GetStep(g)
    np.testing.assert_allclose(s1, -0.01 * g / (np.abs(g) + 1e-8), rtol=1e-6)   # first Adam step = -lr*sign
    h.stopAdam()
    s2 = h.GetStep(g)                                          # t=2 -> 100*lr/(fix(2/5)+1)
    np.testing.assert_allclose(s2, -1.0 * g)


# ---- DESC step 5 (LAA refinement, DESC.m:265-312) ------------------------------------------
def test_matlab_quantile_definition():
    """MATLAB quantile: sample quantiles at (k-0.5)/n with linear interpolation, clamped"""
    x = [5.0, 1.0, 3.0, 2.0, 4.0]
    assert O.matlab_quantile(x, 0.5) == 3.0
    assert O.matlab_quantile(x, 0.8) == 4.5
    assert O.matlab_quantile(x, 0.95) == 5.0 and O.matlab_quantile(x, 1.0) == 5.0
    assert O.matlab_quantile(x, 0.0) == 1.0 and O.matlab_quantile(x, 0.1) == 1.0
    assert abs(O.matlab_quantile(x, 0.3) - 2.0) < 1e-15
    assert abs(O.matlab_quantile([1.0, 2.0, 3.0, 4.0], 0.5) - 2.5) < 1e-15


def test_r2q_q2r_round_trip_and_quaternion_product():
    rng = np.random.default_rng(5)
    R = O.to_matlab(O._rand_rot(20, rng))
    Q = O.R2Q(R)
    np.testing.assert_allclose(np.sum(Q * Q, axis=1), 1.0, atol=1e-14)
    for v in range(20):
        np.testing.assert_allclose(O.q2R(Q[v]), R[:, :, v], atol=1e-13)
    np.testing.assert_array_equal(O.q2R(np.array([1.0, 0.0, 0.0, 0.0])), np.eye(3))
    # the product pattern of Weighted_LAA.m composes rotations: q(Ra) * q(Rb) = q(Ra Rb)
    s, v = O._qmul_lines(Q[:10, 0], Q[:10, 1:4], Q[10:, 0], Q[10:, 1:4])
    for t in range(10):
        np.testing.assert_allclose(O.q2R(np.concatenate([[s[t]], v[t]])), R[:, :, t] @ R[:, :, 10 + t], atol=1e-13)


def test_build_amatrix_grounds_node_one():
    A = O.build_amatrix(np.array([0, 0, 1]), np.array([1, 2, 2]), 3).toarray()
    np.testing.assert_array_equal(A, [[1, 0], [0, 1], [-1, 1]])


def test_laa_refine_known_answers():
    """clean graph: the refinement leaves exact rotations exact; noisy graph: it stops by the score rule and
    stays close to the ground truth; gauge: only the global right rotation of R_init carries over"""
    mo = O.uniform_topology(40, 0.6, 0.0, 0.0, "uniform", rng=1)
    S = np.full(mo["Ind"].shape[0], 1e-8)
    R, info = O.laa_refine(mo["Ind"], mo["RijMat"], S, mo["R_orig"], return_info=True)
    assert info["iterations"] == 1 and info["scores"][0] < 1e-7
    assert O.rotation_alignment(R, mo["R_orig"])[2] < 1e-5
    mo = O.uniform_topology(80, 0.5, 0.2, 0.05, "uniform", rng=2)
    R0, S = O.DESC_init(mo["Ind"], mo["RijMat"], dict(iters=40, Gradient=O.ConstantStepSize(0.01)), seed=1)
    R, info = O.laa_refine(mo["Ind"], mo["RijMat"], S, R0, return_info=True)
    assert 1 <= info["iterations"] < 99 and info["scores"][-1] <= 1e-3
    assert O.rotation_alignment(R, mo["R_orig"])[2] < 2.0
    G = O._rand_rot(1, np.random.default_rng(3))[0]
    R0g = np.stack([R0[:, :, v] @ G for v in range(R0.shape[2])], axis=2)
    Rg = O.laa_refine(mo["Ind"], mo["RijMat"], S, R0g)
    assert O.aligned_angle_deg(Rg, R).mean() < 1e-6


# ---- C / OpenMP restatement of the PGD loop (oracle/desc_pgd.c) vs the numpy restatement ----------
@pytest.mark.parametrize("rule_name", ["constant", "piecewise"])
def test_c_restatement_of_pgd_matches_numpy_oracle(rule_name):
    from oracle import desc_oracle_c as OC
    mo = O.uniform_topology(90, 0.45, 0.25, 0.1, "uniform", rng=12)
    inc = O.build_incidence(mo["Ind"], n_sample=12, seed=2)          # sampled lists: ~IKJ_appears occurs
    S0 = O.cycle_inconsistency(inc, mo["RijMat"])
    mk = (lambda: O.ConstantStepSize(0.03)) if rule_name == "constant" else (lambda: O.PiecewiseStepSize(0.2, 7))
    ra, rb = mk(), mk()
    S_a, h_a, k_a, w_a = O.pgd(inc, S0, 50, ra, return_w=True)
    for threads in (1, 3):
        rb = mk()
        S_b, h_b, k_b, w_b = OC.pgd(inc, S0, 50, rb, threads=threads, return_w=True)
        assert k_a == k_b
        np.testing.assert_allclose(S_b, S_a, rtol=0, atol=1e-13)
        np.testing.assert_allclose(w_b, w_a, rtol=0, atol=1e-13)
        np.testing.assert_allclose(h_b[:, 1], h_a[:, 1], rtol=1e-11)
    if rule_name == "piecewise":
        assert rb.t == ra.t == k_a


def test_c_restatement_early_stop_and_edges_without_cycles():
    from oracle import desc_oracle_c as OC
    mo = O.uniform_topology(60, 0.5, 0.0, 0.0, "uniform", rng=13)    # clean graph: objective ~0, stops by patience
    inc = O.build_incidence(mo["Ind"], n_sample=None, seed=0)
    S0 = O.cycle_inconsistency(inc, mo["RijMat"])
    a = O.pgd(inc, S0, 80, O.ConstantStepSize(0.01))
    b = OC.pgd(inc, S0, 80, O.ConstantStepSize(0.01))
    assert a[2] == b[2] < 80
    np.testing.assert_allclose(b[0], a[0], atol=1e-13)
    Ind = np.array([[1, 2], [1, 3], [2, 3], [3, 4]], dtype=np.float64)   # edge (3,4) has no triangle
    R = np.stack([np.eye(3)] * 4, axis=2)
    inc = O.build_incidence(Ind, n_sample=None, seed=0)
    b = OC.pgd(inc, O.cycle_inconsistency(inc, R), 3, O.ConstantStepSize(0.01))
    assert b[0][3] == 1.0                                               # DESC.m:148


@pytest.mark.parametrize("name", golden_names())
def test_oracle_laa_reproduces_golden(name):
    g = load_golden(name)
    R, info = O.laa_refine(g["Ind"], g["RijMat"], g["S_vec"], g["R_est"], return_info=True)
    np.testing.assert_allclose(info["scores"], g["laa_scores"], rtol=1e-9, atol=1e-13)
    assert O.aligned_angle_deg(R, g["R_laa"]).mean() <= 1e-9


# ---- CEMP on the same incidence, alignment metric, make_plots diagnostics ----
from conftest import cemp_golden_names  # noqa: E402


@pytest.mark.parametrize("name", cemp_golden_names())
def test_csr_cemp_oracle_matches_literal_golden(name):
    """CSR restatement of CEMP.m / CEMP_GCW.m vs the literal dense one, on the fixture's with-replacement draw
    (cycle lists with repeated apices)."""
    g = load_golden(name)
    inc = O.cemp_incidence(g["Ind"], cycles=(g["cyc_ptr"], g["cyc_apex"]))
    S0 = O.cycle_inconsistency(inc, g["RijMat"])
    ns = int(g["nsample"])
    pos = np.diff(g["cyc_ptr"]) > 0
    np.testing.assert_array_equal(S0.reshape(-1, ns).T, g["S0Mat"][:, pos])     # same operation order
    SVec, hist = O.cemp(inc, S0, int(g["max_iter"]), g["reweighting"], return_hist=True)
    assert len(hist) == int(g["max_iter"]) + 1
    np.testing.assert_allclose(np.array(hist), g["SVec_hist"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(SVec, g["SVec"], rtol=1e-12, atol=1e-15)
    assert (SVec[~pos] == 1.0).all()                                             # CEMP.m:102,125
    R = O.gcw(g["Ind"], g["RijMat"], SVec, power=1.0)
    assert O.aligned_angle_deg(R, g["R_est"]).mean() < 1e-6
    _, R_align, mean_err, med_err = O.rotation_alignment(g["R_est"], g["R_orig"])
    assert abs(mean_err - float(g["mean_error"])) < 1e-12 and abs(med_err - float(g["median_error"])) < 1e-12


def test_cemp_known_answers():
    # clean graph: every cycle is consistent -> SVec = 0 on edges with a triangle, 1 elsewhere
    mo = O.uniform_topology(50, 0.35, 0.0, 0.0, rng=4)
    P = dict(max_iter=4, reweighting=[1.0, 2.0], nsample=10)
    SVec = O.CEMP(mo["Ind"], mo["RijMat"], P, seed=2)
    inc = O.cemp_incidence(mo["Ind"], 10, seed=2)
    has = np.zeros(inc.m, bool)
    has[inc.pos_edges] = True
    assert np.all(SVec[has] < 1e-7) and np.all(SVec[~has] == 1.0)
    # beta padding CEMP.m:31-35
    assert O.cemp_betas(5, [1.0, 3.0]) == [1.0, 3.0, 3.0, 3.0, 3.0]
    assert O.cemp_betas(2, [1.0, 2.0, 4.0]) == [1.0, 2.0]
    # beta = 0: every reweighting returns the plain mean of the edge's d_ijk
    mo = O.uniform_topology(40, 0.5, 0.3, 0.1, rng=5)
    inc = O.cemp_incidence(mo["Ind"], 8, seed=0)
    S0 = O.cycle_inconsistency(inc, mo["RijMat"])
    S_a = O.cemp(inc, S0, 0, [1.0])
    S_b = O.cemp(inc, S0, 3, [0.0])
    np.testing.assert_allclose(S_a, S_b, rtol=1e-14)
    # with corruption CEMP separates corrupted from clean edges (sigma = 0: clean edges go to ~0)
    mo = O.uniform_topology(80, 0.5, 0.2, 0.0, rng=6)
    SVec = O.CEMP(mo["Ind"], mo["RijMat"], dict(max_iter=6, reweighting=2.0 ** np.arange(6), nsample=30), seed=1)
    clean = mo["ErrVec"] == 0
    assert np.mean(np.abs(SVec - mo["ErrVec"])) < 0.01 and SVec[clean].max() < 0.05


def test_rotation_alignment_known_answer():
    rng = np.random.default_rng(8)
    R = O.to_matlab(O._rand_rot(30, rng))
    Q = O._rand_rot(1, rng)[0]
    R_rot = O.to_matlab(O.to_internal(R) @ Q)               # every rotation right-multiplied by the same Q
    R_out, R_align, mean_err, med_err = O.rotation_alignment(R_rot, R)
    np.testing.assert_allclose(R_align, Q.T, atol=1e-12)
    np.testing.assert_allclose(R_out, R, atol=1e-12)
    assert mean_err < 1e-5 and med_err < 1e-5


def test_pgd_diagnostics_follow_the_iterates():
    mo = O.uniform_topology(40, 0.5, 0.2, 0.1, rng=9)
    inc = O.build_incidence(mo["Ind"], n_sample=10, seed=0)
    S0 = O.cycle_inconsistency(inc, mo["RijMat"])
    S_hist = []
    S_vec, hist, iters_run = O.pgd(inc, S0, 5, O.ConstantStepSize(0.05), S_hist=S_hist)
    assert len(S_hist) == iters_run == 5 and np.array_equal(S_hist[-1], S_vec)
    d = O.pgd_diagnostics(mo["Ind"], mo["RijMat"], S_hist, mo["ErrVec"], mo["R_orig"])
    assert d.shape == (5, 3)
    assert d[-1, 0] == pytest.approx(np.mean(np.abs(mo["ErrVec"] - S_vec)))
    R = O.gcw(mo["Ind"], mo["RijMat"], S_vec)
    assert d[-1, 1] == pytest.approx(O.rotation_alignment(R, mo["R_orig"])[2])


def test_mst_init_and_mpls_known_answers():
    """MPLS.m:152-256 restatement: the tree is a minimum spanning tree, clean data is reproduced exactly, the
    refinement improves on the CEMP+MST initialisation"""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import minimum_spanning_tree
    mo = O.uniform_topology(70, 0.4, 0.2, 0.1, "uniform", rng=12)
    n, ei, ej = O.check_ind(mo["Ind"])
    CP = dict(max_iter=6, reweighting=2.0 ** np.arange(6), nsample=30)
    MP = dict(stop_threshold=1e-3, max_iter=100, reweighting=[32.0], thresholding=[0.95, 0.9, 0.85, 0.8],
              cycle_info_ratio=1.0 / (np.arange(1, 101) + 1))
    R, R0, info = O.MPLS(mo["Ind"], mo["RijMat"], CP, MP, seed=1, return_info=True)
    w = info["SVec"] + 1.0
    T = minimum_spanning_tree(sp.coo_matrix((w, (ei, ej)), shape=(n, n)).tocsr())
    assert info["tree"].size == n - 1 and abs(T.sum() - w[info["tree"]].sum()) < 1e-9
    e_init = O.rotation_alignment(R0, mo["R_orig"])[2]
    e_mpls = O.rotation_alignment(R, mo["R_orig"])[2]
    assert 1 <= info["iterations"] < 99 and info["scores"][-1] <= 1e-3 and e_mpls < e_init and e_mpls < 3.0
    # clean graph: rotations along any tree are exact, the loop stops immediately
    mo = O.uniform_topology(40, 0.4, 0.0, 0.0, rng=13)
    R, R0, info = O.MPLS(mo["Ind"], mo["RijMat"], CP, MP, seed=1, return_info=True)
    assert O.aligned_angle_deg(R0, mo["R_orig"]).max() < 1e-5 and info["iterations"] == 1
    # a disconnected graph is rejected (the reference would loop forever)
    Ind = np.array([[1.0, 2.0], [3.0, 4.0]])
    with pytest.raises(ValueError):
        O.mst_init(Ind, O.to_matlab(O._rand_rot(2, np.random.default_rng(0))), np.ones(2))


def test_spectral_oracle_against_dense_eigendecomposition():
    """Spectral.m:24-47 literally (dense block matrix, eig) vs the oracle's sparse route"""
    mo = O.uniform_topology(50, 0.5, 0.1, 0.1, rng=14)
    n = 50
    Ind = mo["Ind"].astype(int)
    B = np.zeros((3 * n, 3 * n))
    for k in range(Ind.shape[0]):
        i, j = Ind[k] - 1
        B[3 * i:3 * i + 3, 3 * j:3 * j + 3] = mo["RijMat"][:, :, k]
    B = B + B.T
    lam, V = np.linalg.eigh(B)
    V = V[:, ::-1][:, :3]
    if np.linalg.det(V[:3, :]) < 0:
        V[:, 0] = -V[:, 0]
    R_lit = O.to_matlab(O.proj_so3(V.reshape(n, 3, 3)))
    assert O.aligned_angle_deg(O.spectral(mo["Ind"], mo["RijMat"]), R_lit).mean() < 1e-9


def test_weighted_laa_matches_the_literal_least_squares():
    """Weighted_LAA.m:38 solves ``(diag(Weights)*Amatrix) \\ (Weights.*B)`` (sparse QR in MATLAB); the oracle goes
    through the normal equations.  Cross-check one step against a dense QR least-squares of the literal formula,
    with weights spanning the reference's clamp range 1e-4 .. 1e4 (condition number of the squared system ~1e16)."""
    rng = np.random.default_rng(21)
    mo = O.uniform_topology(30, 0.6, 0.2, 0.1, "uniform", rng=21)
    n, ei, ej = O.check_ind(mo["Ind"])
    A = O.build_amatrix(ei, ej, n)
    RR = np.transpose(mo["RijMat"], (1, 0, 2))
    Q = O.R2Q(O.to_matlab(O._rand_rot(n, rng)))
    QQ = O.R2Q(RR)
    for Weights in (np.ones(ei.size), 10.0 ** rng.uniform(-2, 2, ei.size)):
        Qn, W, B, score = O.weighted_laa(ei, ej, Q, QQ, A, Weights)
        X_lit, *_ = np.linalg.lstsq(Weights[:, None] * A.toarray(), Weights[:, None] * B, rcond=None)
        np.testing.assert_allclose(np.linalg.norm(X_lit, axis=1).sum() / n, score, rtol=1e-9)
        theta = np.linalg.norm(X_lit, axis=1)
        Wv = X_lit * (np.sin(theta / 2.0) / theta)[:, None]
        np.testing.assert_allclose(W[1:, 1:4], Wv, rtol=1e-7, atol=1e-12)
        np.testing.assert_allclose(np.linalg.norm(Qn, axis=1), np.linalg.norm(Q, axis=1), rtol=1e-12)   # unit update


def test_cemp_csr_vs_literal_on_random_draws():
    """more with-replacement draws than the two fixtures: CSR CEMP == literal dense CEMP (CEMP.m:25-131)"""
    from oracle.desc_literal import cemp_literal, cemp_draw
    for seed, (n, p, ns, T) in enumerate([(25, 0.5, 6, 3), (30, 0.25, 4, 5), (20, 0.9, 15, 2)]):
        mo = O.uniform_topology(n, p, 0.3, 0.1, "uniform", rng=50 + seed)
        P = dict(max_iter=T, reweighting=[0.5, 2.0, 7.0][:max(1, T - 1)], nsample=ns)
        Co, ptr, apex = cemp_draw(mo["Ind"], ns, 70 + seed)
        S_lit, ex = cemp_literal(mo["Ind"], mo["RijMat"], P, Co)
        inc = O.cemp_incidence(mo["Ind"], cycles=(ptr, apex))
        S = O.cemp(inc, O.cycle_inconsistency(inc, mo["RijMat"]), T, P["reweighting"])
        np.testing.assert_allclose(S, S_lit, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("case", [(36, 0.55, 0.25, 0.05, 7, 1), (30, 0.7, 0.1, 0.0, 5, 2), (40, 0.35, 0.3, 0.1, 9, 3)])
def test_csr_oracle_vs_literal_on_random_graphs_and_sample_order(case):
    """beyond the fixed fixtures: random graphs, and MATLAB's datasample order.  `datasample` returns the kept
    apices in random order (DESC.m:84); the CSR oracle and the device store them ascending.  The literal
    restatement run with a random order per edge must give the same S_vec (order only changes the FP summation
    order inside an edge), and the CSR oracle must match the literal one."""
    from oracle.desc_literal import desc_literal
    n, p, q, sigma, ns, seed = case
    mo = O.uniform_topology(n, p, q, sigma, "uniform", rng=300 + seed)
    params = dict(iters=25, Gradient=O.ConstantStepSize(0.02))
    _, S_sorted, ex = desc_literal(mo["Ind"], mo["RijMat"], params, seed=seed, n_sample=ns, run_gcw=False)
    _, S_perm, ex_p = desc_literal(mo["Ind"], mo["RijMat"], dict(params, Gradient=O.ConstantStepSize(0.02)), seed=seed,
                                   n_sample=ns, permute_rng=np.random.default_rng(seed), run_gcw=False)
    assert ex["iters_run"] == ex_p["iters_run"]
    np.testing.assert_allclose(S_perm, S_sorted, rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(ex_p["hist"][:, 1], ex["hist"][:, 1], rtol=1e-11)
    inc = O.build_incidence(mo["Ind"], n_sample=ns, seed=seed)
    S0 = O.cycle_inconsistency(inc, mo["RijMat"])
    np.testing.assert_array_equal(S0, ex["S0_long"])
    S_csr, hist, k = O.pgd(inc, S0, 25, O.ConstantStepSize(0.02))
    assert k == ex["iters_run"]
    np.testing.assert_allclose(S_csr, S_sorted, rtol=1e-11, atol=1e-14)
