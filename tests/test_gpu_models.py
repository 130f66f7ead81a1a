"""GPU parity tests of the device data generators (csrc/gen.cu) against the counter-based CPU
restatement of Models/Uniform_Topology.m / Nonuniform_Topology.m (oracle/desc_models_ctr.py).

Bar: integer / index / mask outputs bit-exact; rotations to 1e-9 absolute (entries are O(1); the SVD projection of a
Gaussian 3x3 amplifies the last-ulp differences of log/cos/sin between the device and numpy by 1/(singular-value gap));
ErrVec to 1e-8 relative above the sqrt(eps) floor of acos at 1."""
import numpy as np
import pytest

import desc_b200
from oracle import desc_models_ctr as M
from oracle import desc_oracle as O

pytestmark = pytest.mark.gpu

ROT_ATOL = 1e-9


def _compare(got, ref):
    np.testing.assert_array_equal(got["Ind"], ref["Ind"])
    assert got["Ind"].shape[1] == 2 and got["RijMat"].shape == ref["RijMat"].shape
    np.testing.assert_array_equal(got["corrupted"], ref["corrupted"])
    np.testing.assert_allclose(got["R_orig"], ref["R_orig"], atol=ROT_ATOL, rtol=0)
    np.testing.assert_allclose(got["Rij_orig"], ref["Rij_orig"], atol=ROT_ATOL, rtol=0)
    np.testing.assert_allclose(got["RijMat"], ref["RijMat"], atol=ROT_ATOL, rtol=0)
    e, r = got["ErrVec"].ravel(), ref["ErrVec"].ravel()
    assert np.max(np.abs(e - r) - 1e-8 * np.abs(r)) <= 1e-7
    big = r > 1e-3
    assert np.max(np.abs(e[big] - r[big]) / r[big], initial=0.0) <= 1e-8
    # every output matrix is a rotation
    R = O.to_internal(got["RijMat"])
    np.testing.assert_allclose(R @ R.transpose(0, 2, 1), np.broadcast_to(np.eye(3), R.shape), atol=1e-12)
    np.testing.assert_allclose(np.linalg.det(R), 1.0, atol=1e-12)


@pytest.mark.parametrize("case", [(200, 0.5, 0.2, 0.0, "uniform", 0), (150, 0.3, 0.3, 0.1, "uniform", 1),
                                  (120, 0.6, 0.25, 0.05, "self-consistent", 2), (33, 1.0, 0.0, 0.2, "uniform", 3),
                                  (64, 0.05, 1.0, 0.0, "self-consistent", 4)])
def test_uniform_topology_matches_counter_oracle(case):
    n, p, q, sigma, model, seed = case
    got = desc_b200.Uniform_Topology(n, p, q, sigma, model, seed=seed, want_adj=True)
    ref = M.uniform_topology(n, p, q, sigma, model, seed=seed)
    _compare(got, ref)
    A = got["AdjMat"]
    assert A.shape == (n, n) and (A == A.T).all() and A.sum() == 2 * got["Ind"].shape[0]


@pytest.mark.parametrize("crpt_type", ["uniform", "self-consistent", "adv"])
@pytest.mark.parametrize("case", [(100, 0.5, 0.3, 0.5, 0.05, 0.1, 7), (257, 0.2, 0.5, 0.9, 0.0, 0.0, 8),
                                  (60, 0.7, 1.0, 1.0, 0.1, 0.0, 9), (50, 0.5, 0.0, 0.5, 0.1, 0.1, 10)])
def test_nonuniform_topology_matches_counter_oracle(case, crpt_type):
    n, p, pn, pe, s_in, s_out, seed = case
    got = desc_b200.Nonuniform_Topology(n, p, pn, pe, s_in, s_out, crpt_type, seed=seed)
    ref = M.nonuniform_topology(n, p, pn, pe, s_in, s_out, crpt_type, seed=seed)
    _compare(got, ref)


@pytest.mark.parametrize("case", [(400, 20, 40, 0.2, 0.05, 11), (101, 30, 50, 0.1, 0.0, 12), (90, 10, 60, 0.3, 0.1, 13)])
def test_ring_topology_matches_counter_oracle(case):
    n, deg, window, q, sigma, seed = case
    got = desc_b200.Ring_Topology(n, deg, window, q, sigma, seed=seed)
    ref = M.uniform_topology(n, deg / (2.0 * window), q, sigma, "uniform", seed=seed, ring=window)
    _compare(got, ref)


def test_generated_model_feeds_the_solver_without_leaving_the_device():
    """Model buffers (HBM, MATLAB layout) go straight into the solver; same result as through host arrays"""
    with desc_b200.Uniform_Topology(300, 0.4, 0.2, 0.1, seed=21, on_device=True) as mo:
        assert mo.n == 300 and mo.m > 0 and mo.launches >= 6 and mo.gen_ms > 0.0
        host = mo.to_host()
        params = dict(iters=20, Gradient=desc_b200.ConstantStepSize(0.01), seed=2)
        R_dev, S_dev = desc_b200.DESC_init(mo.Ind, mo.RijMat, params)
    R_host, S_host = desc_b200.DESC_init(host["Ind"], host["RijMat"], dict(params, Gradient=desc_b200.ConstantStepSize(0.01)))
    np.testing.assert_array_equal(S_dev, S_host)
    np.testing.assert_array_equal(R_dev, R_host)
    # and the solve recovers the planted corruption / rotations
    assert np.mean(np.abs(S_dev.ravel() - host["ErrVec"].ravel())) < 0.05
    _, _, mean_err, _ = desc_b200.Rotation_Alignment(R_dev, host["R_orig"])
    assert mean_err < 3.0


def test_generator_is_a_pure_function_of_its_seed():
    a = desc_b200.Uniform_Topology(120, 0.5, 0.2, 0.1, seed=5)
    b = desc_b200.Uniform_Topology(120, 0.5, 0.2, 0.1, seed=5)
    c = desc_b200.Uniform_Topology(120, 0.5, 0.2, 0.1, seed=6)
    np.testing.assert_array_equal(a["RijMat"], b["RijMat"])
    np.testing.assert_array_equal(a["Ind"], b["Ind"])
    assert a["Ind"].shape != c["Ind"].shape or not np.array_equal(a["Ind"], c["Ind"])


def test_headline_size_generates_in_milliseconds():
    """cfg 4 shape (n = 10^4, p = 0.1): ~5e6 edges; distribution checks instead of the O(n^2) numpy oracle"""
    with desc_b200.Uniform_Topology(10000, 0.1, 0.2, 0.1, seed=1, on_device=True) as mo:
        assert abs(mo.m - 0.1 * 10000 * 9999 / 2) < 5 * np.sqrt(0.1 * 0.9 * 10000 * 9999 / 2)
        h = mo.to_host()
        assert mo.gen_ms < 2000.0
    assert abs(h["corrupted"].mean() - 0.2) < 2e-3
    i, j = h["Ind"][:, 0], h["Ind"][:, 1]
    assert (i < j).all() and (np.diff(i * 10001 + j) > 0).all()
    assert 0.2 < np.median(h["ErrVec"].ravel()[h["corrupted"]]) and np.median(h["ErrVec"].ravel()[~h["corrupted"]]) < 0.1
