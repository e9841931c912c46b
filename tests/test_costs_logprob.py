"""The steps either side of the hot path: cost constructors and LinearDynamics.log_prob on the
device, against the reference's literal known answers (tests/costs_test.py:196-257), the oracle and live scipy."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from donk_testutil import relerr, to_np
from oracle import costs as ocosts


def test_oracle_costs_match_reference_literals():
    C, c, cc = ocosts.cost_l2(np.array([1, 2, 3]), np.array([1, 1, 2]))
    assert_allclose(C, np.diag([1, 1, 2])); assert_allclose(c, [-1, -2, -6]); assert_allclose(cc, 0.5 + 2 + 9)
    C, c, cc = ocosts.cost_l1(np.array([3, 2]), np.array([2, 2]), np.array([1, 2]), 1e-2)
    assert_allclose(C, np.diag([0.01 / 1.01**1.5, 20]))
    assert_allclose(c, [1 / 1.01**0.5 - 0.03 / 1.01**1.5, -40])
    assert_allclose(cc, 1.01**0.5 - 3 / 1.01**0.5 + 0.045 / 1.01**1.5 + 0.2 + 40)


def test_oracle_log_prob_matches_scipy():
    from scipy.stats import multivariate_normal

    rng = np.random.default_rng(0)
    T, dX, dU, N = 4, 5, 2, 3
    F, f = rng.standard_normal((T, dX, dX + dU)), rng.standard_normal((T, dX))
    A = rng.standard_normal((T, dX, dX))
    covar = A @ np.swapaxes(A, 1, 2) + 0.1 * np.eye(dX)
    X, U = rng.standard_normal((N, T + 1, dX)), rng.standard_normal((N, T, dU))
    got = ocosts.dynamics_log_prob(X, U, F, f, covar)
    for n in range(N):
        for t in range(T):
            mean = F[t] @ np.concatenate([X[n, t], U[n, t]]) + f[t]
            assert_allclose(got[n, t], multivariate_normal.logpdf(X[n, t + 1], mean, covar[t]), rtol=1e-12)


def test_cost_constructors_reference_known_answers(backend):
    """tests/costs_test.py:196-257, literally, through the reference-shaped class (device kernels underneath)."""
    from donk_b200.costs import QuadraticCosts

    cf = QuadraticCosts.quadratic_cost_approximation_l2(t=np.array([1, 2, 3]), w=np.array([1, 1, 2]))
    assert_allclose(cf.C, np.diag([1, 1, 2])); assert_allclose(cf.c, [-1, -2, -6]); assert_allclose(cf.cc, 0.5 + 2 + 9)
    cf = QuadraticCosts.quadratic_cost_approximation_l2(t=np.array([[-2, 2, 0], [-1, 1, 0], [0, 0, 0]]),
                                                        w=np.array([[1, 1, 2], [1, 1, 2], [10, 10, 0]]))
    assert_allclose(cf.C, [np.diag([1, 1, 2]), np.diag([1, 1, 2]), np.diag([10, 10, 0])])
    assert_allclose(cf.c, [[2, -2, 0], [1, -1, 0], [0, 0, 0]]); assert_allclose(cf.cc, [4, 1, 0])
    cf = QuadraticCosts.quadratic_cost_approximation_l1(xu=np.array([3, 2]), t=np.array([2, 2]), w=np.array([1, 2]), alpha=1e-2)
    assert_allclose(cf.C, np.diag([0.01 / 1.01**1.5, 20]))
    assert_allclose(cf.c, [1 / 1.01**0.5 - 0.03 / 1.01**1.5, -40])
    assert_allclose(cf.cc, 1.01**0.5 - 3 / 1.01**0.5 + 0.045 / 1.01**1.5 + 0.2 + 40)
    cf = QuadraticCosts.quadratic_cost_approximation_l1(xu=np.array([[-2, 0], [0, 0]]), t=np.array([[1, 2], [10, 0]]),
                                                        w=np.array([[1, 2], [10, 0]]), alpha=1e-6)
    assert_allclose(cf.C, np.zeros((2, 2, 2)), atol=1e-6); assert_allclose(cf.c, [[-1, -2], [-10, 0]], atol=1e-6)
    assert_allclose(cf.cc, [3 * 1 - 2 + 2 * 2, 10 * 10], atol=1e-6)


def test_batched_cost_constructors_bit_exact_vs_oracle(backend):
    from donk_b200 import batched

    rng = np.random.default_rng(4)
    B, T, p = 3, 7, 9
    t, w, xu = rng.standard_normal((B, T + 1, p)), rng.random((B, T + 1, p)), rng.standard_normal((B, T + 1, p))
    C, c, cc = batched.quadratic_cost_approximation_l2(t, w)
    Co, co, cco = ocosts.cost_l2(t, w)
    assert np.array_equal(to_np(C), Co) and np.array_equal(to_np(c), co) and relerr(cc, cco) < 1e-14
    C, c, cc = batched.quadratic_cost_approximation_l1(xu, t, w, 1e-3)
    Co, co, cco = ocosts.cost_l1(xu, t, w, 1e-3)
    assert relerr(C, Co) < 1e-14 and relerr(c, co) < 1e-14 and relerr(cc, cco) < 1e-14
    with pytest.raises(ValueError):
        batched.quadratic_cost_approximation_l2(t, w[:, :-1])


@pytest.mark.parametrize("dX,dU", [(5, 2), (20, 6)])
def test_batched_log_prob_vs_oracle(backend, dX, dU):
    from donk_b200 import batched, synthetic

    B, T, N = 2, 6, 11
    conds = [synthetic.condition(41, i, T, dX, dU, N + 2 * (2 * dX + dU)) for i in range(B)]
    X, U = np.stack([c.X for c in conds]), np.stack([c.U for c in conds])
    F, f, covar, info = batched.fit_lr(X, U)
    lp = to_np(batched.dynamics_log_prob(X[:, :N], U[:, :N], F, f, covar))
    assert lp.shape == (B, N, T)
    for b in range(B):
        ref = ocosts.dynamics_log_prob(X[b, :N], U[b, :N], to_np(F[b]), to_np(f[b]), to_np(covar[b]))
        assert relerr(lp[b], ref) < 1e-8
    bad = to_np(covar).copy()
    bad[1, 2] = -np.eye(dX)
    with pytest.raises(np.linalg.LinAlgError):
        batched.dynamics_log_prob(X[:, :N], U[:, :N], F, f, bad)
