"""Oracle: quadratic cost constructors and transition log-probabilities (numpy restatement).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates
  QuadraticCosts.quadratic_cost_approximation_l2   donk/costs/quadratic_costs.py:121-139
  QuadraticCosts.quadratic_cost_approximation_l1   donk/costs/quadratic_costs.py:141-166
  LinearDynamics.log_prob                          donk/dynamics/linear_dynamics.py:54-64
      (scipy.stats.multivariate_normal.logpdf, allow_singular=False: pseudo-determinant and Mahalanobis distance from
       the eigendecomposition of the covariance; scipy 1.18.1 ``_PSD``)

Pinned by the reference's own known answers (tests/costs_test.py:196-257 literals, tests/dynamics_test.py:180-204) in
``tests/test_costs_logprob.py`` and, for log_prob, against live scipy in the same file.
"""
from __future__ import annotations

import numpy as np


def cost_l2(t, w):
    """quadratic_costs.py:133-139: C = diag(w), c = -w t, cc = t w t / 2."""
    t, w = np.asarray(t, dtype=np.float64), np.asarray(w, dtype=np.float64)
    p = t.shape[-1]
    C = np.zeros(t.shape + (p,))
    idx = np.arange(p)
    C[..., idx, idx] = w
    return C, -(w * t), np.sum(t * w * t, axis=-1) / 2


def cost_l1(xu, t, w, alpha):
    """quadratic_costs.py:155-166."""
    xu, t, w = (np.asarray(a, dtype=np.float64) for a in (xu, t, w))
    p = xu.shape[-1]
    sq = (t - xu) ** 2 + alpha
    ad = np.sqrt(sq)
    cu = sq * ad
    C = np.zeros(xu.shape + (p,))
    idx = np.arange(p)
    C[..., idx, idx] = alpha * w / cu
    c = w * ((xu - t) / ad - alpha * xu / cu)
    cc = np.sum(w * (ad + xu * (t - xu) / ad + alpha * xu ** 2 / cu / 2), axis=-1)
    return C, c, cc


def dynamics_log_prob(X, U, F, f, covar):
    """linear_dynamics.py:54-64 for ``X (..., T+1, dX), U (..., T, dU)`` -> ``(..., T)``."""
    X, U = np.asarray(X, dtype=np.float64), np.asarray(U, dtype=np.float64)
    T, dX = f.shape
    xu = np.concatenate([X[..., :-1, :], U], axis=-1)
    means = np.einsum("tij,...tj->...ti", F, xu) + f
    diff = X[..., 1:, :] - means
    out = np.empty(diff.shape[:-1])
    for t in range(T):
        s, u = np.linalg.eigh(covar[t])
        maha = np.sum(np.square((diff[..., t, :] @ u) / np.sqrt(s)), axis=-1)
        out[..., t] = -0.5 * (dX * np.log(2 * np.pi) + np.sum(np.log(s)) + maha)
    return out
