"""Oracle: full-covariance Gaussian-mixture EM and the GMM dynamics prior.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference delegates this arithmetic to scikit-learn
(``GMMPrior.update`` -> ``sklearn.mixture.GaussianMixture.fit``,
donk/dynamics/prior/prior_gmm.py:14-24; ``GMMPrior.eval`` -> ``predict_proba``,
prior_gmm.py:26-41).  scikit-learn is a third-party dependency that is not in
the reference tree (pinned ``scikit-learn~=1.0`` in setup.py:17; 1.9.0 is
installed here), so the published EM iteration is restated below and pinned
against live scikit-learn output in tests/golden (explicit initial parameters,
``tol=0``, fixed iteration counts, and the default stop rule).

EM iteration (sklearn/mixture/_base.py:263-278, _gaussian_mixture.py:312-313,
193-196, 364-370, 490-553):
  E: y = x P_k - mu_k P_k (P_k upper-triangular precision Cholesky = L_k^-T),
     logp_k = -(d log 2pi + |y|^2)/2 + sum log diag P_k + log pi_k,
     lse = logsumexp_k logp_k, log_resp = logp - lse, lower bound = mean(lse).
  M: r = exp(log_resp); nk = sum r + 10 eps; mu_k = r^T X / nk;
     Sigma_k = (r_k (X-mu_k))^T (X-mu_k) / nk + reg I; pi = nk / sum nk;
     L_k = chol(Sigma_k); P_k = (L_k^-1)^T.
  stop when |lb_i - lb_{i-1}| < tol.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
from scipy.linalg import solve_triangular


@dataclass
class GMMParams:
    weights: np.ndarray  # (K,)
    means: np.ndarray  # (K,d)
    covariances: np.ndarray  # (K,d,d)
    precisions_cholesky: np.ndarray  # (K,d,d) upper triangular
    lower_bound: float = -np.inf
    n_iter: int = 0
    converged: bool = False
    lower_bounds: list = field(default_factory=list)


def precision_cholesky(covariances):
    """P_k = (L_k^-1)^T with L_k the lower Cholesky factor of Sigma_k."""
    K, d, _ = covariances.shape
    P = np.empty_like(covariances)
    for k in range(K):
        try:
            L = np.linalg.cholesky(covariances[k])
        except np.linalg.LinAlgError:
            raise ValueError("ill-defined empirical covariance")
        P[k] = solve_triangular(L, np.eye(d), lower=True).T
    return P


def params_from_precisions(weights, means, precisions):
    """Parameters for an explicit (weights_init, means_init, precisions_init) start.

    sklearn derives P_k from the given precision matrices with
    ``_compute_precision_cholesky_from_precisions``: flip, Cholesky, flip, so that
    P_k is the *upper* factor with P_k P_k^T = precision_k.  Covariances are
    undefined until the first M-step; they are filled with inverse(precision).
    """
    K, d, _ = precisions.shape
    P = np.empty_like(precisions)
    cov = np.empty_like(precisions)
    for k in range(K):
        flipped = precisions[k][::-1, ::-1]
        P[k] = np.linalg.cholesky(flipped)[::-1, ::-1]
        cov[k] = np.linalg.inv(precisions[k])
    return GMMParams(np.asarray(weights, float).copy(), np.asarray(means, float).copy(), cov, P)


def weighted_log_prob(X, prm: GMMParams):
    n, d = X.shape
    K = prm.means.shape[0]
    out = np.empty((n, K))
    for k in range(K):
        P = prm.precisions_cholesky[k]
        y = X @ P - prm.means[k] @ P
        out[:, k] = -0.5 * (d * np.log(2 * np.pi) + np.sum(y * y, axis=1)) + np.sum(np.log(np.diag(P)))
    return out + np.log(prm.weights)


def gmm_e_step(X, prm: GMMParams):
    """-> (mean log-likelihood, log_resp (n,K))."""
    wlp = weighted_log_prob(X, prm)
    m = np.max(wlp, axis=1, keepdims=True)
    lse = m[:, 0] + np.log(np.sum(np.exp(wlp - m), axis=1))
    return float(np.mean(lse)), wlp - lse[:, None]


def gmm_m_step(X, resp, reg_covar: float = 1e-6) -> GMMParams:
    n, d = X.shape
    K = resp.shape[1]
    nk = resp.sum(axis=0) + 10 * np.finfo(np.float64).eps
    means = resp.T @ X / nk[:, None]
    cov = np.empty((K, d, d))
    for k in range(K):
        diff = X - means[k]
        cov[k] = (resp[:, k] * diff.T) @ diff / nk[k]
        cov[k].flat[:: d + 1] += reg_covar
    w = nk / nk.sum()  # sklearn 1.9 _m_step: weights_ = nk, then weights_ /= weights_.sum()
    return GMMParams(w, means, cov, precision_cholesky(cov))


def gmm_fit(X, init, tol: float = 1e-3, max_iter: int = 100, reg_covar: float = 1e-6, warm: bool = False) -> GMMParams:
    """EM loop of ``BaseMixture.fit_predict`` (sklearn/mixture/_base.py:241-305).

    ``init`` is either a GMMParams (explicit start / warm start) or an (n,)
    integer label array (hard one-hot responsibilities followed by one M-step,
    what the k-means initialisation of _base.py:119-128,159 produces).  With
    ``warm=True`` the first change is measured against ``init.lower_bound``
    (_base.py:257).
    """
    if isinstance(init, GMMParams):
        prm = init
        lb = init.lower_bound if warm else -np.inf
    else:
        labels = np.asarray(init)
        K = int(labels.max()) + 1
        resp = np.zeros((X.shape[0], K))
        resp[np.arange(X.shape[0]), labels] = 1
        prm = gmm_m_step(X, resp, reg_covar)
        # _initialize divides by n but does not renormalise (same value up to rounding)
        prm.weights = (resp.sum(axis=0) + 10 * np.finfo(np.float64).eps) / X.shape[0]
        lb = -np.inf
    lbs = []
    converged = False
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        prev = lb
        lb, log_resp = gmm_e_step(X, prm)
        prm = gmm_m_step(X, np.exp(log_resp), reg_covar)
        lbs.append(lb)
        if abs(lb - prev) < tol:
            converged = True
            break
    prm.lower_bound, prm.n_iter, prm.converged, prm.lower_bounds = lb, n_iter, converged, lbs
    return prm


def gmm_eval_prior(prm: GMMParams, n_pool: int, xux):
    """GMMPrior.eval, prior_gmm.py:26-41 -> (mu_p, Phi_p, N_p)."""
    _, log_resp = gmm_e_step(xux, prm)
    wts = np.mean(np.exp(log_resp), axis=0)
    mu_p = np.sum(wts[:, None] * prm.means, axis=0)
    Phi_p = np.sum(wts[:, None, None] * prm.covariances, axis=0)
    N_p = np.sum(wts * prm.weights) * n_pool
    return mu_p, Phi_p, N_p
