"""Oracle: time-varying linear-Gaussian dynamics fit (numpy restatement).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates
  fit_lr                       donk/dynamics/linear_dynamics.py:137-189
  NormalInverseWishart         donk/dynamics/prior/prior.py:9-59
  to_transitions               donk/dynamics/utils.py:4-17
  StateDistribution.fit        donk/samples/state_dist.py:27-37
"""
from __future__ import annotations

import numpy as np


def to_transitions(X, U):
    """(N,T+1,dX),(N,T,dU) -> (N*T, dX+dU+dX); row n*T+t = [x_t|u_t|x_{t+1}] (utils.py:4-17)."""
    N, T, _ = U.shape
    return np.concatenate([X[:, :-1], U, X[:, 1:]], axis=-1).reshape(N * T, -1)


def state_fit(X0):
    """Mean and *unbiased* covariance of the initial states (state_dist.py:27-37)."""
    return np.mean(X0, axis=0), np.cov(X0, rowvar=False)


def niw_posterior(mu0, Phi, N_mean, N_covar, emp_mean, emp_covar, N):
    """NIW conjugate update, prior.py:18-36."""
    dm = emp_mean - mu0
    return (
        (N_mean * mu0 + N * emp_mean) / (N_mean + N),
        Phi + N * emp_covar + N_mean * N / (N_mean + N) * np.outer(dm, dm),
        N_mean + N,
        N_covar + N,
    )


def niw_map_covar(Phi, N_covar, d):
    """prior.py:42-45."""
    return Phi / (N_covar + d + 1)


def fit_lr(X, U, prior=None, regularization: float = 1e-6, prior_weight: float = 1.0):
    """Per-timestep conditioning of the joint Gaussian of [x_t;u_t;x_{t+1}].

    X (N,T+1,dX), U (N,T,dU) -> F (T,dX,p), f (T,dX), covar (T,dX,dX).
    ``prior`` is None or a callable ``xux (N,d) -> (mu_p, Phi_p, N_p)`` giving the
    GMM-mixed prior mean, covariance and strength (prior_gmm.py:26-41; see
    ``oracle.gmm.gmm_eval_prior``), or returning an object with ``mu0, Phi, N_mean, N_covar`` (an explicit NIW).
    """
    N, _, dX = X.shape
    _, T, dU = U.shape
    p = dX + dU
    d = p + dX
    if N <= 1:
        raise ValueError(f"Cannot fit dynamics to {N} sample(s)")  # linear_dynamics.py:154-155
    F = np.empty((T, dX, p))
    f = np.empty((T, dX))
    covar = np.empty((T, dX, dX))
    ip = np.arange(p)
    for t in range(T):
        xux = np.concatenate([X[:, t], U[:, t], X[:, t + 1]], axis=1)
        mu = xux.mean(axis=0)
        diff = xux - mu
        sigma = diff.T @ diff / N  # biased, linear_dynamics.py:165
        if prior is not None:
            pr = prior(xux)
            if hasattr(pr, "mu0"):  # an explicit NIW, as any DynamicsPrior.eval returns (prior.py:62-82)
                m0, Ph, Nm, Nc = pr.mu0, pr.Phi, pr.N_mean, pr.N_covar
            else:
                mu_p, Phi_p, N_p = pr
                # non_informative_prior(d).posterior(mu_p, Phi_p, N_p)  (prior_gmm.py:41)
                m0, Ph, Nm, Nc = niw_posterior(np.zeros(d), np.zeros((d, d)), 0, -(1 + d), mu_p, Phi_p, N_p)
            # .posterior(empmu, empsig, N / prior_weight).map_covar()  (linear_dynamics.py:172-175)
            _, Ph, _, Nc = niw_posterior(m0, Ph, Nm, Nc, mu, sigma, N / prior_weight)
            sigma = niw_map_covar(Ph, Nc, d)
        lo = np.min(np.linalg.eigvalsh(sigma[:p, :p]))  # linear_dynamics.py:178
        if lo < regularization:
            sigma[ip, ip] += regularization - lo  # only the p x p block's diagonal
        Ft = np.linalg.solve(sigma[:p, :p], sigma[:p, p:]).T  # linear_dynamics.py:184
        F[t] = Ft
        f[t] = mu[p:] - Ft @ mu[:p]
        covar[t] = sigma[p:, p:] - Ft @ sigma[:p, :p] @ Ft.T
    covar = (covar + np.swapaxes(covar, 1, 2)) / 2  # linear_dynamics.py:188
    return F, f, covar
