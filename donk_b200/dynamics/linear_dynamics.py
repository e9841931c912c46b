"""Time-varying linear-Gaussian dynamics and their least-squares fit
(donk/dynamics/linear_dynamics.py:16-64,137-189), fitted on the B200."""
from __future__ import annotations

import numpy as np

from donk_b200.dynamics.dynamics import DynamicsModel
from donk_b200.models import TimeVaryingLinearGaussian


class LinearDynamics(DynamicsModel, TimeVaryingLinearGaussian):
    """``x' ~ N(F_t [x;u] + f_t, covar_t)``."""

    def __init__(self, F: np.ndarray, f: np.ndarray, covar: np.ndarray):
        TimeVaryingLinearGaussian.__init__(self, F, f, covar)
        _, self.dX = f.shape
        self.dU = F.shape[2] - self.dX

    F = TimeVaryingLinearGaussian.coefficients
    f = TimeVaryingLinearGaussian.intercept

    def predict(self, x, u, t: int = None, noise=None):
        return TimeVaryingLinearGaussian.predict(self, np.concatenate([x, u], axis=-1), t, noise)

    def __str__(self) -> str:
        return f"LinearDynamics[T={self.T}, dX={self.dX}, dU={self.dU}]"

    def log_prob(self, X, U):
        """Log-probabilities of transitions under this model (linear_dynamics.py:54-64): ``X (..., T+1, dX)``,
        ``U (..., T, dU)`` -> ``(..., T)``, evaluated on the device (``donk_dynamics_log_prob_f64``: Cholesky factor of
        every ``covar_t``, one thread per transition).  A ``covar_t`` that is not positive definite raises
        ``numpy.linalg.LinAlgError``, as scipy's ``multivariate_normal.logpdf`` does for the reference.
        """
        from donk_b200 import batched

        X, U = np.asarray(X, dtype=np.float64), np.asarray(U, dtype=np.float64)
        lead = X.shape[:-2]
        Xb = X.reshape((1, -1) + X.shape[-2:])
        Ub = U.reshape((1, -1) + U.shape[-2:])
        out = batched.dynamics_log_prob(Xb, Ub, self.F[None], self.f[None], self.covar[None])
        return out[0].cpu().numpy().reshape(lead + (self.T,))


def fit_lr(X: np.ndarray, U: np.ndarray, prior=None, regularization: float = 1e-6,
           prior_weight: float = 1) -> LinearDynamics:
    """Fit dynamics by per-timestep linear regression under an optional NIW prior.

    Same signature and error behaviour as the reference (linear_dynamics.py:137-189): ``X (N,T+1,dX)``,
    ``U (N,T,dU)``, ``ValueError`` when ``N <= 1``.  A ``GMMPrior`` is evaluated inside the fit kernel (its mixture lives on the
    device); any other ``DynamicsPrior`` is evaluated by its own ``eval`` and enters the kernel as explicit NIW parameters.
    """
    from donk_b200 import batched
    from donk_b200.dynamics.prior import GMMPrior

    X, U = np.asarray(X, dtype=np.float64), np.asarray(U, dtype=np.float64)
    N = X.shape[0]
    if N <= 1:
        raise ValueError(f"Cannot fit dynamics to {N} sample(s)")
    if prior is None or isinstance(prior, GMMPrior):
        gmm = None if prior is None else prior._gmm
        F, f, covar, info = batched.fit_lr(X[None], U[None], gmm, regularization, prior_weight)
    else:
        # any other DynamicsPrior (prior.py:62-82): its eval() is user code and runs on the host, one call per time
        # step as in the reference (linear_dynamics.py:171); the regression itself still runs on the device
        T = U.shape[1]
        niw = [prior.eval(np.concatenate([X[:, t], U[:, t], X[:, t + 1]], axis=1)) for t in range(T)]
        F, f, covar, info = batched.fit_lr_niw(
            X[None], U[None], np.stack([q.mu0 for q in niw])[None], np.stack([q.Phi for q in niw])[None],
            np.array([[float(q.N_mean) for q in niw]]), np.array([[float(q.N_covar) for q in niw]]),
            regularization, prior_weight)
    if int(info.sum()) != 0:
        raise np.linalg.LinAlgError("Singular matrix")  # numpy.linalg.solve, linear_dynamics.py:184
    return LinearDynamics(F[0].cpu().numpy(), f[0].cpu().numpy(), covar[0].cpu().numpy())
