"""Deterministic synthetic GPS problems.

Numpy / torch only; shared by ``tests/`` and ``bench.py``.  Not part of the
reference: the reference ships no data generators for its hot path beyond the
tiny random helpers of ``tests/utils.py:13-31``.

A *condition* is one trajectory-optimisation problem: stable linear system
``x' = A x + B u + noise`` rolled out under a random time-varying
linear-Gaussian policy, with an L2 cost (``quadratic_costs.py:121-139`` with
target 0 and weights ``[1]*dX + [1e-2]*dU``).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Condition:
    X: np.ndarray  # (N,T+1,dX) roll-out states
    U: np.ndarray  # (N,T,dU) roll-out actions
    prev_K: np.ndarray  # (T,dU,dX)
    prev_k: np.ndarray  # (T,dU)
    prev_covar: np.ndarray  # (T,dU,dU)
    C: np.ndarray  # (T+1,p,p)
    c: np.ndarray  # (T+1,p)
    cc: np.ndarray  # (T+1,)


def l2_costs(T: int, dX: int, dU: int):
    p = dX + dU
    w = np.concatenate([np.ones(dX), 1e-2 * np.ones(dU)])
    C = np.broadcast_to(np.diag(w), (T + 1, p, p)).copy()
    return C, np.zeros((T + 1, p)), np.zeros(T + 1)


def condition(config: int, i: int, T: int, dX: int, dU: int, N: int) -> Condition:
    """Condition ``i`` of benchmark config ``config`` (seed ``1000*config + i``)."""
    rng = np.random.default_rng(1000 * config + i)
    Q, _ = np.linalg.qr(rng.standard_normal((dX, dX)))
    A = 0.97 * Q
    B = 0.1 * rng.standard_normal((dX, dU))
    K = 0.05 * rng.standard_normal((T, dU, dX))
    k = 0.1 * rng.standard_normal((T, dU))
    covar = np.broadcast_to(0.5 * np.eye(dU), (T, dU, dU)).copy()
    X = np.empty((N, T + 1, dX))
    U = np.empty((N, T, dU))
    X[:, 0] = rng.standard_normal((N, dX))
    for t in range(T):
        U[:, t] = X[:, t] @ K[t].T + k[t] + np.sqrt(0.5) * rng.standard_normal((N, dU))
        X[:, t + 1] = X[:, t] @ A.T + U[:, t] @ B.T + 0.05 * rng.standard_normal((N, dX))
    C, c, cc = l2_costs(T, dX, dU)
    return Condition(X, U, K, k, covar, C, c, cc)


def rollout_batch_torch(B: int, T: int, dX: int, dU: int, N: int, seed: int, device):
    """Batched roll-outs for ``B`` conditions, generated on ``device`` with torch.

    Same construction as :func:`condition` but one torch generator for the
    whole batch (the per-condition numpy streams are too slow for B=4096).
    Returns dict of float64 tensors: X (B,N,T+1,dX), U (B,N,T,dU), prev_K
    (B,T,dU,dX), prev_k (B,T,dU), prev_covar (B,T,dU,dU).
    """
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f64 = dict(dtype=torch.float64, device=device, generator=g)
    # orthogonal factors on the host: the CUDA QR of a batch of small matrices is a loop of ~8 tiny launches per matrix
    # (33 k launches at B = 4096), which takes seconds plain and hours under a profiler
    Q = torch.linalg.qr(torch.randn(B, dX, dX, **f64).cpu())[0].to(device)
    A = 0.97 * Q
    Bm = 0.1 * torch.randn(B, dX, dU, **f64)
    K = 0.05 * torch.randn(B, T, dU, dX, **f64)
    k = 0.1 * torch.randn(B, T, dU, **f64)
    covar = (0.5 * torch.eye(dU, dtype=torch.float64, device=device)).expand(B, T, dU, dU).contiguous()
    X = torch.empty(B, N, T + 1, dX, dtype=torch.float64, device=device)
    U = torch.empty(B, N, T, dU, dtype=torch.float64, device=device)
    X[:, :, 0] = torch.randn(B, N, dX, **f64)
    s = float(np.sqrt(0.5))
    for t in range(T):
        u = torch.einsum("bnx,bux->bnu", X[:, :, t], K[:, t]) + k[:, t, None, :] + s * torch.randn(B, N, dU, **f64)
        U[:, :, t] = u
        X[:, :, t + 1] = (
            torch.einsum("bnx,byx->bny", X[:, :, t], A)
            + torch.einsum("bnu,byu->bny", u, Bm)
            + 0.05 * torch.randn(B, N, dX, **f64)
        )
    return dict(X=X, U=U, prev_K=K, prev_k=k, prev_covar=covar)


def gmm_pool(n: int, d: int, K: int, seed: int):
    """Config-4 style transition pool: rows = centers[label] + N(0,I).

    Returns (X (n,d), init_weights (K,), init_means (K,d), init_precisions (K,d,d)).
    """
    rng = np.random.default_rng(seed)
    centers = 3.0 * rng.standard_normal((K, d))
    means0 = centers + 0.5 * rng.standard_normal((K, d))
    label = rng.integers(0, K, size=n)
    X = centers[label] + rng.standard_normal((n, d))
    prec0 = np.broadcast_to(np.eye(d), (K, d, d)).copy()
    return X, np.full(K, 1.0 / K), means0, prec0


def gmm_pool_torch(n: int, d: int, K: int, seed: int, device):
    """Same distribution as :func:`gmm_pool`, generated on ``device`` (n up to 1e7)."""
    import torch

    rng = np.random.default_rng(seed)
    centers = 3.0 * rng.standard_normal((K, d))
    means0 = centers + 0.5 * rng.standard_normal((K, d))
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    label = torch.randint(0, K, (n,), device=device, generator=g)
    X = torch.randn(n, d, dtype=torch.float64, device=device, generator=g)
    X += torch.as_tensor(centers, device=device)[label]
    prec0 = np.broadcast_to(np.eye(d), (K, d, d)).copy()
    return X, np.full(K, 1.0 / K), means0, prec0


def gmm_prior(dX: int, dU: int, K: int, seed: int):
    """A K-cluster mixture over transitions ``[x_t | u_t | x_{t+1}]`` with well-conditioned random covariances: the
    dynamics prior of the benchmark configurations that time ``fit_lr`` under a ``GMMPrior`` (both arms of bench.py are
    handed these very parameters).  Returns ``(weights (K,), means (K,d), covariances (K,d,d))``."""
    d = 2 * dX + dU
    rng = np.random.default_rng(seed)
    w = rng.uniform(0.5, 1.5, K)
    w /= w.sum()
    means = 0.3 * rng.standard_normal((K, d))
    A = rng.standard_normal((K, d, d)) / np.sqrt(d)
    cov = A @ np.swapaxes(A, -1, -2) + 0.5 * np.eye(d)
    return w, means, cov


def precision_cholesky(covariances):
    """``precisions_cholesky_`` as scikit-learn defines it: upper-triangular ``P_k`` with ``P_k P_k^T = Sigma_k^-1``
    (the transpose of the inverse of the lower Cholesky factor of ``Sigma_k``)."""
    from scipy.linalg import solve_triangular

    cov = np.asarray(covariances, dtype=np.float64)
    out = np.empty_like(cov)
    eye = np.eye(cov.shape[-1])
    for k in range(cov.shape[0]):
        out[k] = solve_triangular(np.linalg.cholesky(cov[k]), eye, lower=True).T
    return out
