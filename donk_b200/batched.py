"""Batched, device-resident entry points of the hot path (torch tensors in, torch tensors out).

The reference API is single-problem (one condition, one eta per call); these are the additive
batched forms the B200 path is built around — leading ``(B,)`` conditions and ``(B, E)`` eta
candidates.  The reference-shaped classes in ``donk_b200.traj_opt``,
``donk_b200.dynamics`` are thin B=1 views over this module.

Everything here calls the CUDA library through ``donk_b200._lib.native()``; there is no CPU path.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from donk_b200 import _lib
from donk_b200._lib import ILQR_BACKWARD, ILQR_COST, ILQR_FORWARD, ILQR_KL, DonkError


def _dev(x, dtype=torch.float64):
    return _lib.as_device(x, dtype)


def spd_factor(A: torch.Tensor, want_inv=True, want_chol=False, want_hld=True):
    """Cholesky-based inverse / factor / half log-determinant of a batch of SPD matrices.

    ``A (..., d, d)``.  Raises ``numpy.linalg.LinAlgError`` if any matrix is not positive definite
    (as ``numpy.linalg.cholesky`` does in donk/utils/batched.py:48).
    """
    nat = _lib.native()
    A = _dev(A)
    d = A.shape[-1]
    n = A.numel() // (d * d)
    inv = torch.empty_like(A) if want_inv else None
    chol = torch.empty_like(A) if want_chol else None
    hld = torch.empty(A.shape[:-2], dtype=A.dtype, device=A.device) if want_hld else None
    info = torch.zeros(max(n, 1), dtype=torch.int32, device=A.device)
    nat.call("spd_factor_f64", n, d, A, inv, chol, hld, info)
    if int(info.sum().item()) != 0:
        raise np.linalg.LinAlgError("Matrix is not positive definite")
    return inv, chol, hld


def extended_costs_kl(K: torch.Tensor, k: torch.Tensor, inv_covar: torch.Tensor):
    """``(..., T, dU, dX), (..., T, dU), (..., T, dU, dU) -> C_kl (..., T, p, p), c_kl (..., T, p)``."""
    nat = _lib.native()
    K, k, inv_covar = _dev(K), _dev(k), _dev(inv_covar)
    dU, dX = K.shape[-2:]
    p = dX + dU
    n = K.numel() // (dU * dX)
    C = torch.empty(K.shape[:-2] + (p, p), dtype=K.dtype, device=K.device)
    c = torch.empty(K.shape[:-2] + (p,), dtype=K.dtype, device=K.device)
    nat.call("extended_costs_kl_f64", n, dX, dU, K, k, inv_covar, C, c)
    return C, c


def kl_divergence_action(X, K, k, covar, prev_K, prev_k, prev_covar, prev_inv_covar=None):
    """Batched ``kl_divergence_action`` for given state means ``X (B,T,dX)`` -> ``kl (B,)``."""
    nat = _lib.native()
    X, K, k, covar = _dev(X), _dev(K), _dev(k), _dev(covar)
    prev_K, prev_k, prev_covar = _dev(prev_K), _dev(prev_k), _dev(prev_covar)
    B, T, dX = X.shape
    dU = K.shape[-2]
    _, _, hld = spd_factor(covar, want_inv=False)
    if prev_inv_covar is None:
        prev_inv, _, phld = spd_factor(prev_covar)
    else:
        prev_inv = _dev(prev_inv_covar)
        _, _, phld = spd_factor(prev_covar, want_inv=False)
    kl = torch.empty(B, dtype=X.dtype, device=X.device)
    nat.call("kl_divergence_action_f64", B, T, dX, dU, X, K, k, covar, hld, prev_K, prev_k, prev_inv, phld, kl)
    return kl


@dataclass
class BatchedStepResult:
    """Result of one batched iLQR step; leading dims ``(B, E)``."""

    eta: torch.Tensor  # (B,E)
    K: torch.Tensor  # (B,E,T,dU,dX)
    k: torch.Tensor  # (B,E,T,dU)
    covar: torch.Tensor  # (B,E,T,dU,dU)
    inv_covar: torch.Tensor  # (B,E,T,dU,dU)
    kl_div: torch.Tensor  # (B,E)
    expected_costs: torch.Tensor  # (B,E)
    info: torch.Tensor  # (B,E) int32
    traj_mean: torch.Tensor | None = None  # (B,E,T+1,p)
    traj_covar: torch.Tensor | None = None  # (B,E,T+1,p,p)


def lqr_backward(F, f, C, c, gamma: float = 1.0):
    """Batched ``backward`` (lqg.py:126-182): ``F (B,T,dX,p) ... -> K, k, covar, inv_covar, info``."""
    nat = _lib.native()
    F, f, C, c = _dev(F), _dev(f), _dev(C), _dev(c)
    B, T, dX, p = F.shape
    dU = p - dX
    o = dict(dtype=F.dtype, device=F.device)
    K = torch.empty(B, 1, T, dU, dX, **o)
    k = torch.empty(B, 1, T, dU, **o)
    cov = torch.empty(B, 1, T, dU, dU, **o)
    inv = torch.empty(B, 1, T, dU, dU, **o)
    info = torch.zeros(B, 1, dtype=torch.int32, device=F.device)
    nat.call("ilqr_step_f64", B, 1, T, dX, dU, ILQR_BACKWARD, F, f, None, C, c, None, None, None, None, None, None,
             None, None, None, None, float(gamma), 0.0, None, K, k, cov, inv, None, None, None, None, info)
    return K[:, 0], k[:, 0], cov[:, 0], inv[:, 0], info[:, 0]


def lqr_forward(F, f, dyn_covar, K, k, pol_covar, x0_mean, x0_covar, regularization: float = 1e-6,
                C=None, c=None, cc=None):
    """Batched ``forward`` (lqg.py:185-244) -> ``mean (B,T+1,p), covar (B,T+1,p,p)[, cost (B,)]``.

    With costs ``C, c, cc`` the summed ``expected_costs`` (quadratic_costs.py:106-119) is returned too.
    """
    nat = _lib.native()
    F, f, dyn_covar = _dev(F), _dev(f), _dev(dyn_covar)
    K, k, pol_covar = _dev(K).unsqueeze(1).contiguous(), _dev(k).unsqueeze(1).contiguous(), _dev(pol_covar).unsqueeze(1).contiguous()
    x0_mean, x0_covar = _dev(x0_mean), _dev(x0_covar)
    B, T, dX, p = F.shape
    dU = p - dX
    o = dict(dtype=F.dtype, device=F.device)
    mean = torch.empty(B, 1, T + 1, p, **o)
    covar = torch.empty(B, 1, T + 1, p, p, **o)
    info = torch.zeros(B, 1, dtype=torch.int32, device=F.device)
    flags = ILQR_FORWARD
    cost = None
    if C is not None:
        C, c, cc = _dev(C), _dev(c), _dev(cc)
        cost = torch.empty(B, 1, **o)
        flags |= ILQR_COST
    nat.call("ilqr_step_f64", B, 1, T, dX, dU, flags, F, f, dyn_covar, C, c, cc, None, None, None, None, None, None,
             x0_mean, x0_covar, None, 1.0, float(regularization), None, K, k, pol_covar, pol_covar, None, cost, mean,
             covar, info)
    if cost is not None:
        return mean[:, 0], covar[:, 0], cost[:, 0]
    return mean[:, 0], covar[:, 0]


class BatchedILQR:
    """``ILQR(dynamics, prev_pol, costs, initial_state)`` (lqg.py:29-123) for B conditions at once.

    All arguments carry a leading ``B``: dynamics ``F (B,T,dX,p), f (B,T,dX), dyn_covar (B,T,dX,dX)``;
    previous policy ``prev_K (B,T,dU,dX), prev_k (B,T,dU)`` and ``prev_covar`` and/or ``prev_inv_covar
    (B,T,dU,dU)``; costs ``C (B,T+1,p,p), c (B,T+1,p), cc (B,T+1)``; ``x0_mean (B,dX), x0_covar (B,dX,dX)``.
    """

    def __init__(self, F, f, dyn_covar, prev_K, prev_k, C, c, cc, x0_mean, x0_covar, prev_covar=None,
                 prev_inv_covar=None):
        self.F, self.f, self.dyn_covar = _dev(F), _dev(f), _dev(dyn_covar)
        self.prev_K, self.prev_k = _dev(prev_K), _dev(prev_k)
        self.C, self.c, self.cc = _dev(C), _dev(c), _dev(cc)
        self.x0_mean, self.x0_covar = _dev(x0_mean), _dev(x0_covar)
        self.B, self.T, self.dX, p = self.F.shape
        self.dU = p - self.dX
        if prev_covar is None and prev_inv_covar is None:
            raise ValueError("Must provide covar or inv_covar.")
        # lazy covar / chol_covar / inv_covar of the previous policy (donk/models/tvlg.py:62-99)
        if prev_covar is not None:
            self.prev_covar = _dev(prev_covar)
            inv, _, self.prev_hld = spd_factor(self.prev_covar)
            self.prev_inv_covar = inv if prev_inv_covar is None else _dev(prev_inv_covar)
        else:
            self.prev_inv_covar = _dev(prev_inv_covar)
            self.prev_covar, _, _ = spd_factor(self.prev_inv_covar, want_hld=False)
            _, _, self.prev_hld = spd_factor(self.prev_covar, want_inv=False)
        self.C_kl, self.c_kl = extended_costs_kl(self.prev_K, self.prev_k, self.prev_inv_covar)
        self._ws = {}

    # -- workspace ---------------------------------------------------------------------------
    def _buffers(self, E: int, want_traj: bool):
        key = (E, want_traj)
        ws = self._ws.get(key)
        if ws is None:
            B, T, dX, dU = self.B, self.T, self.dX, self.dU
            p = dX + dU
            o = dict(dtype=self.F.dtype, device=self.F.device)
            ws = dict(
                K=torch.empty(B, E, T, dU, dX, **o), k=torch.empty(B, E, T, dU, **o),
                covar=torch.empty(B, E, T, dU, dU, **o), inv_covar=torch.empty(B, E, T, dU, dU, **o),
                kl=torch.zeros(B, E, **o), cost=torch.zeros(B, E, **o),
                info=torch.zeros(B, E, dtype=torch.int32, device=self.F.device),
                tmean=torch.empty(B, E, T + 1, p, **o) if want_traj else None,
                tcov=torch.empty(B, E, T + 1, p, p, **o) if want_traj else None,
            )
            self._ws = {key: ws}  # keep one set only: the policy buffers dominate memory
        return ws

    def step(self, eta, active: torch.Tensor | None = None, want_traj: bool = False) -> BatchedStepResult:
        """``ILQR.step`` (lqg.py:55-75) for ``eta (B,E)`` (or ``(B,)``, or a scalar broadcast to all)."""
        nat = _lib.native()
        eta = torch.as_tensor(eta, dtype=self.F.dtype).to(self.F.device)
        if eta.ndim == 0:
            eta = eta.expand(self.B, 1)
        elif eta.ndim == 1:
            eta = eta[:, None]
        eta = eta.contiguous()
        E = eta.shape[1]
        ws = self._buffers(E, want_traj)
        nat.call("ilqr_step_f64", self.B, E, self.T, self.dX, self.dU,
                 ILQR_BACKWARD | ILQR_FORWARD | ILQR_KL | ILQR_COST,
                 self.F, self.f, self.dyn_covar, self.C, self.c, self.cc, self.C_kl, self.c_kl,
                 self.prev_K, self.prev_k, self.prev_inv_covar, self.prev_hld, self.x0_mean, self.x0_covar,
                 eta, 1.0, 1e-6, active, ws["K"], ws["k"], ws["covar"], ws["inv_covar"], ws["kl"], ws["cost"],
                 ws["tmean"], ws["tcov"], ws["info"])
        return BatchedStepResult(eta, ws["K"], ws["k"], ws["covar"], ws["inv_covar"], ws["kl"], ws["cost"],
                                 ws["info"], ws["tmean"], ws["tcov"])

    def step_f32(self, eta) -> BatchedStepResult:
        """``ILQR.step`` in single precision (``donk_ilqr_step_f32``): same kernel, float storage and FFMA math.

        Inputs are cast once and cached.  Stated tolerance against the fp64 path / the reference for
        well-conditioned problems (tests/test_ilqr_kernels.py::test_step_f32_tolerance): gains ``K`` within
        2e-3 max-norm relative, KL and expected cost within 5e-3 relative, for eta in [1e-2, 1e2].  At the extremes
        of the eta range ``C/eta + C_kl`` loses ``C_kl`` (or ``C``) entirely in fp32 — use fp64 there.
        """
        nat = _lib.native()
        f32 = torch.float32
        if getattr(self, "_f32", None) is None:
            self._f32 = {k: getattr(self, k).to(f32).contiguous() for k in
                         ("F", "f", "dyn_covar", "C", "c", "cc", "C_kl", "c_kl", "prev_K", "prev_k", "prev_inv_covar",
                          "prev_hld", "x0_mean", "x0_covar")}
        d = self._f32
        eta = torch.as_tensor(eta, dtype=f32).to(self.F.device)
        if eta.ndim == 0:
            eta = eta.expand(self.B, 1)
        elif eta.ndim == 1:
            eta = eta[:, None]
        eta = eta.contiguous()
        E = eta.shape[1]
        B, T, dX, dU = self.B, self.T, self.dX, self.dU
        o = dict(dtype=f32, device=self.F.device)
        K, k = torch.empty(B, E, T, dU, dX, **o), torch.empty(B, E, T, dU, **o)
        cov, inv = torch.empty(B, E, T, dU, dU, **o), torch.empty(B, E, T, dU, dU, **o)
        kl, cost = torch.zeros(B, E, **o), torch.zeros(B, E, **o)
        info = torch.zeros(B, E, dtype=torch.int32, device=self.F.device)
        nat.call("ilqr_step_f32", B, E, T, dX, dU, ILQR_BACKWARD | ILQR_FORWARD | ILQR_KL | ILQR_COST,
                 d["F"], d["f"], d["dyn_covar"], d["C"], d["c"], d["cc"], d["C_kl"], d["c_kl"], d["prev_K"], d["prev_k"],
                 d["prev_inv_covar"], d["prev_hld"], d["x0_mean"], d["x0_covar"], eta, 1.0, 1e-6, None, K, k, cov, inv,
                 kl, cost, None, None, info)
        return BatchedStepResult(eta, K, k, cov, inv, kl, cost, info)

    def sample_surface(self, min_eta: float = 1e-6, max_eta: float = 1e16, N: int = 100) -> BatchedStepResult:
        """``ILQR.sample_surface`` (lqg.py:77-88): N log-spaced eta candidates for every condition."""
        etas = np.logspace(np.log10(min_eta), np.log10(max_eta), num=N)
        return self.step(torch.as_tensor(etas).expand(self.B, N))

    def optimize(self, kl_step, min_eta: float = 1e-6, max_eta: float = 1e16, rtol: float = 1e-2,
                 want_traj: bool = False, raise_on_failure: bool = True):
        """``ILQR.optimize`` (lqg.py:90-123) for all B conditions in lock-step, as one CUDA-graph launch.

        The whole search — the check at ``min_eta``, the check at ``max_eta``, brentq's two bracket evaluations, the
        Brent rounds until the last condition has converged, the step at the root — runs on the device
        (``donk_ilqr_optimize_f64``: a WHILE graph node around the Brent rounds); the host reads three counters once.

        ``kl_step`` scalar or ``(B,)``.  Returns ``(BatchedStepResult with E=1, status (B,) int32, n_steps)``; status
        0 = root found, 1 = constraint inactive at ``min_eta`` (lqg.py:108-109), 2 = ``max_eta`` too low (the
        reference raises ValueError, lqg.py:112-113), 3 = Brent failed (scipy raises ValueError / RuntimeError).
        ``n_steps`` counts the batched ``step`` evaluations the reference's control flow needs for the slowest
        condition.  Results live in this object's workspace and are valid until the next ``step`` / ``optimize``.
        """
        nat = _lib.native()
        dev, dt, B = self.F.device, self.F.dtype, self.B
        ws = self._buffers(1, want_traj)
        opt = getattr(self, "_opt", None)
        if opt is None:
            i32 = dict(dtype=torch.int32, device=dev)
            opt = self._opt = dict(
                kl_step=torch.empty(B, dtype=dt, device=dev), eta=torch.empty(B, 1, dtype=dt, device=dev),
                active=torch.zeros(B, **i32), status=torch.zeros(B, **i32), counters=torch.zeros(3, **i32),
                work=torch.empty(nat.size("ilqr_optimize_workspace_bytes", B) // 8 + 1, dtype=dt, device=dev))
        opt["kl_step"].copy_(torch.as_tensor(kl_step, dtype=dt).to(dev).expand(B))
        # brentq evaluates f(log min_eta), f(log max_eta) itself (exp(log(x)) may differ from x by an ulp)
        xa, xb = float(np.log(min_eta)), float(np.log(max_eta))
        nat.call("ilqr_optimize_f64", B, self.T, self.dX, self.dU, self.F, self.f, self.dyn_covar, self.C, self.c,
                 self.cc, self.C_kl, self.c_kl, self.prev_K, self.prev_k, self.prev_inv_covar, self.prev_hld,
                 self.x0_mean, self.x0_covar, opt["kl_step"], float(min_eta), float(max_eta), float(np.exp(xa)),
                 float(np.exp(xb)), xa, xb, float(rtol), 100, opt["eta"], opt["active"], ws["K"], ws["k"],
                 ws["covar"], ws["inv_covar"], ws["kl"], ws["cost"], ws["tmean"], ws["tcov"], ws["info"],
                 opt["status"], opt["counters"], opt["work"], opt["work"].numel() * opt["work"].element_size())
        rounds, err, searched = opt["counters"].tolist()  # the one device->host read of the optimisation
        if err & 1:
            raise np.linalg.LinAlgError("Quu is not positive definite")
        if raise_on_failure:
            if err & 2:
                raise ValueError(f"max_eta eta to low ({max_eta})")
            if err & 4:
                raise ValueError("f(a) and f(b) must have different signs")
            if err & 8:
                raise RuntimeError("Failed to converge after 100 iterations.")
        n_steps = 2 + (3 + rounds if searched > 0 else 0)
        result = BatchedStepResult(opt["eta"], ws["K"], ws["k"], ws["covar"], ws["inv_covar"], ws["kl"], ws["cost"],
                                   ws["info"], ws["tmean"], ws["tcov"])
        return result, opt["status"], n_steps


# =====================================================================================================
# Dynamics fit
# =====================================================================================================
def to_transitions(X, U) -> torch.Tensor:
    """``(N,T+1,dX), (N,T,dU) -> (N*T, 2dX+dU)`` rows ``[x_t | u_t | x_{t+1}]`` (donk/dynamics/utils.py:4-17)."""
    nat = _lib.native()
    X, U = _dev(X), _dev(U)
    N, T, dU = U.shape
    dX = X.shape[-1]
    out = torch.empty(N * T, 2 * dX + dU, dtype=X.dtype, device=X.device)
    nat.call("to_transitions_f64", N, T, dX, dU, X, U, out)
    return out


def state_fit(X0):
    """``StateDistribution.fit`` for ``X0 (B,N,dX)``: mean ``(B,dX)`` and unbiased covariance ``(B,dX,dX)``."""
    nat = _lib.native()
    X0 = _dev(X0)
    B, N, dX = X0.shape
    mean = torch.empty(B, dX, dtype=X0.dtype, device=X0.device)
    covar = torch.empty(B, dX, dX, dtype=X0.dtype, device=X0.device)
    nat.call("state_fit_f64", B, N, dX, X0, dX, N * dX, mean, covar)
    return mean, covar


def state_fit_from_rollouts(X):
    """Initial-state distribution straight from roll-outs ``X (B,N,T+1,dX)`` (no gather copy)."""
    nat = _lib.native()
    X = _dev(X)
    B, N, T1, dX = X.shape
    mean = torch.empty(B, dX, dtype=X.dtype, device=X.device)
    covar = torch.empty(B, dX, dX, dtype=X.dtype, device=X.device)
    nat.call("state_fit_f64", B, N, dX, X, T1 * dX, N * T1 * dX, mean, covar)
    return mean, covar


def fit_lr(X, U, prior: "DeviceGMM | None" = None, regularization: float = 1e-6, prior_weight: float = 1.0):
    """Batched ``fit_lr`` (donk/dynamics/linear_dynamics.py:137-189).

    ``X (B,N,T+1,dX), U (B,N,T,dU) -> F (B,T,dX,p), f (B,T,dX), covar (B,T,dX,dX), info (B,T)``.
    ``prior`` is a fitted :class:`DeviceGMM` (the GMM behind ``GMMPrior``) or ``None``.
    """
    nat = _lib.native()
    X, U = _dev(X), _dev(U)
    B, N, T1, dX = X.shape
    _, _, T, dU = U.shape
    if T1 != T + 1 or U.shape[0] != B or U.shape[1] != N:
        raise ValueError(f"inconsistent roll-out shapes {tuple(X.shape)} / {tuple(U.shape)}")
    if N <= 1:
        raise ValueError(f"Cannot fit dynamics to {N} sample(s)")  # linear_dynamics.py:154-155
    p = dX + dU
    o = dict(dtype=X.dtype, device=X.device)
    F = torch.empty(B, T, dX, p, **o)
    f = torch.empty(B, T, dX, **o)
    covar = torch.empty(B, T, dX, dX, **o)
    info = torch.zeros(B, T, dtype=torch.int32, device=X.device)
    if prior is None:
        nat.call("fit_lr_f64", B, N, T, dX, dU, X, U, 0, None, None, None, None, 0.0, 1.0, float(regularization),
                 F, f, covar, info)
    else:
        if prior.d != 2 * dX + dU:
            raise ValueError(f"prior dimension {prior.d} != {2 * dX + dU}")
        nat.call("fit_lr_f64", B, N, T, dX, dU, X, U, prior.K, prior.weights, prior.means, prior.covariances,
                 prior.precisions_cholesky, float(prior.N), float(prior_weight), float(regularization), F, f, covar,
                 info)
    return F, f, covar, info


def fit_lr_f32(X, U, prior: "DeviceGMM | None" = None, regularization: float = 1e-6, prior_weight: float = 1.0):
    """:func:`fit_lr` in single precision (``donk_fit_lr_f32``): float32 tensors in and out, half the HBM traffic.

    Stated tolerance against the fp64 path for well-conditioned data (N > 2(dX+dU)): ``F`` and ``f`` within 2e-3,
    ``covar`` within 5e-3, max-norm relative (tests/test_fit_gmm_kernels.py::test_fit_lr_f32_tolerance).  The joint
    covariance loses ~7 digits to the subtraction of the means; ill-conditioned fits (N close to dX+dU, where the
    eigenvalue lift decides the result) belong on the fp64 path.
    """
    nat = _lib.native()
    f32 = torch.float32
    X, U = _dev(X, f32), _dev(U, f32)
    B, N, T1, dX = X.shape
    _, _, T, dU = U.shape
    if T1 != T + 1 or U.shape[0] != B or U.shape[1] != N:
        raise ValueError(f"inconsistent roll-out shapes {tuple(X.shape)} / {tuple(U.shape)}")
    if N <= 1:
        raise ValueError(f"Cannot fit dynamics to {N} sample(s)")
    p = dX + dU
    o = dict(dtype=f32, device=X.device)
    F, f, covar = torch.empty(B, T, dX, p, **o), torch.empty(B, T, dX, **o), torch.empty(B, T, dX, dX, **o)
    info = torch.zeros(B, T, dtype=torch.int32, device=X.device)
    if prior is None:
        nat.call("fit_lr_f32", B, N, T, dX, dU, X, U, 0, None, None, None, None, 0.0, 1.0, float(regularization),
                 F, f, covar, info)
    else:
        if prior.d != 2 * dX + dU:
            raise ValueError(f"prior dimension {prior.d} != {2 * dX + dU}")
        g = [t.to(f32).contiguous() for t in (prior.weights, prior.means, prior.covariances, prior.precisions_cholesky)]
        nat.call("fit_lr_f32", B, N, T, dX, dU, X, U, prior.K, g[0], g[1], g[2], g[3], float(prior.N),
                 float(prior_weight), float(regularization), F, f, covar, info)
    return F, f, covar, info


def fit_lr_niw(X, U, mu0, Phi, N_mean, N_covar, regularization: float = 1e-6, prior_weight: float = 1.0):
    """``fit_lr`` under an explicit normal-inverse-Wishart prior per (condition, t) — the form any
    ``DynamicsPrior.eval`` result takes (linear_dynamics.py:168-175, prior.py:18-45).

    ``mu0 (B,T,d), Phi (B,T,d,d), N_mean (B,T), N_covar (B,T)`` with ``d = 2dX+dU``; returns as :func:`fit_lr`.
    """
    nat = _lib.native()
    X, U = _dev(X), _dev(U)
    B, N, T1, dX = X.shape
    _, _, T, dU = U.shape
    if N <= 1:
        raise ValueError(f"Cannot fit dynamics to {N} sample(s)")
    d, p = 2 * dX + dU, dX + dU
    mu0, Phi, N_mean, N_covar = _dev(mu0), _dev(Phi), _dev(N_mean), _dev(N_covar)
    if mu0.shape != (B, T, d) or Phi.shape != (B, T, d, d) or N_mean.shape != (B, T) or N_covar.shape != (B, T):
        raise ValueError("NIW prior arrays must be (B,T,d), (B,T,d,d), (B,T), (B,T)")
    o = dict(dtype=X.dtype, device=X.device)
    F, f, covar = torch.empty(B, T, dX, p, **o), torch.empty(B, T, dX, **o), torch.empty(B, T, dX, dX, **o)
    info = torch.zeros(B, T, dtype=torch.int32, device=X.device)
    nat.call("fit_lr_niw_f64", B, N, T, dX, dU, X, U, mu0, Phi, N_mean, N_covar, float(prior_weight),
             float(regularization), F, f, covar, info)
    return F, f, covar, info


def dynamics_log_prob(X, U, F, f, covar):
    """Batched ``LinearDynamics.log_prob`` (linear_dynamics.py:54-64).

    ``X (B,N,T+1,dX), U (B,N,T,dU)``, dynamics ``F (B,T,dX,p), f (B,T,dX), covar (B,T,dX,dX)`` -> ``(B,N,T)``
    log-densities of every observed transition.  Raises ``numpy.linalg.LinAlgError`` when a ``covar_t`` is not positive
    definite (scipy's ``multivariate_normal.logpdf`` with ``allow_singular=False`` fails there too).
    """
    nat = _lib.native()
    X, U, F, f, covar = _dev(X), _dev(U), _dev(F), _dev(f), _dev(covar)
    B, N, T1, dX = X.shape
    T, dU = U.shape[2], U.shape[3]
    if T1 != T + 1 or F.shape != (B, T, dX, dX + dU):
        raise ValueError(f"inconsistent shapes {tuple(X.shape)} / {tuple(U.shape)} / {tuple(F.shape)}")
    inv, _, hld = spd_factor(covar)
    out = torch.empty(B, N, T, dtype=X.dtype, device=X.device)
    nat.call("dynamics_log_prob_f64", B, N, T, dX, dU, X, U, F, f, inv, hld, out)
    return out


def quadratic_cost_approximation_l2(t, w):
    """``QuadraticCosts.quadratic_cost_approximation_l2`` (quadratic_costs.py:121-139) for targets / weights
    ``t, w (..., p)`` -> ``C (...,p,p), c (...,p), cc (...)`` on the device."""
    nat = _lib.native()
    t, w = _dev(t), _dev(w)
    if t.shape != w.shape:
        raise ValueError(f"{tuple(t.shape)} != {tuple(w.shape)}")
    p = t.shape[-1]
    n = t.numel() // p
    o = dict(dtype=t.dtype, device=t.device)
    C, c, cc = torch.empty(t.shape + (p,), **o), torch.empty(t.shape, **o), torch.empty(t.shape[:-1], **o)
    nat.call("quadratic_cost_l2_f64", n, p, t, w, C, c, cc)
    return C, c, cc


def quadratic_cost_approximation_l1(xu, t, w, alpha: float):
    """``QuadraticCosts.quadratic_cost_approximation_l1`` (quadratic_costs.py:141-166): second-order expansion of
    ``sum_i w_i sqrt((t_i - xu_i)^2 + alpha)`` at the trajectory ``xu (..., p)``."""
    nat = _lib.native()
    xu, t, w = _dev(xu), _dev(t), _dev(w)
    if not (xu.shape == t.shape == w.shape):
        raise ValueError(f"{tuple(xu.shape)} / {tuple(t.shape)} / {tuple(w.shape)}")
    p = t.shape[-1]
    n = t.numel() // p
    o = dict(dtype=t.dtype, device=t.device)
    C, c, cc = torch.empty(t.shape + (p,), **o), torch.empty(t.shape, **o), torch.empty(t.shape[:-1], **o)
    nat.call("quadratic_cost_l1_f64", n, p, xu, t, w, float(alpha), C, c, cc)
    return C, c, cc


# =====================================================================================================
# GMM EM (GMMPrior.update)
# =====================================================================================================
class DeviceGMM:
    """Full-covariance Gaussian mixture fitted by EM on the device; rows may be sharded over ranks.

    Mirrors what ``GMMPrior`` needs from ``sklearn.mixture.GaussianMixture`` (prior_gmm.py:10-41):
    ``weights (K)``, ``means (K,d)``, ``covariances (K,d,d)``, ``precisions_cholesky (K,d,d)``,
    ``lower_bound``, ``n_iter``, ``converged`` and ``N`` (pool size seen by the last ``fit``).
    Defaults follow sklearn: ``tol=1e-3, reg_covar=1e-6, max_iter=100``.
    """

    def __init__(self, n_clusters: int, tol: float = 1e-3, reg_covar: float = 1e-6, max_iter: int = 100,
                 precision: str = "fp64"):
        self.K = int(n_clusters)
        self.kernels = "auto"  # "auto" | "scalar" | "dmma": see fit()
        # "fp64" (default: parity with scikit-learn at 1e-8) or "tf32x3": the E pass and the scatter statistics of the M
        # pass run on the tcgen05 tensor cores at a stated fp32 tolerance (include/donk_b200.h, donk_gmm_tc_estep_f32 /
        # donk_gmm_tc_mstep_f32); the partial sums, the all-reduce, the finalisation and all parameters stay fp64.
        # tc_m_pass = False keeps the fp64 DMMA M pass behind the tensor-core E pass.
        if precision not in ("fp64", "tf32x3"):
            raise ValueError(f"precision must be 'fp64' or 'tf32x3', got {precision!r}")
        self.precision = precision
        self.tc_m_pass = True
        self._tc = None
        self.tol, self.reg_covar, self.max_iter = tol, reg_covar, max_iter
        self.d = None
        self.weights = self.means = self.covariances = self.precisions_cholesky = None
        self.lower_bound = -np.inf
        self.n_iter = 0
        self.converged = False
        self.N = 0
        self.lower_bounds: list[float] = []

    # -- initialisation ----------------------------------------------------------------------------
    def set_params(self, weights, means, precisions=None, covariances=None, precisions_cholesky=None):
        """Explicit start (sklearn ``weights_init / means_init / precisions_init``)."""
        nat = _lib.native()
        # own copies: EM updates the parameters in place, the caller's arrays must stay untouched
        self.weights, self.means = _dev(weights).clone(), _dev(means).clone()
        self.K, self.d = self.means.shape
        if precisions_cholesky is not None:
            self.precisions_cholesky = _dev(precisions_cholesky).clone()
        elif precisions is not None:
            prec = _dev(precisions)
            self.precisions_cholesky = torch.empty_like(prec)
            info = torch.zeros(self.K, dtype=torch.int32, device=prec.device)
            nat.call("gmm_prec_chol_from_precisions_f64", self.d, self.K, prec, self.precisions_cholesky, info)
            if int(info.sum().item()):
                raise ValueError("'full precision' should be symmetric, positive-definite")
        self.covariances = _dev(covariances).clone() if covariances is not None else torch.zeros(
            self.K, self.d, self.d, dtype=torch.float64, device=self.means.device)
        self._cov_known = covariances is not None  # else a placeholder until the first M step fills it
        return self

    def init_from_labels(self, X_local, labels_local, group=None):
        """Hard-assignment start: one-hot responsibilities followed by an M-step
        (what sklearn's k-means initialisation feeds ``_initialize``, _base.py:119-128,159).

        Two statistics passes: the first (centred on 0) yields the cluster means, the second is centred
        on those means so the scatter matrices carry no cancellation error.
        """
        nat = _lib.native()
        X = _dev(X_local)
        n, d = X.shape
        self.d = d
        K = self.K
        o = dict(dtype=X.dtype, device=X.device)
        labels = torch.as_tensor(labels_local).to(X.device).long()
        onehot = torch.zeros(n, K, **o)
        onehot[torch.arange(n, device=X.device), labels] = 1
        n_total = torch.tensor([float(n)], **o)
        self._allreduce(n_total, group)
        self.weights = torch.full((K,), 1.0 / K, **o)
        self.means = torch.zeros(K, d, **o)
        self.covariances = torch.empty(K, d, d, **o)
        self.precisions_cholesky = torch.empty(K, d, d, **o)
        ws = torch.empty(nat.size("gmm_workspace_bytes", n, d, K) // 8 + 1, **o)
        stats = torch.empty(nat.size("gmm_stats_size", d, K), **o)
        lb = torch.zeros(1, **o)
        info = torch.zeros(K, dtype=torch.int32, device=X.device)
        shift = torch.empty(K, d, **o)
        for _ in range(2):
            nat.call("gmm_em_step_f64", n, d, K, X, None, self.means, None, onehot, stats, shift, ws,
                     ws.numel() * ws.element_size())
            self._allreduce(stats, group)
            nat.call("gmm_m_finalize_f64", float(n_total.item()), d, K, stats, shift, float(self.reg_covar),
                     self.weights, self.means, self.covariances, self.precisions_cholesky, lb, info)
        if int(info.sum().item()):
            raise ValueError("Fitting the mixture model failed because some components have ill-defined "
                             "empirical covariance")
        return self

    @staticmethod
    def _allreduce(t, group):
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    # -- EM ------------------------------------------------------------------------------------------
    def em_iteration(self, X, n_total: float, ws, stats, group=None) -> float:
        """One E+M iteration over this rank's rows; returns the mean log-likelihood (pre-M-step)."""
        nat = _lib.native()
        n, d = X.shape
        if getattr(self, "_shift", None) is None or self._shift.shape != self.means.shape:
            self._shift = torch.empty_like(self.means)
        mode = int(getattr(self, "_mode", 0))
        tc = self._tc
        tc_m = tc is not None and tc.get("mplanes") is not None
        if tc is not None and not tc_m:  # E pass on the tensor cores: responsibilities and lse totals straight into ws
            nat.call("gmm_tc_estep_f32", n, d, self.K, tc["planes"], tc["center"], self.weights, self.means,
                     self.precisions_cholesky, ws, ws[n * self.K:], tc["work"], tc["work"].numel() * 8)
            mode = 3
        if tc_m:
            # E and M pass on the tensor cores: the E pass leaves the transposed single-precision responsibilities in the M
            # pass's scratch (no fp64 round trip, no transpose kernel); scatter statistics about the same centre, drained
            # into fp64 partial sums
            nat.call("gmm_tc_estep_m_f32", n, d, self.K, tc["planes"], tc["center"], self.weights, self.means,
                     self.precisions_cholesky, ws, ws[n * self.K:], tc["work"], tc["work"].numel() * 8, tc["mwork"],
                     tc["mwork"].numel() * 8)
            nat.call("gmm_tc_mstep_f32", n, d, self.K, tc["mplanes"], tc["center"], None, ws[n * self.K:], stats, self._shift,
                     tc["mwork"], tc["mwork"].numel() * 8)
        else:
            nat.call("gmm_em_step_mode_f64", n, d, self.K, X, self.weights, self.means, self.precisions_cholesky, None,
                     stats, self._shift, ws, ws.numel() * ws.element_size(), mode)
        self._allreduce(stats, group)  # the one exchange step of the path (NCCL over NVLink)
        nat.call("gmm_m_finalize_f64", float(n_total), d, self.K, stats, self._shift, float(self.reg_covar),
                 self.weights, self.means, self.covariances, self.precisions_cholesky, self._lb, self._info)
        self._cov_known = True
        return self._lb

    def _centring_hardness(self) -> torch.Tensor:
        """``max_kj (mu_kj - x_bar_j)^2 / var_kj`` of the current parameters (device scalar, no host sync).

        ``var`` is the diagonal of the covariances; for an explicit precision start (``set_params(precisions=...)``,
        covariances unknown) it is bounded from below by ``1 / (P P^T)_jj`` (``(A^-1)_jj >= 1 / A_jj`` for SPD A), which
        over-estimates the hardness — the safe side.
        """
        if getattr(self, "_cov_known", True):
            var = torch.diagonal(self.covariances, dim1=-2, dim2=-1)
        else:
            var = 1.0 / (self.precisions_cholesky ** 2).sum(-1)
        xbar = (self.weights[:, None] * self.means).sum(0)
        return (((self.means - xbar) ** 2) / var).max()

    def _pick_mode(self, hardness: float) -> int:
        if self.kernels == "scalar":
            return 1
        if self.kernels == "dmma":
            return 2
        return 1 if (not np.isfinite(hardness) or hardness * np.finfo(np.float64).eps > 1e-9) else 0

    def fit(self, X_local, warm_start: bool = False, group=None, check_every: int = 1):
        """EM until ``|delta lower bound| < tol`` or ``max_iter`` (sklearn/mixture/_base.py:241-305).

        ``X_local (n_local,d)``: this rank's rows of the pool.  Parameters must have been initialised
        (``set_params`` / ``init_from_labels``) or come from a previous ``fit`` (``warm_start``).
        """
        import warnings

        nat = _lib.native()
        X = _dev(X_local)
        n, d = X.shape
        if self.means is None:
            raise DonkError("DeviceGMM.fit: parameters are not initialised")
        if d != self.d:
            raise ValueError(f"X has {d} features, the mixture has {self.d}")
        o = dict(dtype=X.dtype, device=X.device)
        n_total_t = torch.tensor([float(n)], **o)
        self._allreduce(n_total_t, group)
        n_total = float(n_total_t.item())
        ws = torch.empty(nat.size("gmm_workspace_bytes", n, d, self.K) // 8 + 1, **o)
        stats = torch.empty(nat.size("gmm_stats_size", d, self.K), **o)
        self._lb = torch.zeros(1, **o)
        self._info = torch.zeros(self.K, dtype=torch.int32, device=X.device)
        # Kernel family: the DMMA pass centres its statistics on the mixture mean; its rounding error grows as
        # eps (|mu_k - x_bar| / sigma_k)^2.  Where that would exceed ~1e-9 the scalar kernels (centred on every
        # cluster's own mean) are used whatever the pool size: results first, speed second.  The bound is evaluated
        # from the parameters going INTO every iteration (clusters move during EM), see _centring_hardness.
        self._mode = self._pick_mode(float(self._centring_hardness()))
        self._tc = self._tc_setup(X, m_pass=True) if self.precision == "tf32x3" else None
        lb = self.lower_bound if warm_start else -np.inf
        self.converged = False
        self.lower_bounds = []
        n_iter = 0
        for n_iter in range(1, self.max_iter + 1):
            prev = lb
            self.em_iteration(X, n_total, ws, stats, group)
            # one device->host read per iteration: the lower bound and the centring bound of the NEXT iteration
            lb, hard = torch.cat([self._lb, self._centring_hardness().reshape(1)]).tolist()
            self._mode = self._pick_mode(hard)
            if int(self._info.sum().item()):
                raise ValueError("Fitting the mixture model failed because some components have ill-defined "
                                 "empirical covariance (for instance caused by singleton or collapsed samples).")
            self.lower_bounds.append(lb)
            if abs(lb - prev) < self.tol:
                self.converged = True
                break
        self.lower_bound, self.n_iter, self.N = lb, n_iter, int(n_total)
        self._tc = None  # the split copy of the pool is released with the fit
        if not self.converged and self.max_iter > 0:
            from donk_b200.dynamics.prior.prior_gmm import ConvergenceWarning

            warnings.warn("Best performing initialization did not converge. Try different init parameters, or "
                          "increase max_iter, tol, or check for degenerate data.", ConvergenceWarning)
        return self

    def _tc_setup(self, X, m_pass: bool = False):
        """Split copy of the rows for the tensor-core E pass: two TF32 planes of ``x - c`` (``c`` = the mixture mean going
        into the fit, fixed over its iterations so that the copy is made once)."""
        nat = _lib.native()
        n, d = X.shape
        if not nat.size("gmm_tc_supported", d, self.K):
            raise DonkError(f"precision='tf32x3': d={d}, K={self.K} is outside the tensor-core E pass (d <= 80); use 'fp64'")
        o = dict(dtype=torch.float64, device=X.device)
        center = (self.weights[:, None] * self.means).sum(0).contiguous()
        planes = torch.empty(nat.size("gmm_tc_planes_bytes", n, d) // 8 + 32, **o)
        off = (-planes.data_ptr() // 8) % 32  # 256-byte aligned start
        planes = planes[off:]
        work = torch.empty(nat.size("gmm_tc_workspace_bytes", n, d, self.K) // 8 + 1, **o)
        nat.call("gmm_tc_prepare_f32", n, d, X, center, planes, planes.numel() * 8)
        tc = dict(center=center, planes=planes, work=work, mplanes=None, mwork=None)
        if m_pass and self.tc_m_pass and nat.size("gmm_tc_m_supported", d, self.K):
            mplanes = torch.empty(nat.size("gmm_tc_mplanes_bytes", n, d) // 8 + 32, **o)
            tc["mplanes"] = mplanes[(-mplanes.data_ptr() // 8) % 32:]
            tc["mwork"] = torch.empty(nat.size("gmm_tc_mwork_bytes", n, d, self.K) // 8 + 1, **o)
            nat.call("gmm_tc_mprepare_f32", n, d, X, center, tc["mplanes"], tc["mplanes"].numel() * 8)
        return tc

    def predict_proba_tf32x3(self, X) -> torch.Tensor:
        """``predict_proba`` through the tensor-core E pass (fp32 tolerance, see ``precision``)."""
        nat = _lib.native()
        X = _dev(X)
        n, d = X.shape
        tc = self._tc_setup(X)
        out = torch.empty(n * self.K + (n + 63) // 64, dtype=X.dtype, device=X.device)
        nat.call("gmm_tc_estep_f32", n, d, self.K, tc["planes"], tc["center"], self.weights, self.means,
                 self.precisions_cholesky, out, out[n * self.K:], tc["work"], tc["work"].numel() * 8)
        return out[:n * self.K].view(n, self.K)

    def eval_prior(self, XUX):
        """``GMMPrior.eval`` (prior_gmm.py:26-41) on the device: the NIW hyper-parameters the mixture assigns to the
        transitions ``XUX (N,d)`` — ``(mu0 (d), Phi (d,d), N_mean, N_covar)`` as device tensors / floats."""
        resp = self.predict_proba(XUX)
        share = resp.mean(dim=0)                                   # average responsibility of every cluster
        strength = float(torch.dot(share, self.weights)) * self.N  # the prior's weight in samples
        mu0 = share @ self.means
        Phi = torch.tensordot(share, self.covariances, dims=1) * strength
        return mu0, Phi, strength, strength - (1 + self.d)

    def predict_proba(self, X) -> torch.Tensor:
        nat = _lib.native()
        X = _dev(X)
        n, d = X.shape
        resp = torch.empty(n, self.K, dtype=X.dtype, device=X.device)
        nat.call("gmm_predict_proba_f64", n, d, self.K, X, self.weights, self.means, self.precisions_cholesky, resp)
        return resp


# =====================================================================================================
# Host-resident inputs: chunked, double-buffered H2D -> step -> D2H pipeline
# =====================================================================================================
class HostPipelinedILQR:
    """``ILQR.step`` for conditions whose inputs live in (pinned) host memory.

    The conditions are cut into chunks; chunk c+1 is copied to the device on one stream while chunk c is
    being solved on the other, so a step costs about max(PCIe time, kernel time) instead of their sum.
    ``host`` maps the ``BatchedILQR`` argument names (F, f, dyn_covar, prev_K, prev_k, prev_covar, C, c, cc,
    x0_mean, x0_covar) to CPU tensors with leading dimension B (pinned for asynchronous copies).
    """

    KEYS = ("F", "f", "dyn_covar", "prev_K", "prev_k", "prev_covar", "C", "c", "cc", "x0_mean", "x0_covar")

    def __init__(self, host: dict, E: int, chunks: int = 8):
        nat = _lib.native()
        self.host = {k: host[k] for k in self.KEYS}
        self.B = self.host["F"].shape[0]
        self.E = E
        self.chunks = max(1, min(chunks, self.B))
        self.bc = (self.B + self.chunks - 1) // self.chunks
        self.dev = nat.device
        self.cuda = self.dev.type == "cuda"
        n_slots = 2 if self.cuda else 1
        self.streams = [torch.cuda.Stream(self.dev) for _ in range(n_slots)] if self.cuda else [None]
        # per-slot device staging buffers and policy scratch (allocated once)
        self.slots = []
        for _ in range(n_slots):
            bufs = {k: torch.empty((self.bc,) + tuple(v.shape[1:]), dtype=v.dtype, device=self.dev)
                    for k, v in self.host.items()}
            bufs["eta"] = torch.empty(self.bc, E, dtype=torch.float64, device=self.dev)
            self.slots.append(dict(bufs=bufs, ws=None))
        pin = (lambda t: t.pin_memory()) if self.cuda else (lambda t: t)
        self.kl = pin(torch.empty(self.B, E, dtype=torch.float64))
        self.cost = pin(torch.empty(self.B, E, dtype=torch.float64))
        self.info = pin(torch.zeros(self.B, E, dtype=torch.int32))
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.host.values()) + self.B * E * 8
        self.d2h_bytes = self.kl.numel() * 8 * 2 + self.info.numel() * 4

    def step(self, eta_host: torch.Tensor):
        """Solve all conditions for ``eta_host (B,E)`` (CPU tensor); returns pinned ``kl, cost, info``."""
        import contextlib

        for c in range(self.chunks):
            lo, hi = c * self.bc, min((c + 1) * self.bc, self.B)
            if lo >= hi:
                break
            n = hi - lo
            slot = self.slots[c % len(self.slots)]
            stream = self.streams[c % len(self.streams)]
            ctx = torch.cuda.stream(stream) if self.cuda else contextlib.nullcontext()
            with ctx:
                d = {}
                for k, v in self.host.items():
                    d[k] = slot["bufs"][k][:n]
                    d[k].copy_(v[lo:hi], non_blocking=True)
                eta = slot["bufs"]["eta"][:n]
                eta.copy_(eta_host[lo:hi], non_blocking=True)
                il = BatchedILQR(d["F"], d["f"], d["dyn_covar"], d["prev_K"], d["prev_k"], d["C"], d["c"], d["cc"],
                                 d["x0_mean"], d["x0_covar"], prev_covar=d["prev_covar"])
                if slot["ws"] is not None and n == self.bc:
                    il._ws = slot["ws"]
                r = il.step(eta)
                if n == self.bc:
                    slot["ws"] = il._ws
                self.kl[lo:hi].copy_(r.kl_div, non_blocking=True)
                self.cost[lo:hi].copy_(r.expected_costs, non_blocking=True)
                self.info[lo:hi].copy_(r.info, non_blocking=True)
        if self.cuda:
            for s in self.streams:
                s.synchronize()
        return self.kl, self.cost, self.info


# =====================================================================================================
# Full GPS inner iteration for B conditions at once (TrajOptAlgorithm.iteration, traj_opt.py:35-96)
# =====================================================================================================
class BatchedTrajOpt:
    """``TrajOptAlgorithm`` (donk/traj_opt/traj_opt.py:12-96) over a batch of conditions, device resident.

    Per condition: fit the initial-state distribution, fit the dynamics (optionally under a shared GMM prior),
    run KL-constrained iLQR with ``kl_step * kl_step_mult`` and adapt ``kl_step_mult`` from predicted vs. actual
    improvement (``step_adjust``, lqg.py:309-350).  ``kl_step_mult`` is a ``(B,)`` tensor.
    """

    def __init__(self, kl_step: float, max_step_mult: float = 10, min_step_mult: float = 0.1,
                 dynamics_regularization: float = 1e-6):
        self.kl_step = kl_step
        self.max_step_mult, self.min_step_mult = max_step_mult, min_step_mult
        self.dynamics_regularization = dynamics_regularization
        self.kl_step_mult = None
        self.dyn = self.pol = self.x0 = self.costs = None
        self.last = None

    def _total_cost(self, dyn, pol, x0, costs):
        _, _, cost = lqr_forward(dyn[0], dyn[1], dyn[2], pol[0], pol[1], pol[2], x0[0], x0[1], 1e-6, *costs)
        return cost

    def iteration(self, X, U, C, c, cc, prev_K=None, prev_k=None, prev_covar=None, prior: "DeviceGMM | None" = None):
        """One iteration on roll-outs ``X (B,N,T+1,dX), U (B,N,T,dU)`` with quadratic costs ``C, c, cc``.

        The policy to stay close to is ``(prev_K, prev_k, prev_covar)`` or, if omitted, the policy of the previous
        iteration.  Returns the ``BatchedStepResult`` of the optimisation (E = 1).
        """
        X, U = _dev(X), _dev(U)
        B = X.shape[0]
        prev_dyn, prev_x0, prev_costs = self.dyn, self.x0, self.costs
        if prev_K is None:
            prev_K, prev_k, prev_covar = self.pol
        prev_pol = (_dev(prev_K), _dev(prev_k), _dev(prev_covar))
        if self.kl_step_mult is None:
            self.kl_step_mult = torch.ones(B, dtype=torch.float64, device=X.device)
        x0 = state_fit_from_rollouts(X)
        F, f, S, info = fit_lr(X, U, prior, self.dynamics_regularization)
        return self._optimize_and_adjust((F, f, S), info, x0, (_dev(C), _dev(c), _dev(cc)), prev_pol, prev_dyn, prev_x0,
                                         prev_costs)

    def _optimize_and_adjust(self, dyn, info, x0, costs, prev_pol, prev_dyn, prev_x0, prev_costs):
        F, f, S = dyn
        if int(info.sum().item()):
            raise np.linalg.LinAlgError("Singular matrix")
        self.x0 = x0
        self.dyn = (F, f, S)
        self.costs = costs
        ilqr = BatchedILQR(F, f, S, prev_pol[0], prev_pol[1], *self.costs, self.x0[0], self.x0[1],
                           prev_covar=prev_pol[2])
        res, status, n_steps = ilqr.optimize(self.kl_step * self.kl_step_mult)
        self.pol = (res.K[:, 0].clone(), res.k[:, 0].clone(), res.covar[:, 0].clone())
        self.last = dict(status=status, n_steps=n_steps, eta=res.eta[:, 0].clone(), kl=res.kl_div[:, 0].clone())
        if prev_dyn is not None:  # step_adjust (lqg.py:309-350)
            previous = self._total_cost(prev_dyn, prev_pol, prev_x0, prev_costs)
            predicted = self._total_cost(prev_dyn, self.pol, prev_x0, prev_costs)
            actual = self._total_cost(self.dyn, self.pol, self.x0, self.costs)
            mult = self.kl_step_mult * 0.5 * (predicted - previous) / (predicted - actual)
            self.kl_step_mult = torch.clamp(mult, self.min_step_mult, self.max_step_mult)
            self.last.update(previous_costs=previous, predicted_new_costs=predicted, actual_new_costs=actual)
        return res

    def iteration_from_host(self, X, U, C, c, cc, prior: "DeviceGMM | None" = None, chunks: int = 4, out_policy=None):
        """:meth:`iteration` for roll-outs and costs that live in (pinned) host memory, as a sampler leaves them.

        The conditions are cut into ``chunks``; a copy stream uploads chunk i + 1 (and, behind the roll-outs, the costs,
        which only the iLQR stage reads) while the state fit and ``fit_lr`` of chunk i run, so the upload hides behind the
        dynamics fit instead of preceding it (config 5: 8.6 GB per 128 conditions, 168 ms at 51 GB/s, against ~190 ms of
        ``fit_lr``).  The previous policy is the one of the previous iteration (a first iteration goes through
        :meth:`iteration`).  ``out_policy``: optional pinned host tensors ``(K, k, covar)`` the new policy is copied into
        (asynchronously on the current stream: synchronise before reading them).
        """
        if self.pol is None:
            raise DonkError("iteration_from_host continues a running optimisation: run iteration() once first")
        dev = self.pol[0].device
        B = X.shape[0]
        if dev.type != "cuda":  # host emulation (tests): nothing to overlap
            res = self.iteration(_dev(X), _dev(U), _dev(C), _dev(c), _dev(cc), prior=prior)
        else:
            o = dict(dtype=torch.float64, device=dev)
            key = (tuple(X.shape), tuple(U.shape), tuple(C.shape))
            if getattr(self, "_host_key", None) != key:  # device-side landing buffers, kept between iterations
                self._host_key = key
                self._host_buf = [torch.empty(t.shape, **o) for t in (X, U)]
                # two sets of cost buffers in turn: step_adjust reads the costs of the previous iteration after this one's
                # have landed
                self._host_costs = [[torch.empty(t.shape, **o) for t in (C, c, cc)] for _ in range(2)]
                self._host_turn = 0
                self._copy_stream = torch.cuda.Stream(device=dev)
            dX_, dU_ = self._host_buf
            self._host_turn ^= 1
            dC_, dc_, dcc_ = self._host_costs[self._host_turn]
            main = torch.cuda.current_stream(dev)
            bounds = [(i * B) // chunks for i in range(chunks + 1)]
            ready = []
            self._copy_stream.wait_stream(main)  # the buffers may still be read by the previous iteration
            with torch.cuda.stream(self._copy_stream):
                for lo, hi in zip(bounds[:-1], bounds[1:]):
                    dX_[lo:hi].copy_(X[lo:hi], non_blocking=True)
                    dU_[lo:hi].copy_(U[lo:hi], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self._copy_stream)
                    ready.append(ev)
                for d_, h_ in ((dC_, C), (dc_, c), (dcc_, cc)):
                    d_.copy_(h_, non_blocking=True)
                costs_ready = torch.cuda.Event()
                costs_ready.record(self._copy_stream)
            parts = []
            for (lo, hi), ev in zip(zip(bounds[:-1], bounds[1:]), ready):
                if hi == lo:
                    continue
                main.wait_event(ev)
                x0 = state_fit_from_rollouts(dX_[lo:hi])
                parts.append((x0,) + tuple(fit_lr(dX_[lo:hi], dU_[lo:hi], prior, self.dynamics_regularization)))
            x0 = tuple(torch.cat([q[0][i] for q in parts]) for i in range(2))
            F, f, S, info = (torch.cat([q[i] for q in parts]) for i in range(1, 5))
            main.wait_event(costs_ready)
            prev_dyn, prev_x0, prev_costs = self.dyn, self.x0, self.costs
            res = self._optimize_and_adjust((F, f, S), info, x0, (dC_, dc_, dcc_), self.pol, prev_dyn, prev_x0, prev_costs)
        if out_policy is not None:
            for h, t in zip(out_policy, self.pol):
                h.copy_(t, non_blocking=True)
        return res


# =====================================================================================================
# Device k-means: initial labelling for GMMPrior.update on pools that live on the GPU
# =====================================================================================================
class DeviceKMeans:
    """Lloyd's algorithm with greedy k-means++ seeding on the device (``donk_kmeans_pass_f64``).

    Stands in for the ``sklearn.cluster.KMeans(n_clusters, n_init=1, random_state)`` call that scikit-learn's
    ``GaussianMixture`` uses to initialise EM (sklearn/mixture/_base.py:119-128).  Same algorithm — greedy
    k-means++ with ``2 + log(K)`` candidates per centre, Lloyd iterations until the centres move by less than
    ``tol = 1e-4`` x mean feature variance, ``max_iter = 300`` — but a different random stream, so labels agree with
    scikit-learn statistically, not bit for bit.  One Lloyd iteration is one pass over the rows (labels and the
    per-cluster sums come out of the same kernel).  Rows may be sharded over ranks: the sums are all-reduced every
    iteration; seeding uses rank 0's rows and broadcasts the centres.
    """

    def __init__(self, n_clusters: int, random_state=None, max_iter: int = 300, tol: float = 1e-4):
        self.K, self.max_iter, self.tol = int(n_clusters), max_iter, tol
        self.seed = 0 if random_state is None else int(random_state)
        self.centers = None
        self.labels = None
        self.n_iter = 0

    @staticmethod
    def _pass(X, centers, labels=False, mind2=False, dist=False, sums=None, ws=None):
        nat = _lib.native()
        n, d = X.shape
        K = centers.shape[0]
        o = dict(dtype=X.dtype, device=X.device)
        lab = torch.empty(n, dtype=torch.int32, device=X.device) if labels else None
        m2 = torch.empty(n, **o) if mind2 else None
        D = torch.empty(n, K, **o) if dist else None
        nat.call("kmeans_pass_f64", n, d, K, X, centers.contiguous(), lab, m2, D, sums, ws,
                 ws.numel() * ws.element_size() if ws is not None else 0)
        return lab, m2, D

    def _seed(self, X):
        """Greedy k-means++ (sklearn/cluster/_kmeans.py:_kmeans_plusplus): each new centre is the best of
        ``2 + log K`` candidates drawn with probability proportional to the current squared distance (inverse-CDF
        sampling on the running sum, as scikit-learn does).  No host synchronisation: candidates, the winner and the
        updated distances stay on the device."""
        import math

        n, d = X.shape
        K = self.K
        g = torch.Generator(device=X.device)
        g.manual_seed(self.seed)
        trials = 2 + int(math.log(K))
        centers = torch.empty(K, d, dtype=X.dtype, device=X.device)
        nat = _lib.native()
        centers[0] = X[torch.randint(0, n, (1,), device=X.device, generator=g)][0]
        _, closest, _ = self._pass(X, centers[:1], mind2=True)
        ws = torch.empty(nat.size("kmeans_workspace_bytes", d, trials) // 8 + 1, dtype=X.dtype, device=X.device)
        dmin = torch.empty(n, trials, dtype=X.dtype, device=X.device)
        pot = torch.empty(trials, dtype=X.dtype, device=X.device)
        for i in range(1, K):
            cum = torch.cumsum(closest, 0)
            u = torch.rand(trials, dtype=X.dtype, device=X.device, generator=g) * cum[-1]
            cand = torch.searchsorted(cum, u).clamp_(max=n - 1)  # all distances zero (fewer distinct rows than centres): row 0
            # one read of the rows: distances to the candidates clamped by the current nearest distance + their column sums
            nat.call("kmeanspp_round_f64", n, d, trials, X, X[cand].contiguous(), closest, dmin, pot, ws,
                     ws.numel() * ws.element_size())
            best = pot.argmin().reshape(1)
            closest = dmin.index_select(1, best)[:, 0].contiguous()
            centers[i] = X[cand.index_select(0, best)][0]
        return centers

    def fit(self, X_local, group=None):
        import torch.distributed as dist

        nat = _lib.native()
        X = _dev(X_local)
        n, d = X.shape
        K = self.K
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        rank = dist.get_rank(group) if multi else 0
        o = dict(dtype=X.dtype, device=X.device)
        centers = self._seed(X) if rank == 0 else torch.empty(K, d, **o)
        if multi:
            dist.broadcast(centers, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        stat = torch.stack([X.sum(0), (X * X).sum(0), torch.full((d,), float(n), **o)])
        DeviceGMM._allreduce(stat, group)
        var = stat[1] / stat[2] - (stat[0] / stat[2]) ** 2
        tol = self.tol * float(var.mean())
        ws = torch.empty(nat.size("kmeans_workspace_bytes", d, K) // 8 + 1, **o)
        sums = torch.empty(K * d + K, **o)
        for it in range(1, self.max_iter + 1):
            self._pass(X, centers, sums=sums, ws=ws)
            DeviceGMM._allreduce(sums, group)
            cnt = sums[K * d:]
            new = torch.where(cnt[:, None] > 0, sums[:K * d].reshape(K, d) / cnt.clamp(min=1)[:, None], centers)
            shift = float(((new - centers) ** 2).sum())
            centers = new
            self.n_iter = it
            if shift <= tol:
                break
        self.centers = centers
        self.labels, _, _ = self._pass(X, centers, labels=True)
        return self


# =====================================================================================================
# Device-resident transition pool
# =====================================================================================================
class DeviceTransitionPool:
    """``donk.samples.TransitionPool`` (donk/samples/pool.py:6-56) with the rows kept in device memory.

    Same semantics — rows ``[x_t | u_t | x_{t+1}]`` in roll-out order, the newest ``capacity`` rows are kept,
    ``get_transitions(N)`` returns the ``N`` newest — but ``add`` takes roll-outs that may already live on the GPU and
    writes their transitions straight into a ring buffer (``donk_to_transitions_f64``), so a GPS iteration uploads only
    its new roll-outs instead of the whole pool (5.76 GB at 10 M x 72).  ``get_transitions`` returns a view of the
    buffer when the requested window is contiguous and one device-to-device gather otherwise (2 ms per 5.76 GB: noise
    next to an EM iteration).  With roll-outs sharded over ranks every rank keeps its own pool;
    ``GMMPrior.update(pool.get_transitions())`` then all-reduces the EM statistics (``DeviceGMM.fit``).
    """

    def __init__(self, capacity: int = None) -> None:
        self.capacity = capacity
        self._buf = None   # (rows_allocated, d)
        self._start = 0    # ring start (oldest row) when capacity is set
        self._size = 0

    def __len__(self) -> int:
        return self._size

    def add(self, X, U) -> None:
        assert X.shape[0] == U.shape[0], f"{X.shape[0]} != {U.shape[0]}"
        assert X.shape[1] == U.shape[1] + 1, f"{X.shape[1]} != {U.shape[1] + 1}"
        X, U = _dev(X), _dev(U)
        if self._buf is None:
            self.dX, self.dU = X.shape[-1], U.shape[-1]
        else:
            assert X.shape[-1] == self.dX, f"{X.shape[-1]} != {self.dX}"
            assert U.shape[-1] == self.dU, f"{U.shape[-1]} != {self.dU}"
        d = 2 * self.dX + self.dU
        N, T = U.shape[0], U.shape[1]
        rows = N * T
        nat = _lib.native()
        if self.capacity is None:  # growing buffer, geometric reallocation
            need = self._size + rows
            if self._buf is None or need > self._buf.shape[0]:
                grown = torch.empty(max(need, 2 * (self._buf.shape[0] if self._buf is not None else 0)), d,
                                    dtype=X.dtype, device=X.device)
                if self._size:
                    grown[:self._size] = self._buf[:self._size]
                self._buf = grown
            nat.call("to_transitions_f64", N, T, self.dX, self.dU, X, U, self._buf[self._size:need])
            self._size = need
            return
        cap = self.capacity
        if self._buf is None:
            self._buf = torch.empty(cap, d, dtype=X.dtype, device=X.device)
        if rows >= cap:  # only the newest `cap` rows of this batch survive
            new = torch.empty(rows, d, dtype=X.dtype, device=X.device)
            nat.call("to_transitions_f64", N, T, self.dX, self.dU, X, U, new)
            self._buf.copy_(new[rows - cap:])
            self._start, self._size = 0, cap
            return
        end = (self._start + self._size) % cap
        if end + rows <= cap:  # fits without wrapping: the kernel writes in place
            nat.call("to_transitions_f64", N, T, self.dX, self.dU, X, U, self._buf[end:end + rows])
        else:
            new = torch.empty(rows, d, dtype=X.dtype, device=X.device)
            nat.call("to_transitions_f64", N, T, self.dX, self.dU, X, U, new)
            first = cap - end
            self._buf[end:] = new[:first]
            self._buf[:rows - first] = new[first:]
        overflow = max(0, self._size + rows - cap)
        self._start = (self._start + overflow) % cap
        self._size = min(cap, self._size + rows)

    def get_transitions(self, N: int = None) -> torch.Tensor:
        """All, or the ``N`` newest, transitions ``(n, d)`` on the device, oldest first (pool.py:44-56)."""
        if self._buf is None:
            raise ValueError("pool is empty")
        n = self._size if N is None else min(int(N), self._size)
        if self.capacity is None:
            return self._buf[self._size - n:self._size]
        cap = self.capacity
        lo = (self._start + self._size - n) % cap
        if lo + n <= cap:
            return self._buf[lo:lo + n]
        return torch.cat([self._buf[lo:], self._buf[:lo + n - cap]], dim=0)
