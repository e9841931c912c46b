"""Oracle: KL-constrained iLQR (array-in / array-out numpy restatement).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates ``donk/traj_opt/lqg.py`` of the reference:
  backward               lqg.py:126-182
  forward                lqg.py:185-244
  extended_costs_kl      lqg.py:247-273
  kl_divergence_action   lqg.py:276-306
  ILQR.step              lqg.py:55-75
  ILQR.optimize          lqg.py:90-123   (+ scipy.optimize.brentq, restated below)
  step_adjust            lqg.py:309-350
and ``QuadraticCosts.expected_costs`` (donk/costs/quadratic_costs.py:55-92,106-119).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy.linalg import lapack


def _sym(A: np.ndarray) -> np.ndarray:
    """(A + A^T)/2 on the last two axes, new array (donk/utils/batched.py:52-59)."""
    return (A + np.swapaxes(A, -1, -2)) / 2


def _spd_inverse(A: np.ndarray) -> np.ndarray:
    """A^-1 through LAPACK posv with identity right-hand side.

    Same factorisation/solve the reference reaches via
    ``scipy.linalg.solve(A, eye, assume_a="pos")`` (lqg.py:165).
    """
    _, x, info = lapack.dposv(A, np.eye(A.shape[0]), lower=0)
    if info != 0:
        raise np.linalg.LinAlgError(f"matrix not positive definite (info={info})")
    return x


def backward(F, f, C, c, gamma: float = 1.0):
    """LQR backward (Riccati) pass, lqg.py:126-182.

    F (T,dX,p), f (T,dX), C (T+1,p,p), c (T+1,p)  ->  K (T,dU,dX), k (T,dU),
    covar (T,dU,dU) = Quu^-1, inv_covar (T,dU,dU) = Quu.
    """
    T, dX, p = F.shape
    dU = p - dX
    K = np.empty((T, dU, dX))
    k = np.empty((T, dU))
    covar = np.empty((T, dU, dU))
    inv_covar = np.empty((T, dU, dU))

    V = C[T, :dX, :dX]  # lqg.py:151
    v = c[T, :dX]
    for t in range(T - 1, -1, -1):
        Ft = F[t]
        Q = _sym(C[t] + gamma * (Ft.T @ V @ Ft))  # lqg.py:160,162
        q = c[t] + gamma * (Ft.T @ (V @ f[t] + v))  # lqg.py:161
        Quu = Q[dX:, dX:]
        Quu_inv = _sym(_spd_inverse(Quu))  # lqg.py:165-166
        K[t] = -Quu_inv @ Q[dX:, :dX]  # lqg.py:170
        k[t] = -Quu_inv @ q[dX:]  # lqg.py:172
        covar[t] = Quu_inv
        inv_covar[t] = Quu
        V = _sym(Q[:dX, :dX] + Q[:dX, dX:] @ K[t])  # lqg.py:178,180
        v = q[:dX] + Q[:dX, dX:] @ k[t]  # lqg.py:179
    return K, k, covar, inv_covar


def forward(F, f, dyn_covar, K, k, pol_covar, x0_mean, x0_covar, regularization: float = 1e-6):
    """Gaussian state-action marginals, lqg.py:185-244.

    Returns mean (T+1,p), covar (T+1,p,p).  The recursion consumes the raw
    covariance; symmetrise + regularise touch the stored output only
    (lqg.py:238-239), and the last row only has its state block symmetrised
    (lqg.py:242).
    """
    T, dX, p = F.shape
    mean = np.zeros((T + 1, p))
    covar = np.zeros((T + 1, p, p))
    mean[0, :dX] = x0_mean
    covar[0, :dX, :dX] = x0_covar
    idx = np.arange(p)
    for t in range(T):
        Sxx = covar[t, :dX, :dX]
        mean[t, dX:] = K[t] @ mean[t, :dX] + k[t]  # lqg.py:227
        Sux = K[t] @ Sxx  # lqg.py:229
        covar[t, dX:, :dX] = Sux
        covar[t, :dX, dX:] = Sux.T
        covar[t, dX:, dX:] = Sux @ K[t].T + pol_covar[t]  # lqg.py:231
        mean[t + 1, :dX] = F[t] @ mean[t] + f[t]  # lqg.py:234
        covar[t + 1, :dX, :dX] = F[t] @ covar[t] @ F[t].T + dyn_covar[t]  # lqg.py:236
        covar[t] = _sym(covar[t])
        covar[t, idx, idx] += regularization
    covar[T, :dX, :dX] = _sym(covar[T, :dX, :dX])
    return mean, covar


def extended_costs_kl(K, k, inv_covar):
    """Quadratic expansion of -log p_prev(u|x), lqg.py:247-273."""
    T, dU, dX = K.shape
    p = dX + dU
    C = np.empty((T, p, p))
    c = np.empty((T, p))
    for t in range(T):
        L = inv_covar[t]
        KtL = K[t].T @ L
        C[t, :dX, :dX] = KtL @ K[t]
        C[t, :dX, dX:] = -KtL
        C[t, dX:, :dX] = -L @ K[t]
        C[t, dX:, dX:] = L
        c[t, :dX] = KtL @ k[t]
        c[t, dX:] = -L @ k[t]
    return C, c


def kl_divergence_action(X, K, k, covar, prev_K, prev_k, prev_covar, prev_inv_covar) -> float:
    """Mean over t of KL(new(u|x_t) || prev(u|x_t)), lqg.py:276-306.

    ``chol_covar`` of either policy is the lower Cholesky factor of its
    covariance (donk/models/tvlg.py:76-86, donk/utils/batched.py:34-49).
    """
    T, dU, _ = K.shape
    total = 0.0
    for t in range(T):
        du = (prev_K[t] - K[t]) @ X[t] + prev_k[t] - k[t]
        Lp = np.linalg.cholesky(prev_covar[t])
        Ln = np.linalg.cholesky(covar[t])
        total += 0.5 * (
            np.einsum("ij,ji->", prev_inv_covar[t], covar[t])
            + du @ prev_inv_covar[t] @ du
            - dU
            + 2 * sum(np.log(np.diag(Lp)) - np.log(np.diag(Ln)))
        )
    return total / T


def expected_costs(C, c, cc, mean, covar, dX: int):
    """Per-timestep expected quadratic cost, quadratic_costs.py:55-92,106-119.

    mean (T+1,p) with the action part of the last row ignored (treated as 0);
    covar (T+1,p,p) as returned by ``forward`` (regularised).
    """
    traj = mean.copy()
    traj[-1, dX:] = 0
    quad = np.einsum("ti,tij,tj->t", traj, C, traj) / 2
    return quad + np.sum(traj * c, axis=-1) + cc + np.einsum("tij,tji->t", C, covar) / 2


@dataclass
class ILQRProblem:
    """Everything ``ILQR(dynamics, prev_pol, costs, initial_state)`` holds (lqg.py:32-53)."""

    F: np.ndarray
    f: np.ndarray
    dyn_covar: np.ndarray
    C: np.ndarray
    c: np.ndarray
    cc: np.ndarray
    prev_K: np.ndarray
    prev_k: np.ndarray
    prev_covar: np.ndarray
    prev_inv_covar: np.ndarray
    x0_mean: np.ndarray
    x0_covar: np.ndarray

    def __post_init__(self):
        self.C_kl, self.c_kl = extended_costs_kl(self.prev_K, self.prev_k, self.prev_inv_covar)


@dataclass
class StepResult:
    eta: float
    K: np.ndarray
    k: np.ndarray
    covar: np.ndarray
    inv_covar: np.ndarray
    kl_div: float
    expected_costs: float
    mean: np.ndarray
    traj_covar: np.ndarray


def ilqr_step(prob: ILQRProblem, eta: float) -> StepResult:
    """ILQR.step, lqg.py:55-75."""
    dX = prob.f.shape[1]
    C_ext = prob.C / eta
    C_ext[:-1] += prob.C_kl
    c_ext = prob.c / eta
    c_ext[:-1] += prob.c_kl
    K, k, covar, inv_covar = backward(prob.F, prob.f, C_ext, c_ext)
    mean, tcov = forward(prob.F, prob.f, prob.dyn_covar, K, k, covar, prob.x0_mean, prob.x0_covar)
    kl = kl_divergence_action(
        mean[:, :dX], K, k, covar, prob.prev_K, prob.prev_k, prob.prev_covar, prob.prev_inv_covar
    )
    cost = float(np.sum(expected_costs(prob.C, prob.c, prob.cc, mean, tcov, dX)))
    return StepResult(eta, K, k, covar, inv_covar, float(kl), cost, mean, tcov)


def brentq(fun, xa: float, xb: float, xtol: float = 2e-12, rtol: float = 1e-2, maxiter: int = 100):
    """Brent's root bracketing, restated from the classic published algorithm as
    implemented by scipy 1.18.1 ``optimize.brentq`` (scipy/optimize/Zeros/brentq.c).

    Not part of the reference tree (scipy~=1.10 pinned in setup.py:16; call
    site lqg.py:121).  Returns (root, list of evaluated abscissae).  The iterate
    sequence is pinned against live scipy in tests/golden.
    """
    trace = []

    def g(x):
        trace.append(x)
        return fun(x)

    xpre, xcur = xa, xb
    xblk = fblk = spre = scur = 0.0
    fpre = g(xpre)
    fcur = g(xcur)
    if fpre == 0:
        return xpre, trace
    if fcur == 0:
        return xcur, trace
    if np.sign(fpre) == np.sign(fcur):
        raise ValueError("f(a) and f(b) must have different signs")
    for _ in range(maxiter):
        if fpre != 0 and fcur != 0 and np.sign(fpre) != np.sign(fcur):
            xblk, fblk = xpre, fpre
            spre = scur = xcur - xpre
        if abs(fblk) < abs(fcur):
            xpre, xcur, xblk = xcur, xblk, xcur
            fpre, fcur, fblk = fcur, fblk, fcur
        delta = (xtol + rtol * abs(xcur)) / 2
        sbis = (xblk - xcur) / 2
        if fcur == 0 or abs(sbis) < delta:
            return xcur, trace
        if abs(spre) > delta and abs(fcur) < abs(fpre):
            if xpre == xblk:
                stry = -fcur * (xcur - xpre) / (fcur - fpre)  # secant
            else:
                dpre = (fpre - fcur) / (xpre - xcur)
                dblk = (fblk - fcur) / (xblk - xcur)
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre))
            if 2 * abs(stry) < min(abs(spre), 3 * abs(sbis) - delta):
                spre, scur = scur, stry
            else:
                spre = scur = sbis
        else:
            spre = scur = sbis
        xpre, fpre = xcur, fcur
        if abs(scur) > delta:
            xcur += scur
        else:
            xcur += delta if sbis > 0 else -delta
        fcur = g(xcur)
    raise RuntimeError("brentq failed to converge")


def ilqr_optimize(prob: ILQRProblem, kl_step: float, min_eta=1e-6, max_eta=1e16, rtol=1e-2):
    """ILQR.optimize, lqg.py:90-123.  Returns (StepResult, list of eta visited)."""
    visited = []

    def step(eta):
        visited.append(eta)
        return ilqr_step(prob, eta)

    if step(min_eta).kl_div <= kl_step:  # lqg.py:108-109
        return step(min_eta), visited
    if step(max_eta).kl_div > kl_step:  # lqg.py:112-113
        raise ValueError(f"max_eta eta to low ({max_eta})")
    root, _ = brentq(lambda le: step(np.exp(le)).kl_div - kl_step, np.log(min_eta), np.log(max_eta), rtol=rtol)
    return step(np.exp(root)), visited


def step_adjust(
    step_mult, dyn, pol, x0, costs, dyn_prev, pol_prev, x0_prev, costs_prev, max_step_mult=10.0, min_step_mult=0.1
):
    """KL step multiplier update, lqg.py:309-350.

    dyn = (F,f,dyn_covar); pol = (K,k,covar); x0 = (mean,covar); costs = (C,c,cc).
    Returns (new_mult, previous_costs, predicted_new_costs, actual_new_costs).
    """

    def total(d, pl, s, cs):
        m, S = forward(d[0], d[1], d[2], pl[0], pl[1], pl[2], s[0], s[1])
        return float(np.sum(expected_costs(cs[0], cs[1], cs[2], m, S, d[1].shape[1])))

    prev = total(dyn_prev, pol_prev, x0_prev, costs_prev)
    pred = total(dyn_prev, pol, x0_prev, costs_prev)
    act = total(dyn, pol, x0, costs)
    mult = step_mult * 0.5 * (pred - prev) / (pred - act)
    return float(np.clip(mult, min_step_mult, max_step_mult)), prev, pred, act
