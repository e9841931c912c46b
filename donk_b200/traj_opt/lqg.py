"""KL-constrained iLQR with the reference's API (donk/traj_opt/lqg.py), computed on the B200.

Drop-in names: ``ILQR``, ``ILQRStepResult``, ``backward``, ``forward``, ``extended_costs_kl``,
``kl_divergence_action``, ``step_adjust``.  Arguments and results are numpy arrays inside the same value
types as the reference; every call copies its inputs to the device, runs the CUDA kernels of
``donk_b200/csrc/ilqr.cuh`` for a batch of one, and copies the results back.  Use
``donk_b200.batched.BatchedILQR`` to keep many conditions resident on the device.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from donk_b200 import batched
from donk_b200.costs.quadratic_costs import QuadraticCosts
from donk_b200.policy import LinearGaussianPolicy
from donk_b200.samples import StateDistribution
from donk_b200.samples.traj_dist import TrajectoryDistribution


def _np(t):
    return t.detach().cpu().numpy()


@dataclass
class ILQRStepResult:
    """Result of one iLQR step (lqg.py:18-26)."""

    eta: float
    policy: LinearGaussianPolicy
    kl_div: float
    expected_costs: float
    trajectory: TrajectoryDistribution


class ILQR:
    """iLQG trajectory optimisation (lqg.py:29-123)."""

    def __init__(self, dynamics, prev_pol: LinearGaussianPolicy, costs: QuadraticCosts,
                 initial_state: StateDistribution) -> None:
        self.dynamics = dynamics
        self.prev_pol = prev_pol
        self.costs = costs
        self.initial_state = initial_state
        a = lambda x: np.asarray(x, dtype=np.float64)[None]  # noqa: E731
        # Only what the caller supplied is shipped; the missing factor is derived on the device.
        covar, inv = prev_pol.given_forms()
        self._b = batched.BatchedILQR(
            a(dynamics.F), a(dynamics.f), a(dynamics.covar), a(prev_pol.K), a(prev_pol.k), a(costs.C), a(costs.c),
            a(costs.cc), a(initial_state.mean), a(initial_state.covar),
            prev_covar=None if covar is None else a(covar), prev_inv_covar=None if inv is None else a(inv))
        self.C_kl, self.c_kl = _np(self._b.C_kl[0]), _np(self._b.c_kl[0])

    def _result(self, r, b=0, e=0) -> ILQRStepResult:
        if int(r.info[b, e]) != 0:
            raise np.linalg.LinAlgError("Matrix is not positive definite")  # scipy solve(assume_a="pos"), lqg.py:165
        pol = LinearGaussianPolicy(_np(r.K[b, e]).copy(), _np(r.k[b, e]).copy(), _np(r.covar[b, e]).copy(),
                                   _np(r.inv_covar[b, e]).copy())
        traj = TrajectoryDistribution(_np(r.traj_mean[b, e]).copy(), _np(r.traj_covar[b, e]).copy(),
                                      self.dynamics.dX, self.dynamics.dU)
        return ILQRStepResult(float(r.eta[b, e]), pol, float(r.kl_div[b, e]), float(r.expected_costs[b, e]), traj)

    def step(self, eta: float) -> ILQRStepResult:
        """Unconstrained optimisation under the Lagrange multiplier ``eta`` (lqg.py:55-75)."""
        return self._result(self._b.step(float(eta), want_traj=True))

    def sample_surface(self, min_eta: float = 1e-6, max_eta: float = 1e16, N: int = 100) -> list[ILQRStepResult]:
        """``N`` log-spaced multipliers evaluated in one batched launch (lqg.py:77-88)."""
        import torch

        etas = np.logspace(np.log10(min_eta), np.log10(max_eta), num=N)
        r = self._b.step(torch.as_tensor(etas)[None], want_traj=True)
        return [self._result(r, 0, e) for e in range(N)]

    def optimize(self, kl_step: float, min_eta: float = 1e-6, max_eta: float = 1e16,
                 rtol: float = 1e-2) -> ILQRStepResult:
        """Find ``eta`` with ``kl_div == kl_step`` by Brent's method in log space (lqg.py:90-123).

        Raises ``ValueError("max_eta eta to low ...")`` like the reference when the constraint cannot be met.
        """
        r, _, _ = self._b.optimize(kl_step, min_eta, max_eta, rtol, want_traj=True)
        return self._result(r)


def backward(dynamics, C: np.ndarray, c: np.ndarray, gamma: float = 1) -> LinearGaussianPolicy:
    """LQR backward pass (lqg.py:126-182); ``numpy.linalg.LinAlgError`` if a ``Quu`` is not PD."""
    K, k, cov, inv, info = batched.lqr_backward(np.asarray(dynamics.F)[None], np.asarray(dynamics.f)[None],
                                                np.asarray(C)[None], np.asarray(c)[None], gamma)
    if int(info[0]) != 0:
        raise np.linalg.LinAlgError("Matrix is not positive definite")
    return LinearGaussianPolicy(_np(K[0]), _np(k[0]), _np(cov[0]), _np(inv[0]))


def forward(dynamics, policy: LinearGaussianPolicy, initial_state: StateDistribution,
            regularization: float = 1e-6) -> TrajectoryDistribution:
    """LQR forward pass: state-action marginals (lqg.py:185-244)."""
    a = lambda x: np.asarray(x, dtype=np.float64)[None]  # noqa: E731
    mean, covar = batched.lqr_forward(a(dynamics.F), a(dynamics.f), a(dynamics.covar), a(policy.K), a(policy.k),
                                      a(policy.covar), a(initial_state.mean), a(initial_state.covar), regularization)
    return TrajectoryDistribution(_np(mean[0]), _np(covar[0]), dynamics.dX, dynamics.dU)


def extended_costs_kl(prev_pol: LinearGaussianPolicy):
    """Quadratic expansion of ``-log p_prev(u|x)`` (lqg.py:247-273) -> ``C (T,p,p), c (T,p)``."""
    C, c = batched.extended_costs_kl(prev_pol.K, prev_pol.k, prev_pol.inv_covar)
    return _np(C), _np(c)


def kl_divergence_action(X: np.ndarray, pol: LinearGaussianPolicy, prev_pol: LinearGaussianPolicy) -> float:
    """Mean KL divergence between new and previous action distributions along ``X (T,dX)`` (lqg.py:276-306)."""
    a = lambda x: np.asarray(x, dtype=np.float64)[None]  # noqa: E731
    kl = batched.kl_divergence_action(a(X), a(pol.K), a(pol.k), a(pol.covar), a(prev_pol.K), a(prev_pol.k),
                                      a(prev_pol.covar), a(prev_pol.inv_covar))
    return float(kl[0])


def _expected_total(dyn, pol, x0, costs) -> float:
    a = lambda x: np.asarray(x, dtype=np.float64)[None]  # noqa: E731
    _, _, cost = batched.lqr_forward(a(dyn.F), a(dyn.f), a(dyn.covar), a(pol.K), a(pol.k), a(pol.covar), a(x0.mean),
                                     a(x0.covar), 1e-6, a(costs.C), a(costs.c), a(costs.cc))
    return float(cost[0])


def step_adjust(step_mult: float, dyn, policy, x0, costs, dyn_prev, policy_prev, x0_prev, costs_prev,
                max_step_mult: float = 10, min_step_mult: float = 0.1) -> float:
    """KL step multiplier from predicted vs. actual improvement (lqg.py:309-350)."""
    from donk_b200 import datalogging

    previous_costs = _expected_total(dyn_prev, policy_prev, x0_prev, costs_prev)
    predicted_new_costs = _expected_total(dyn_prev, policy, x0_prev, costs_prev)
    actual_new_costs = _expected_total(dyn, policy, x0, costs)
    step_mult = step_mult * 0.5 * (predicted_new_costs - previous_costs) / (predicted_new_costs - actual_new_costs)
    step_mult = float(np.clip(step_mult, min_step_mult, max_step_mult))
    datalogging.log(previous_costs=previous_costs, predicted_new_costs=predicted_new_costs,
                    actual_new_costs=actual_new_costs, step_mult=step_mult)
    return step_mult
