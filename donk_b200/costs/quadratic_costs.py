"""Quadratic cost container (donk/costs/quadratic_costs.py:12-139).

What the hot path consumes: ``C, c, cc``, ``compute_costs``, ``expected_costs``, ``+``/``*`` and the L2 / L1
constructors, which run as device kernels.  Symbolic cost authoring stays host-side in
the reference and is out of scope.
"""
from __future__ import annotations

from typing import Any

import numpy as np


class QuadraticCosts:
    """``cost(x,u) = 1/2 [x u]^T C [x u] + [x u]^T c + cc``."""

    def __init__(self, C, c, cc) -> None:
        T = c.shape[:-1]
        dXU = c.shape[-1]
        assert C.shape == (*T, dXU, dXU), f"{C.shape} != {(*T, dXU, dXU)}"
        assert c.shape == (*T, dXU), f"{c.shape} != {(*T, dXU)}"
        assert cc.shape == (*T,), f"{cc.shape} != {(*T, )}"
        self.C = C
        self.c = c
        self.cc = cc

    def __add__(self, other: Any):
        if isinstance(other, QuadraticCosts):
            return QuadraticCosts(self.C + other.C, self.c + other.c, self.cc + other.cc)
        return NotImplemented

    def __mul__(self, other: Any):
        if np.isscalar(other):
            return QuadraticCosts(self.C * other, self.c * other, self.cc * other)
        return NotImplemented

    __rmul__ = __mul__

    def compute_costs(self, X, U):
        """Costs of trajectories ``X (...,T+1,dX), U (...,T,dU)`` -> ``(...,T+1)`` (quadratic_costs.py:55-92)."""
        X, U = np.asarray(X), np.asarray(U)
        pad = np.zeros(U.shape[:-2] + (1, U.shape[-1]))
        traj = np.concatenate([X, np.concatenate([U, pad], axis=-2)], axis=-1)
        return np.einsum("...i,...ij,...j->...", traj, self.C, traj) / 2 + np.sum(traj * self.c, axis=-1) + self.cc

    def quadratic_approximation(self, X, U):
        return self

    def expected_costs(self, traj):
        """Per-step expected cost under a trajectory distribution (quadratic_costs.py:106-119)."""
        return self.compute_costs(traj.X_mean, traj.U_mean) + np.einsum("...ij,...ji->...", self.C, traj.covar) / 2

    @staticmethod
    def quadratic_cost_approximation_l2(t, w):
        """Exact quadratic form of ``0.5 sum_i w_i (t_i - xu_i)^2`` (quadratic_costs.py:121-139), built on the device
        (``donk_quadratic_cost_l2_f64``); ``t, w (..., p)``."""
        from donk_b200 import batched

        C, c, cc = batched.quadratic_cost_approximation_l2(np.asarray(t, dtype=np.float64), np.asarray(w, dtype=np.float64))
        return QuadraticCosts(C.cpu().numpy(), c.cpu().numpy(), cc.cpu().numpy())

    @staticmethod
    def quadratic_cost_approximation_l1(xu, t, w, alpha: float):
        """Second-order expansion of ``sum_i w_i sqrt((t_i - xu_i)^2 + alpha)`` at ``xu`` (quadratic_costs.py:141-166),
        built on the device (``donk_quadratic_cost_l1_f64``)."""
        from donk_b200 import batched

        f64 = lambda a: np.asarray(a, dtype=np.float64)  # noqa: E731
        xu, t, w = np.broadcast_arrays(f64(xu), f64(t), f64(w))
        C, c, cc = batched.quadratic_cost_approximation_l1(xu, t, w, alpha)
        return QuadraticCosts(C.cpu().numpy(), c.cpu().numpy(), cc.cpu().numpy())
