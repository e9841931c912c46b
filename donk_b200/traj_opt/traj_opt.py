"""Outer trajectory-optimisation iteration (donk/traj_opt/traj_opt.py:12-96)."""
from __future__ import annotations

import numpy as np

from donk_b200 import datalogging
from donk_b200.dynamics import linear_dynamics
from donk_b200.samples import StateDistribution
from donk_b200.traj_opt import lqg
from donk_b200.traj_opt.lqg import ILQR


class TrajOptAlgorithm:
    """fit dynamics -> quadratic cost approximation -> KL-constrained iLQR -> step adjustment."""

    def __init__(self, kl_step: float, max_step_mult: float = 10, min_step_mult: float = 0.1,
                 dynamics_regularization: float = 1e-6) -> None:
        self.kl_step = kl_step
        self.kl_step_mult: float = 1
        self.max_step_mult = max_step_mult
        self.min_step_mult = min_step_mult
        self.dynamics_regularization = dynamics_regularization
        self.pol = self.dyn = self.x0 = self.costs = self.ilqr = None

    def iteration(self, X, U, cost_function, prev_pol=None, prior=None) -> None:
        """One iteration on roll-outs ``X (N,T+1,dX), U (N,T,dU)`` (traj_opt.py:35-96)."""
        prev_dyn, prev_x0, prev_costs = self.dyn, self.x0, self.costs
        prev_pol = prev_pol if prev_pol is not None else self.pol
        self.x0 = StateDistribution.fit(X[:, 0])
        self.dyn = linear_dynamics.fit_lr(X, U, prior=prior, regularization=self.dynamics_regularization)
        self.costs = cost_function.quadratic_approximation(np.mean(X, axis=0), np.mean(U, axis=0))
        ilqr = ILQR(self.dyn, prev_pol, self.costs, self.x0)
        self.pol = ilqr.optimize(self.kl_step * self.kl_step_mult).policy
        datalogging.log(prev_pol=prev_pol, pol=self.pol, dyn=self.dyn, x0=self.x0, costs=self.costs,
                        kl_step=self.kl_step * self.kl_step_mult)
        if prev_dyn is not None:
            with datalogging.Context("step_adjust"):
                self.kl_step_mult = lqg.step_adjust(self.kl_step_mult, self.dyn, self.pol, self.x0, self.costs,
                                                    prev_dyn, prev_pol, prev_x0, prev_costs, self.max_step_mult,
                                                    self.min_step_mult)
