"""Gaussian trajectory distribution over ``[x_t | u_t]`` (container semantics of donk/samples/traj_dist.py:6-71).

``mean (..., T+1, dX+dU)`` and ``covar (..., T+1, dX+dU, dX+dU)`` as produced by the forward pass; the state and action
blocks are exposed as views.  Leading batch dimensions are allowed.
"""
from __future__ import annotations

import numpy as np


class TrajectoryDistribution:
    def __init__(self, mean, covar, dX: int, dU: int = None) -> None:
        p = mean.shape[-1]
        self.dX = dX
        self.dU = p - dX if dU is None else dU
        width = self.dX + self.dU
        assert mean.shape[-1:] == (width,), f"{mean.shape[-1:]} != {(width,)}"
        assert covar.shape[-2:] == (width, width), f"{covar.shape[-2:]} != {(width, width)}"
        assert mean.shape[:-1] == covar.shape[:-2], f"{mean.shape[:-1]} != {covar.shape[:-2]}"
        self.mean, self.covar = mean, covar

    # ---- views of the state / action blocks (actions exist for t < T only) ---------------------------------------
    def _states(self) -> slice:
        return slice(None, self.dX)

    def _actions(self) -> slice:
        return slice(self.dX, None)

    X_mean = property(lambda self: self.mean[..., self._states()])
    X_covar = property(lambda self: self.covar[..., self._states(), self._states()])
    U_mean = property(lambda self: self.mean[..., :-1, self._actions()])
    U_covar = property(lambda self: self.covar[..., :-1, self._actions(), self._actions()])

    def sample(self, size, rng):
        """Roll-outs ``X size+(T+1,dX)``, ``U size+(T,dU)`` drawn from the per-step Gaussians.  One joint draw per
        step t < T and one state draw at T, in that order, so a seeded generator yields the reference's samples."""
        steps = self.mean.shape[-2] - 1
        count = int(np.prod(size))
        joint = np.stack([rng.multivariate_normal(self.mean[t], self.covar[t], size=count) for t in range(steps)], axis=1)
        last = rng.multivariate_normal(self.X_mean[steps], self.X_covar[steps], size=count)
        X = np.concatenate([joint[..., : self.dX], last[:, None, :]], axis=1)
        U = joint[..., self.dX:]
        shape = tuple(size)
        return X.reshape(shape + X.shape[1:]), np.ascontiguousarray(U).reshape(shape + U.shape[1:])
