"""Normally distributed trajectory (donk/samples/traj_dist.py:6-49)."""
from __future__ import annotations

import numpy as np


class TrajectoryDistribution:
    def __init__(self, mean, covar, dX: int, dU: int = None) -> None:
        if dU is None:
            dU = mean.shape[-1] - dX
        assert mean.shape[-1:] == (dX + dU,), f"{mean.shape[-1:]} != {(dX + dU,)}"
        assert covar.shape[-2:] == (dX + dU, dX + dU), f"{covar.shape[-2:]} != {(dX + dU, dX + dU)}"
        assert mean.shape[:-1] == covar.shape[:-2], f"{mean.shape[:-1]} != {covar.shape[:-2]}"
        self.mean = mean
        self.covar = covar
        self.dX = dX
        self.dU = dU

    @property
    def X_mean(self):
        return self.mean[..., : self.dX]

    @property
    def X_covar(self):
        return self.covar[..., : self.dX, : self.dX]

    @property
    def U_mean(self):
        return self.mean[..., :-1, self.dX :]

    @property
    def U_covar(self):
        return self.covar[..., :-1, self.dX :, self.dX :]

    def sample(self, size, rng):
        """Roll-outs drawn step by step from the per-step Gaussians (traj_dist.py:51-71): ``X size+(T+1,dX)``,
        ``U size+(T,dU)``; the random stream is consumed in the reference's order."""
        T = self.mean.shape[-2] - 1
        n = int(np.prod(size))
        X = np.empty((n, T + 1, self.dX))
        U = np.empty((n, T, self.dU))
        for t in range(T):
            xu = rng.multivariate_normal(self.mean[t], self.covar[t], size=n)
            X[:, t], U[:, t] = xu[:, :self.dX], xu[:, self.dX:]
        X[:, T] = rng.multivariate_normal(self.X_mean[T], self.X_covar[T], size=n)
        return X.reshape(tuple(size) + (T + 1, self.dX)), U.reshape(tuple(size) + (T, self.dU))
