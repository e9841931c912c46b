"""Normally distributed trajectory (donk/samples/traj_dist.py:6-49)."""
from __future__ import annotations


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
