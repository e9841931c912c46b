"""Normally distributed state (donk/samples/state_dist.py:6-37)."""
from __future__ import annotations

import numpy as np


class StateDistribution:
    def __init__(self, mean, covar) -> None:
        dX = mean.shape[-1]
        assert covar.shape[-2:] == (dX, dX), f"{covar.shape[-2:]} != {(dX, dX)}"
        assert mean.shape[:-1] == covar.shape[:-2], f"{mean.shape[:-1]} != {covar.shape[:-2]}"
        self.mean = mean
        self.covar = covar
        self.dX = dX

    @staticmethod
    def fit(X) -> "StateDistribution":
        """Mean and unbiased covariance of ``X (N,dX)`` (state_dist.py:27-37), on the device."""
        from donk_b200.batched import state_fit

        mean, covar = state_fit(np.asarray(X, dtype=np.float64)[None])
        return StateDistribution(mean[0].cpu().numpy(), covar[0].cpu().numpy())
