"""Normally distributed state (donk/samples/state_dist.py:6-37)."""
from __future__ import annotations

import numpy as np

from donk_b200.utils.shapes import require_shape


class StateDistribution:
    """``N(mean (..., dX), covar (..., dX, dX))``; leading axes are batch axes and must agree."""

    def __init__(self, mean, covar) -> None:
        self.dX = mean.shape[-1]
        require_shape("covar", covar, mean.shape[:-1] + (self.dX, self.dX))
        self.mean, self.covar = mean, covar

    @staticmethod
    def fit(X) -> "StateDistribution":
        """Mean and unbiased covariance of ``X (N,dX)`` (state_dist.py:27-37), on the device."""
        from donk_b200.batched import state_fit

        mean, covar = state_fit(np.asarray(X, dtype=np.float64)[None])
        return StateDistribution(mean[0].cpu().numpy(), covar[0].cpu().numpy())
