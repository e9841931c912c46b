"""Normal-inverse-Wishart prior over linear dynamics (donk/dynamics/prior/prior.py:9-82)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass

import numpy as np


@dataclass
class NormalInverseWishart:
    mu0: np.ndarray
    Phi: np.ndarray
    N_mean: float
    N_covar: float

    def posterior(self, emp_mean, emp_covar, N: float) -> "NormalInverseWishart":
        """Conjugate update after ``N`` (possibly fractional) samples (prior.py:18-36)."""
        dm = emp_mean - self.mu0
        return NormalInverseWishart(
            mu0=(self.N_mean * self.mu0 + N * emp_mean) / (self.N_mean + N),
            Phi=self.Phi + N * emp_covar + self.N_mean * N / (self.N_mean + N) * np.outer(dm, dm),
            N_mean=self.N_mean + N,
            N_covar=self.N_covar + N,
        )

    def map_mean(self) -> np.ndarray:
        return self.mu0.copy()

    def map_covar(self) -> np.ndarray:
        return self.Phi / (self.N_covar + self.mu0.shape[-1] + 1)

    @staticmethod
    def non_informative_prior(d: int) -> "NormalInverseWishart":
        return NormalInverseWishart(mu0=np.zeros((d,)), Phi=np.zeros((d, d)), N_mean=0, N_covar=-(1 + d))


class DynamicsPrior(ABC):
    """A prior over linear dynamics, expressed as normal-inverse-Wishart distributions."""

    @abstractmethod
    def update(self, XUX: np.ndarray) -> None:
        """Refit the prior to transitions ``XUX (N, dX+dU+dX)``."""

    @abstractmethod
    def eval(self, XUX: np.ndarray) -> NormalInverseWishart:
        """Prior for the given transitions."""
