"""Abstract dynamics model (donk/dynamics/dynamics.py:6-26)."""
from abc import ABC, abstractmethod


class DynamicsModel(ABC):
    @abstractmethod
    def predict(self, x, u, t: int = None, noise=None):
        """Next state for state ``x`` and action ``u`` at step ``t``."""
