"""Policies of the hot path (``donk.policy``): the time-varying linear-Gaussian controller iLQR produces."""
from . import lin_gaussian as _lg

LinearGaussianPolicy = _lg.LinearGaussianPolicy

__all__ = ("LinearGaussianPolicy",)
