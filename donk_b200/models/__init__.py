"""Model containers of the hot path: the time-varying linear-Gaussian holder that ``LinearDynamics`` and
``LinearGaussianPolicy`` derive from (mirrors the public names of ``donk.models``)."""
from . import tvlg as _tvlg

TimeVaryingLinearGaussian = _tvlg.TimeVaryingLinearGaussian

__all__ = ("TimeVaryingLinearGaussian",)
