"""Dynamics priors (public names of ``donk.dynamics.prior``): the normal-inverse-Wishart value type on the host and the
GMM prior whose EM runs on the device."""
from . import prior as _prior
from . import prior_gmm as _prior_gmm

DynamicsPrior = _prior.DynamicsPrior
NormalInverseWishart = _prior.NormalInverseWishart
GMMPrior = _prior_gmm.GMMPrior

__all__ = ("GMMPrior", "NormalInverseWishart", "DynamicsPrior")
