from donk_b200.dynamics.prior.prior import DynamicsPrior, NormalInverseWishart
from donk_b200.dynamics.prior.prior_gmm import GMMPrior

__all__ = ["DynamicsPrior", "GMMPrior", "NormalInverseWishart"]
