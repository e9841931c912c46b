"""Time-varying linear-Gaussian policy (donk/policy/lin_gaussian.py:7-46)."""
from donk_b200.models import TimeVaryingLinearGaussian


class LinearGaussianPolicy(TimeVaryingLinearGaussian):
    """``p(u|t,x) = N(K_t x + k_t, covar_t)``; give ``covar``, ``inv_covar`` or both."""

    def __init__(self, K, k, covar=None, inv_covar=None):
        TimeVaryingLinearGaussian.__init__(self, K, k, covar, inv_covar)
        _, self.dU, self.dX = K.shape

    K = TimeVaryingLinearGaussian.coefficients
    k = TimeVaryingLinearGaussian.intercept

    def act(self, x, t: int, noise=None):
        return self.predict(x, t, noise)

    def __str__(self) -> str:
        return f"LinearGaussianPolicy[T={self.T}, dX={self.dX}, dU={self.dU}]"
