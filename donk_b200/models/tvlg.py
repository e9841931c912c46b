"""Time-varying linear-Gaussian model holder (donk/models/tvlg.py:6-129)."""
import numpy as np


class TimeVaryingLinearGaussian:
    """``p(y|t,x) = N(coefficients_t x + intercept_t, covar_t)`` with lazily derived factors.

    Same contract as the reference: exactly ``(T,dY,dX)`` coefficients (tvlg.py:28-36), ``covar`` /
    ``chol_covar`` / ``inv_covar`` computed on first use (tvlg.py:62-99) — here on the device by
    ``donk_spd_factor_f64`` — and ``numpy.linalg.LinAlgError`` if a covariance is not positive definite.
    """

    def __init__(self, coefficients, intercept, covar=None, inv_covar=None) -> None:
        if covar is None and inv_covar is None:
            raise ValueError("Must provide covar or inv_covar.")
        self.T, dY, dX = coefficients.shape
        assert coefficients.shape == (self.T, dY, dX), f"{coefficients.shape} != {(self.T, dY, dX)}"
        assert intercept.shape == (self.T, dY), f"{intercept.shape} != {(self.T, dY)}"
        if covar is not None:
            assert covar.shape == (self.T, dY, dY), f"{covar.shape} != {(self.T, dY, dY)}"
        if inv_covar is not None:
            assert inv_covar.shape == (self.T, dY, dY), f"{inv_covar.shape} != {(self.T, dY, dY)}"
        self._coefficients = coefficients
        self._intercept = intercept
        self._covar = covar
        self._chol_covar = None
        self._inv_covar = inv_covar

    @property
    def coefficients(self):
        return self._coefficients

    @property
    def intercept(self):
        return self._intercept

    @staticmethod
    def _factor(a, **kw):
        from donk_b200.batched import spd_factor

        return [None if r is None else r.cpu().numpy() for r in spd_factor(np.asarray(a, dtype=np.float64), **kw)]

    @property
    def covar(self):
        if self._covar is None:
            self._covar = self._factor(self._inv_covar, want_hld=False)[0]
        return self._covar

    @property
    def chol_covar(self):
        if self._chol_covar is None:
            self._chol_covar = self._factor(self.covar, want_inv=False, want_chol=True, want_hld=False)[1]
        return self._chol_covar

    @property
    def inv_covar(self):
        if self._inv_covar is None:
            self._inv_covar = self._factor(self.covar, want_hld=False)[0]
        return self._inv_covar

    def predict(self, X, t: int = None, noise=None):
        """Model output for inputs ``X`` at step ``t`` (or all steps), tvlg.py:101-129."""
        X = np.asarray(X)
        if t is not None:
            Y = np.einsum("ux,...x->...u", self.coefficients[t], X) + self.intercept[t]
            if noise is not None:
                Y = Y + np.einsum("ij,...j->...i", self.chol_covar[t], noise)
        else:
            dX = self.coefficients.shape[-1]
            if X.shape[-2:] != (self.T, dX):
                raise ValueError(f"Input must have shape (..., {self.T}, {dX}), not {X.shape}")
            Y = np.einsum("tyx,...tx->...ty", self.coefficients, X) + self.intercept
            if noise is not None:
                Y = Y + np.einsum("tij,...tj->...ti", self.chol_covar, noise)
        return Y
