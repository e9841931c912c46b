"""Time-varying linear-Gaussian model holder (donk/models/tvlg.py:6-129)."""
from __future__ import annotations

import numpy as np

from donk_b200.utils.shapes import require_shape


class _SpdForms:
    """The three forms of a batch of SPD matrices — the matrix, its lower Cholesky factor and its inverse —
    derived from whichever was given, on first use, by ``donk_spd_factor_f64`` on the device."""

    def __init__(self, matrix=None, inverse=None):
        self._forms = {"matrix": matrix, "inverse": inverse, "chol": None}

    def _factor(self, of, **want):
        from donk_b200.batched import spd_factor

        inv, chol, _ = spd_factor(np.asarray(of, dtype=np.float64), want_hld=False, **want)
        return inv, chol

    def get(self, form: str):
        have = self._forms
        if have[form] is None:
            if form == "matrix":  # only the inverse was given
                have["matrix"] = self._factor(have["inverse"])[0].cpu().numpy()
            elif form == "chol":
                have["chol"] = self._factor(self.get("matrix"), want_inv=False, want_chol=True)[1].cpu().numpy()
            else:
                have["inverse"] = self._factor(self.get("matrix"))[0].cpu().numpy()
        return have[form]


class TimeVaryingLinearGaussian:
    """``p(y|t,x) = N(coefficients_t x + intercept_t, covar_t)`` with lazily derived factors.

    Contract of the reference: coefficients are exactly ``(T,dY,dX)`` (tvlg.py:28-36), at least one of ``covar`` /
    ``inv_covar`` is given, the missing forms and ``chol_covar`` appear on first use (tvlg.py:62-99), and a covariance
    that is not positive definite raises ``numpy.linalg.LinAlgError``.
    """

    def __init__(self, coefficients, intercept, covar=None, inv_covar=None) -> None:
        if covar is None and inv_covar is None:
            raise ValueError("Must provide covar or inv_covar.")
        if coefficients.ndim != 3:
            raise AssertionError(f"coefficients must be (T, dY, dX), got shape {tuple(coefficients.shape)}")
        self.T, dY, _ = coefficients.shape
        require_shape("intercept", intercept, (self.T, dY))
        for name, m in (("covar", covar), ("inv_covar", inv_covar)):
            if m is not None:
                require_shape(name, m, (self.T, dY, dY))
        self._lin = (coefficients, intercept)
        self._spd = _SpdForms(matrix=covar, inverse=inv_covar)

    coefficients = property(lambda self: self._lin[0])
    intercept = property(lambda self: self._lin[1])
    covar = property(lambda self: self._spd.get("matrix"))
    chol_covar = property(lambda self: self._spd.get("chol"))
    inv_covar = property(lambda self: self._spd.get("inverse"))

    def given_forms(self):
        """``(covar, inv_covar)`` as known so far, ``None`` for a form that has not been given or derived yet
        (lets a caller ship only what exists and derive the rest on the device in the same batch)."""
        return self._spd._forms["matrix"], self._spd._forms["inverse"]

    def predict(self, X, t: int = None, noise=None):
        """Model output for inputs ``X (..., dX)`` at step ``t``, or ``X (..., T, dX)`` for all steps (tvlg.py:101-129);
        ``noise`` (standard normal, output shaped) is coloured by ``chol_covar``."""
        X = np.asarray(X)
        A, b = self._lin
        if t is None:
            if X.shape[-2:] != A.shape[::2]:
                raise ValueError(f"Input must have shape (..., {self.T}, {A.shape[2]}), not {X.shape}")
            L = self.chol_covar if noise is not None else None
        else:
            A, b = A[t], b[t]
            L = self.chol_covar[t] if noise is not None else None
        Y = np.matmul(A, X[..., None])[..., 0] + b
        if L is not None:
            Y = Y + np.matmul(L, np.asarray(noise)[..., None])[..., 0]
        return Y
