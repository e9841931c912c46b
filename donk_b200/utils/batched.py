"""Small-matrix helpers with the reference's in-place semantics (donk/utils/batched.py:6-84).

These are host-side conveniences on numpy arrays (the reference tests pin that ``symmetrize`` and
``regularize`` modify and return their argument, tests/utils_test.py:30-101).  ``batched_cholesky`` and
``batched_inv_spd`` run on the device through ``donk_spd_factor_f64``.
"""
import numpy as np


def symmetrize(A: np.ndarray) -> np.ndarray:
    """In-place ``(A + A^T) / 2`` on the last two axes; returns ``A`` (batched.py:52-59)."""
    A += np.swapaxes(A, -1, -2).copy()
    A /= 2
    return A


def regularize(A: np.ndarray, regularization: float) -> np.ndarray:
    """In-place diagonal shift; returns ``A`` (batched.py:62-69)."""
    idx = np.arange(A.shape[-1])
    A[..., idx, idx] += regularization
    return A


def trace_of_product(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """``Tr(A @ B)`` without forming the product (batched.py:72-84)."""
    return np.einsum("...ij,...ji->...", A, B)


def batched_cholesky(a: np.ndarray) -> np.ndarray:
    """Lower Cholesky factors of ``a (T,d,d)`` (batched.py:34-49); raises LinAlgError if not PD."""
    from donk_b200.batched import spd_factor

    _, chol, _ = spd_factor(a, want_inv=False, want_chol=True, want_hld=False)
    return chol.cpu().numpy()


def batched_inv_spd(a_chol: np.ndarray) -> np.ndarray:
    """Inverse of ``L L^T`` from the Cholesky factors ``a_chol (T,d,d)`` (batched.py:6-31)."""
    from donk_b200.batched import spd_factor

    a_chol = np.asarray(a_chol, dtype=np.float64)
    inv, _, _ = spd_factor(a_chol @ np.swapaxes(a_chol, -1, -2), want_hld=False)
    return inv.cpu().numpy()
