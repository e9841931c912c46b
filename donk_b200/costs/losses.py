"""Per-time-step penalties with their first and second derivatives (donk/costs/losses.py:6-136).

Host-side cost authoring: these run once per GPS iteration on (T, D) arrays and feed
``QuadraticCosts``; they are not part of the device path.  Each function returns ``(l (T,), lx (T,D), lxx (T,D,D))``.
"""
from __future__ import annotations

import numpy as np


def _diag(v: np.ndarray) -> np.ndarray:
    """(T, D) -> (T, D, D) with ``v[t]`` on the diagonal of slice ``t``."""
    out = np.zeros(v.shape + (v.shape[-1],), dtype=v.dtype)
    idx = np.arange(v.shape[-1])
    out[..., idx, idx] = v
    return out


def loss_l2(x, t, w):
    """``sum_j 0.5 w_j (x_j - t_j)^2`` (losses.py:28-60)."""
    err = x - t
    return 0.5 * np.sum(w * err * err, axis=1), w * err, _diag(np.asarray(w, dtype=np.result_type(w, err)))


def loss_l1(x, t, w, alpha):
    """Smoothed absolute error ``sum_j w_j sqrt((x_j - t_j)^2 + alpha)`` (losses.py:63-97)."""
    err = x - t
    r = np.sqrt(err * err + alpha)
    return np.sum(w * r, axis=1), w * err / r, _diag(w * alpha / r**3)


def loss_log_cosh(x, t, w):
    """``sum_j w_j log cosh(x_j - t_j)`` (losses.py:100-136)."""
    err = x - t
    ch = np.cosh(err)
    return np.sum(w * np.log(ch), axis=1), w * np.tanh(err), _diag(w / (ch * ch))


def loss_combined(x, losses):
    """Sum of several penalties: ``losses`` is a list of ``(function, kwargs)`` (losses.py:6-25)."""
    if len(losses) < 1:
        raise ValueError("loss_combined requred at least one loss function to sum up.")
    total = None
    for fn, kwargs in losses:
        part = fn(x, **kwargs)
        total = list(part) if total is None else [a + b for a, b in zip(total, part)]
    return tuple(total)
