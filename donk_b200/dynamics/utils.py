"""Roll-out -> transition-tuple layout (donk/dynamics/utils.py:4-17)."""
import numpy as np


def to_transitions(X: np.ndarray, U: np.ndarray) -> np.ndarray:
    """``X (N,T+1,dX), U (N,T,dU) -> XUX (N*T, dX+dU+dX)``; row ``n*T+t = [x_t | u_t | x_{t+1}]``.

    Host-side layout helper for numpy callers (``TransitionPool``); device-resident roll-outs use
    ``donk_b200.batched.to_transitions`` (``donk_to_transitions_f64``).
    """
    N, T, _ = U.shape
    return np.concatenate([X[:, :-1], U, X[:, 1:]], axis=-1).reshape(N * T, -1)
