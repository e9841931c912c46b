"""FIFO pool of transition rows on the host (semantics of donk/samples/pool.py:6-56; the device-resident version is
``donk_b200.batched.DeviceTransitionPool``)."""
import numpy as np

from donk_b200.dynamics.utils import to_transitions


class TransitionPool:
    """Keeps rows ``[x_t | u_t | x_{t+1}]`` in arrival order; with a ``capacity`` only the newest rows survive."""

    def __init__(self, capacity: int = None) -> None:
        self.capacity = capacity
        self._XUX = None  # (rows, 2 dX + dU) or None while empty

    def _check(self, X, U) -> None:
        assert X.shape[0] == U.shape[0], f"{X.shape[0]} != {U.shape[0]}"
        assert X.shape[1] == U.shape[1] + 1, f"{X.shape[1]} != {U.shape[1] + 1}"
        if self._XUX is None:
            self.dX, self.dU = X.shape[-1], U.shape[-1]  # the first batch fixes the dimensions
        else:
            assert X.shape[-1] == self.dX, f"{X.shape[-1]} != {self.dX}"
            assert U.shape[-1] == self.dU, f"{U.shape[-1]} != {self.dU}"

    def add(self, X, U) -> None:
        """Append the transitions of roll-outs ``X (N,T+1,dX)``, ``U (N,T,dU)``; drop the oldest rows beyond capacity."""
        self._check(X, U)
        rows = to_transitions(X, U)
        merged = rows if self._XUX is None else np.concatenate((self._XUX, rows))
        self._XUX = merged if self.capacity is None else merged[len(merged) - min(len(merged), self.capacity):]

    def get_transitions(self, N: int = None):
        """All rows, or the ``N`` newest, as a view that cannot be written through."""
        rows = self._XUX if N is None else self._XUX[-N:]
        view = rows.view()
        view.setflags(write=False)
        return view
