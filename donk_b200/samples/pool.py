"""FIFO pool of transition tuples (donk/samples/pool.py:6-56)."""
import numpy as np

from donk_b200.dynamics.utils import to_transitions


class TransitionPool:
    """Stores ``[x_t | u_t | x_{t+1}]`` rows; the newest ``capacity`` rows are kept."""

    def __init__(self, capacity: int = None) -> None:
        self._XUX = None
        self.capacity = capacity

    def add(self, X, U) -> None:
        """Append the transitions of roll-outs ``X (N,T+1,dX), U (N,T,dU)`` (pool.py:18-42)."""
        assert X.shape[0] == U.shape[0], f"{X.shape[0]} != {U.shape[0]}"
        assert X.shape[1] == U.shape[1] + 1, f"{X.shape[1]} != {U.shape[1] + 1}"
        if self._XUX is None:
            self.dX, self.dU = X.shape[-1], U.shape[-1]
            self._XUX = to_transitions(X, U)
        else:
            assert X.shape[-1] == self.dX, f"{X.shape[-1]} != {self.dX}"
            assert U.shape[-1] == self.dU, f"{U.shape[-1]} != {self.dU}"
            self._XUX = np.concatenate([self._XUX, to_transitions(X, U)], axis=0)
        if self.capacity is not None:
            self._XUX = self._XUX[-self.capacity :]

    def get_transitions(self, N: int = None):
        """All, or the ``N`` newest, transitions as a read-only view (pool.py:44-56)."""
        XUX = self._XUX if N is None else self._XUX[-N:]
        XUX = XUX.view()
        XUX.flags.writeable = False
        return XUX
