"""Sample containers (public names of ``donk.samples`` that the hot path touches); the device-resident pool is
``donk_b200.batched.DeviceTransitionPool``."""
from . import pool as _pool
from . import state_dist as _state
from . import traj_dist as _traj

TransitionPool = _pool.TransitionPool
StateDistribution = _state.StateDistribution
TrajectoryDistribution = _traj.TrajectoryDistribution

__all__ = ("StateDistribution", "TrajectoryDistribution", "TransitionPool")
