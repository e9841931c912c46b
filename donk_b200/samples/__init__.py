from donk_b200.samples.pool import TransitionPool
from donk_b200.samples.state_dist import StateDistribution
from donk_b200.samples.traj_dist import TrajectoryDistribution

__all__ = ["TransitionPool", "StateDistribution", "TrajectoryDistribution"]
