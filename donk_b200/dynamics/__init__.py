from donk_b200.dynamics.dynamics import DynamicsModel
from donk_b200.dynamics.linear_dynamics import LinearDynamics
from donk_b200.dynamics.utils import to_transitions

__all__ = ["DynamicsModel", "LinearDynamics", "to_transitions"]
