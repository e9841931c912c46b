"""Dynamics models of the hot path (public names of ``donk.dynamics``): ``fit_lr`` lives in ``.linear_dynamics``."""
from . import dynamics as _dynamics
from . import linear_dynamics as _linear
from . import utils as _utils

DynamicsModel = _dynamics.DynamicsModel
LinearDynamics = _linear.LinearDynamics
to_transitions = _utils.to_transitions

__all__ = ("LinearDynamics", "DynamicsModel", "to_transitions")
