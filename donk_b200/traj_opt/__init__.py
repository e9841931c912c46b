"""KL-constrained trajectory optimisation on the device (public names of ``donk.traj_opt``; the module functions
``backward`` / ``forward`` / ``extended_costs_kl`` / ``kl_divergence_action`` / ``step_adjust`` are in ``.lqg``)."""
from . import lqg as _lqg
from . import traj_opt as _algo

ILQR = _lqg.ILQR
ILQRStepResult = _lqg.ILQRStepResult
TrajOptAlgorithm = _algo.TrajOptAlgorithm

__all__ = ("TrajOptAlgorithm", "ILQR", "ILQRStepResult")
