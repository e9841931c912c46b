from donk_b200.traj_opt.lqg import ILQR, ILQRStepResult
from donk_b200.traj_opt.traj_opt import TrajOptAlgorithm

__all__ = ["ILQR", "ILQRStepResult", "TrajOptAlgorithm"]
