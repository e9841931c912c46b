from donk_b200.costs.losses import loss_combined, loss_l1, loss_l2, loss_log_cosh
from donk_b200.costs.quadratic_costs import QuadraticCosts

__all__ = ["QuadraticCosts", "loss_l2", "loss_l1", "loss_log_cosh", "loss_combined"]
