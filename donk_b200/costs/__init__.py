from donk_b200.costs.quadratic_costs import QuadraticCosts

__all__ = ["QuadraticCosts"]
