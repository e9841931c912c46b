from donk_b200.policy.lin_gaussian import LinearGaussianPolicy

__all__ = ["LinearGaussianPolicy"]
