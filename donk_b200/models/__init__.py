from donk_b200.models.tvlg import TimeVaryingLinearGaussian

__all__ = ["TimeVaryingLinearGaussian"]
