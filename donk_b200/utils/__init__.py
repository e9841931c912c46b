from donk_b200.utils.batched import regularize, symmetrize, trace_of_product

__all__ = ["symmetrize", "regularize", "trace_of_product"]
