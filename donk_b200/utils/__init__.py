"""Small batched helpers with the in-place semantics the reference's tests pin (public names of ``donk.utils``)."""
from . import batched as _batched

symmetrize = _batched.symmetrize
regularize = _batched.regularize
trace_of_product = _batched.trace_of_product

__all__ = ("trace_of_product", "regularize", "symmetrize")
