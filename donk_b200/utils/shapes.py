"""Shape validation shared by the value types of the host-side API mirror."""
from __future__ import annotations


def require_shape(what: str, array, expected: tuple, trailing: bool = False) -> None:
    """``AssertionError`` (the exception type the reference's holders raise) unless ``array`` has the expected shape.

    ``trailing``: compare only the last ``len(expected)`` axes (holders that accept leading batch axes).
    """
    got = tuple(array.shape[-len(expected):]) if trailing else tuple(array.shape)
    if got != tuple(expected):
        raise AssertionError(f"{what}: shape {tuple(array.shape)} does not match {'(...,) + ' if trailing else ''}{tuple(expected)}")
