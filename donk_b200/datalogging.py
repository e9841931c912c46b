"""Call-compatible logging hook (donk/datalogging/datalogging.py:84-97).

The reference logs intermediate results through a module-global logger that is a no-op unless the user
installs one.  Observability is out of scope for the hot path; this keeps the call
sites (`log(**kw)`, `Context(name)`) working and lets a user plug in any callable.
"""
from __future__ import annotations

_sink = None
_context: list[str] = []


def set_sink(fn) -> None:
    """``fn(context: tuple[str, ...], **values)`` receives every ``log`` call; ``None`` disables."""
    global _sink
    _sink = fn


def log(**kwargs) -> None:
    if _sink is not None:
        _sink(tuple(_context), **kwargs)


class Context:
    def __init__(self, name: str) -> None:
        self.name = name

    def __enter__(self):
        _context.append(self.name)
        return self

    def __exit__(self, *exc):
        _context.pop()
        return False
