"""Helpers shared by the test modules (import as ``from donk_testutil import ...``)."""
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"


def relerr(ours, ref):
    """Max-norm relative error per tensor: max|ours-ref| / max|ref|."""
    ours = np.asarray(to_np(ours), dtype=np.float64)
    ref = np.asarray(to_np(ref), dtype=np.float64)
    assert ours.shape == ref.shape, f"{ours.shape} != {ref.shape}"
    if ref.size == 0:
        return 0.0
    scale = np.max(np.abs(ref))
    if scale == 0:
        scale = 1.0
    return float(np.max(np.abs(ours - ref)) / scale)


def load_golden(name):
    with np.load(GOLDEN / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


def set_reverse(backend, reverse: int):
    """Emulation only: run every kernel phase back to front (exposes intra-phase dependencies)."""
    if backend.device_type != "cpu":
        return  # a real GPU has one execution order: the parametrisation runs (forward) so its shapes stay covered
    backend.lib.donk_emu_set_reverse(reverse)


def to_np(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else t
