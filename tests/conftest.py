import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

if str(ROOT / "tests") not in sys.path:
    sys.path.insert(0, str(ROOT / "tests"))

from donk_testutil import GOLDEN, relerr  # noqa: E402,F401


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        with np.load(GOLDEN / f"{name}.npz") as z:
            return {k: z[k] for k in z.files}

    return load


@pytest.fixture()
def emu():
    """Route donk_b200 calls to the host-emulation build of the kernel sources (no GPU needed).

    This checks kernel logic and the Python host layer on CPU; the `-m gpu` tests run the same
    assertions against the real CUDA library.
    """
    from donk_b200 import _lib
    from emu.build_emu import backend

    nat = backend()
    _lib.install_test_backend(nat)
    yield nat
    nat.lib.donk_emu_set_reverse(0)
    _lib.install_test_backend(None)
