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


@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    """The kernels under test: ``emu`` = host emulation of the kernel sources (CPU container, checks kernel
    logic + host layer), ``cuda`` = the real sm_100a library through the C ABI (the parity tests proper,
    selected with ``-m gpu``).  The same assertions run against both.
    """
    from donk_b200 import _lib

    if request.param == "cuda":
        _lib.install_test_backend(None)
        nat = _lib.native()  # raises if the CUDA library is missing: no silent fallback
        yield nat
        return
    from emu.build_emu import backend as emu_backend

    nat = emu_backend()
    _lib.install_test_backend(nat)
    yield nat
    nat.lib.donk_emu_set_reverse(0)
    _lib.install_test_backend(None)
