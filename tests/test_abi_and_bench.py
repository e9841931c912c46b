"""C-ABI surface, loader behaviour without a GPU, bench.py contract (dry run on the emulation)."""
import ctypes
import json
import re
import subprocess
import sys

import pytest
import torch

from donk_testutil import ROOT


def _header_symbols():
    text = (ROOT / "include" / "donk_b200.h").read_text()
    return sorted(set(re.findall(r"\b(donk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    """The in-tree CUDA library loads (dlopen needs no device) and exports all of include/donk_b200.h."""
    from donk_b200 import _lib

    if not _lib.LIB_PATH.exists():
        import __graft_entry__ as g

        g.build()
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    syms = _header_symbols()
    assert len(syms) >= 19
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    lib.donk_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.donk_version()
    # the python binding table covers the header
    bound = {"donk_" + n for n in list(_lib.SIGNATURES) + list(_lib._SIZE_FUNCS)} | {
        "donk_version", "donk_last_error", "donk_launch_count"}
    assert set(syms) == bound


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_gpu():
    """No CPU fallback: without a CUDA device every product entry point raises."""
    import numpy as np

    from donk_b200 import _lib
    from donk_b200.dynamics.linear_dynamics import fit_lr

    _lib.install_test_backend(None)
    with pytest.raises((_lib.DonkError, ImportError)):
        fit_lr(np.zeros((3, 5, 2)), np.zeros((3, 4, 1)))
    with pytest.raises((_lib.DonkError, ImportError)):
        _lib.native()


def test_product_code_never_imports_the_oracle():
    for path in (ROOT / "donk_b200").rglob("*.py"):
        text = path.read_text()
        assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), path


def test_bench_contract_dry_run():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--dry-run-emu", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["dtype"] == "f64" and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert line["parity_max_rel_err_vs_oracle"] < 1e-8


def test_bench_reference_arm():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0
    # the unmodified reference when it is installed under baseline/_ref (pip --target), else the numpy port
    expect = "reference" if (ROOT / "baseline" / "_ref" / "donk" / "traj_opt" / "lqg.py").exists() else "port"
    assert line["cpu_baseline"]["kind"] == expect and line["config"]["workload"].startswith("batched ILQR.step")
    assert line["e2e"]["h2d_bytes_per_step"] == 0
