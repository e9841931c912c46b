"""The reference's OWN unit tests, run against this package.

When the reference checkout is present (``/root/reference``; it is not on the GPU box, so this is a CPU test), its
hot-path test modules are executed in a subprocess in which the name ``donk`` is bound to ``donk_b200`` (over the
host emulation of the kernels).  Nothing of the reference's implementation is imported: only its test files and their
data.  Expected outcome: all 25 test functions (88 sub-tests) pass, including the sub-tests of ``Test_LinearDynamics.test_fit_lr`` that
assert ``eigvalsh(covar[t]) >= -1e-16`` on a covariance that is singular by construction (N = 15 samples, p = 17).
"""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

from donk_testutil import ROOT

REF = Path("/root/reference")
MODULES = ["tests/traj_opt_test.py", "tests/dynamics_test.py", "tests/samples_test.py", "tests/utils_test.py"]

DRIVER = textwrap.dedent('''
    import importlib, sys, types
    sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
    import donk_b200
    from donk_b200 import _lib
    from emu.build_emu import backend
    _lib.install_test_backend(backend())
    alias = types.ModuleType("donk")
    alias.__path__ = list(donk_b200.__path__)
    sys.modules["donk"] = alias
    for name in ("dynamics", "dynamics.linear_dynamics", "dynamics.prior", "dynamics.utils", "dynamics.dynamics", "traj_opt",
                 "traj_opt.lqg", "traj_opt.traj_opt", "policy", "policy.lin_gaussian", "models", "models.tvlg", "costs",
                 "costs.quadratic_costs", "costs.losses", "samples", "samples.pool", "samples.state_dist", "samples.traj_dist",
                 "utils", "utils.batched", "datalogging"):
        mod = importlib.import_module("donk_b200." + name)
        sys.modules["donk." + name] = mod
        if "." not in name:
            setattr(alias, name, mod)
    import donk  # the alias
    assert "reference" not in str(getattr(sys.modules["donk.traj_opt"], "__file__", ""))
    import pytest
    sys.exit(pytest.main(["-p", "no:cacheprovider", "-q", "--tb=line", "-rf"] + {modules!r}))
''')


@pytest.mark.skipif(not (REF / "tests" / "traj_opt_test.py").exists(), reason="reference checkout not present")
def test_reference_unit_tests_pass_against_this_package():
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    env.pop("DONK_GMM_MMA", None)
    # tests/dynamics_test.py:102-178 pins values that depend on scikit-learn's k-means random stream: the opt-in host
    # initialisation reproduces them; the package default (device k-means) agrees statistically only
    env["DONK_GMM_INIT"] = "sklearn"
    code = DRIVER.format(root=str(ROOT), modules=MODULES)
    out = subprocess.run([sys.executable, "-c", code], cwd=REF, env=env, capture_output=True, text=True, timeout=900)
    text = out.stdout + out.stderr
    failed = [ln for ln in text.splitlines() if ln.startswith(("FAILED", "SUBFAILED", "ERROR"))]
    assert failed == [], text[-3000:]
    summary = [ln for ln in text.splitlines() if " passed" in ln][-1]
    passed = int(summary.split(" passed")[0].split()[-1])
    assert passed >= 25 and "88 subtests passed" in summary, summary  # what the reference scores against itself
