"""CPU oracle for the donk.ai trajectory-optimisation hot path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy/scipy restatement of the
reference algorithm (fit_lr, NIW/GMM prior, GMM EM, LQR backward/forward, KL,
Brent eta search).  It exists to check the CUDA path and to serve as the CPU
baseline leg of ``bench.py``.  Nothing under ``donk_b200/`` may import it; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs do.

Pinning: every function here is checked in ``tests/test_oracle_golden.py``
against full-precision vectors produced by the *unmodified* reference
(``tests/golden/make_golden.py`` imports ``/root/reference`` in the build
container, re-runs the reference's own known-answer tests, and stores inputs +
outputs under ``tests/golden/*.npz``).  Third-party arithmetic that is not in
the reference tree (scikit-learn 1.9.0 ``GaussianMixture`` EM, scipy 1.18.1
``optimize.brentq``) is restated from the published algorithm and pinned
against those libraries' live output in the same golden files.

All citations ``file:line`` are relative to the reference repository root.
"""
from oracle.lqg import (  # noqa: F401
    backward,
    brentq,
    extended_costs_kl,
    forward,
    expected_costs,
    ilqr_optimize,
    ilqr_step,
    kl_divergence_action,
    step_adjust,
)
from oracle.dynamics import fit_lr, niw_map_covar, state_fit, to_transitions  # noqa: F401
from oracle.gmm import GMMParams, gmm_e_step, gmm_eval_prior, gmm_fit, gmm_m_step  # noqa: F401
