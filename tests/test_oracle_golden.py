"""Pin the CPU oracle (oracle/) against vectors produced by the unmodified reference
(tests/golden/make_golden.py; reference tests/traj_opt_test.py, tests/dynamics_test.py)."""
import numpy as np
import pytest

import oracle
from oracle.gmm import GMMParams, params_from_precisions
from oracle.lqg import ILQRProblem
from donk_testutil import relerr

TOL = 1e-10  # oracle vs reference: same LAPACK/BLAS calls, different association order


def test_forward_matches_reference(golden):
    g = golden("lqg_small")
    mean, covar = oracle.forward(g["fw_F"], g["fw_f"], g["fw_dyn_covar"], g["fw_K"], g["fw_k"], g["fw_pol_covar"],
                                 g["fw_x0_mean"], g["fw_x0_covar"])
    assert relerr(mean, g["fw_mean"]) < TOL
    assert relerr(covar, g["fw_covar"]) < TOL
    assert np.array_equal(covar, np.swapaxes(covar, 1, 2))
    # literal known answers of tests/traj_opt_test.py:27-35 (first and last rows)
    np.testing.assert_allclose(mean[0], [0.12573022, -0.13210486, 0.64042265, -2.21332089, 1.2562705], rtol=1e-7)
    np.testing.assert_allclose(mean[5], [31.0077354, 54.39317675, -68.29849519, 0, 0], rtol=1e-7)


def test_backward_matches_reference(golden):
    g = golden("lqg_small")
    K, k, covar, inv = oracle.backward(g["bw_F"], g["bw_f"], g["bw_C"], g["bw_c"])
    for a, b in ((K, "bw_K"), (k, "bw_k"), (covar, "bw_covar"), (inv, "bw_inv_covar")):
        assert relerr(a, g[b]) < TOL
    np.testing.assert_allclose(K[0], [[1.30823423, 1.38208415, 4.50630981], [-0.71211717, -0.19657582, -2.59400007]],
                               rtol=1e-7)  # tests/traj_opt_test.py:82-84
    K, k, covar, _ = oracle.backward(g["bw_F"], g["bw_f"], g["bw_C"], g["bw_c"], gamma=0.9)
    assert relerr(K, g["bwg_K"]) < TOL and relerr(covar, g["bwg_covar"]) < TOL


def test_extended_costs_and_kl(golden):
    g = golden("lqg_small")
    C, c = oracle.extended_costs_kl(g["ex_K"], g["ex_k"], g["ex_inv_covar"])
    assert relerr(C, g["ex_C"]) < TOL and relerr(c, g["ex_c"]) < TOL
    kl = oracle.kl_divergence_action(g["kl_X"], g["kl_K"], g["kl_k"], g["kl_covar"], g["kl_pK"], g["kl_pk"],
                                     g["kl_pcovar"], g["kl_pinv"])
    assert abs(kl - 3.914744335914712) < 1e-12  # tests/traj_opt_test.py:145


def test_fit_lr_traj00(golden):
    g = golden("fitlr_traj00")
    F, f, covar = oracle.fit_lr(g["X"], g["U"], regularization=1e-6)
    assert relerr(F, g["F"]) < 1e-9 and relerr(f, g["f"]) < 1e-9 and relerr(covar, g["covar"]) < 1e-9
    assert np.array_equal(covar, np.swapaxes(covar, 1, 2))
    with pytest.raises(ValueError):
        oracle.fit_lr(g["X"][:1], g["U"][:1])


def _gmm(g, pre="gmm_"):
    return GMMParams(g[pre + "weights"], g[pre + "means"], g[pre + "covariances"], g[pre + "prec_chol"])


def test_fit_lr_with_gmm_prior(golden):
    g = golden("fitlr_traj00")
    prm = _gmm(g)
    n_pool = int(g["gmm_N"])
    mu_p, Phi_p, N_p = oracle.gmm_eval_prior(prm, n_pool, g["ev_xux"])
    d = g["ev_xux"].shape[1]
    assert relerr(mu_p, g["ev_mu0"]) < TOL
    assert relerr(N_p * Phi_p, g["ev_Phi"]) < TOL
    assert abs(N_p - g["ev_N_mean"]) < 1e-9 * N_p and abs(N_p - (1 + d) - g["ev_N_covar"]) < 1e-8
    prior = lambda xux: oracle.gmm_eval_prior(prm, n_pool, xux)  # noqa: E731
    F, f, covar = oracle.fit_lr(g["X"][:3], g["U"][:3], prior, regularization=0)
    assert relerr(F, g["pF"]) < 1e-8 and relerr(f, g["pf"]) < 1e-8 and relerr(covar, g["pcovar"]) < 1e-8
    F, f, covar = oracle.fit_lr(g["X"][:6], g["U"][:6], prior, regularization=1e-6, prior_weight=0.5)
    assert relerr(F, g["wF"]) < 1e-8 and relerr(f, g["wf"]) < 1e-8 and relerr(covar, g["wcovar"]) < 1e-8


def test_gmm_label_init_reproduces_sklearn_fit(golden):
    """k-means labels -> one-hot M-step -> EM with the default stop rule == sklearn fit."""
    g = golden("fitlr_traj00")
    XUX = oracle.to_transitions(g["X"], g["U"])
    prm = oracle.gmm_fit(XUX, g["kmeans_labels"])
    assert prm.n_iter == int(g["gmm_n_iter"])
    assert abs(prm.lower_bound - float(g["gmm_lower_bound"])) < 1e-9 * abs(float(g["gmm_lower_bound"]))
    assert relerr(prm.means, g["gmm_means"]) < 1e-9
    assert relerr(prm.covariances, g["gmm_covariances"]) < 1e-9
    assert relerr(prm.weights, g["gmm_weights"]) < 1e-9


def test_gmm_em_iterations(golden):
    g = golden("gmm_em")
    X = g["X"]
    for it in (1, 2, 3, 5):
        prm = oracle.gmm_fit(X, params_from_precisions(g["w0"], g["m0"], g["p0"]), tol=0, max_iter=it)
        assert relerr(prm.weights, g[f"it{it}_weights"]) < TOL
        assert relerr(prm.means, g[f"it{it}_means"]) < TOL
        assert relerr(prm.covariances, g[f"it{it}_covariances"]) < TOL
        assert relerr(prm.precisions_cholesky, g[f"it{it}_prec_chol"]) < TOL
        assert abs(prm.lower_bound - g[f"it{it}_lower_bound"]) < 1e-10 * abs(g[f"it{it}_lower_bound"])
    prm = oracle.gmm_fit(X, params_from_precisions(g["w0"], g["m0"], g["p0"]))
    assert prm.n_iter == int(g["d_n_iter"]) and prm.converged == bool(g["d_converged"])
    assert relerr(prm.covariances, g["d_covariances"]) < TOL
    warm = oracle.gmm_fit(g["X2"], prm, warm=True)
    assert warm.n_iter == int(g["w_n_iter"])
    assert relerr(warm.means, g["w_means"]) < TOL and relerr(warm.covariances, g["w_covariances"]) < TOL
    _, log_resp = oracle.gmm_e_step(g["pp_rows"], warm)
    assert relerr(np.exp(log_resp), g["pp"]) < TOL


def _problem(g):
    return ILQRProblem(g["F"], g["f"], g["dyn_covar"], g["C"], g["c"], g["cc"], g["prev_K"], g["prev_k"],
                       g["prev_covar"], g["prev_inv_covar"], g["x0_mean"], g["x0_covar"])


@pytest.mark.parametrize("case", ["ilqr_c1", "ilqr_c2"])
def test_ilqr_step_and_optimize(golden, case):
    g = golden(case)
    prob = _problem(g)
    assert relerr(prob.C_kl, g["C_kl"]) < TOL and relerr(prob.c_kl, g["c_kl"]) < TOL
    for i, eta in enumerate(g["etas"]):
        r = oracle.ilqr_step(prob, float(eta))
        assert relerr(r.K, g[f"s{i}_K"]) < 1e-9 and relerr(r.k, g[f"s{i}_k"]) < 1e-9
        assert relerr(r.covar, g[f"s{i}_covar"]) < 1e-9
        assert relerr(r.mean, g[f"s{i}_mean"]) < 1e-9
        assert abs(r.kl_div - g[f"s{i}_kl"]) < 1e-9 * max(1.0, abs(g[f"s{i}_kl"]))
        assert abs(r.expected_costs - g[f"s{i}_cost"]) < 1e-9 * abs(g[f"s{i}_cost"])
        if f"s{i}_tcovar" in g:
            assert relerr(r.traj_covar, g[f"s{i}_tcovar"]) < 1e-9
    for j in (0, 1):
        r, visited = oracle.ilqr_optimize(prob, float(g[f"o{j}_kl_step"]))
        ref_visited = g[f"o{j}_visited"]
        assert len(visited) == len(ref_visited)  # same Brent path length
        np.testing.assert_allclose(visited, ref_visited, rtol=1e-9)
        assert abs(r.eta - g[f"o{j}_eta"]) < 1e-9 * g[f"o{j}_eta"]
        assert relerr(r.K, g[f"o{j}_K"]) < 1e-8 and relerr(r.k, g[f"o{j}_k"]) < 1e-8


def test_fit_lr_synthetic_with_prior(golden):
    g = golden("ilqr_c2")
    prm = _gmm(g)
    prior = lambda xux: oracle.gmm_eval_prior(prm, int(g["gmm_N"]), xux)  # noqa: E731
    F, f, covar = oracle.fit_lr(g["X"], g["U"], prior)
    assert relerr(F, g["F"]) < 1e-9 and relerr(f, g["f"]) < 1e-9 and relerr(covar, g["dyn_covar"]) < 1e-9
    m, S = oracle.state_fit(g["X"][:, 0])
    assert relerr(m, g["x0_mean"]) < TOL and relerr(S, g["x0_covar"]) < TOL


def test_brentq_iterates_match_scipy(golden):
    g = golden("brent")
    funs = {
        "cubic": lambda x: (x - 1.3) ** 3 + 0.2 * (x - 1.3),
        "tanh": lambda x: np.tanh(0.3 * (x - 5.0)) + 0.1,
        "exp": lambda x: np.exp(-0.2 * x) - 0.05,
        "kink": lambda x: (abs(x + 3.0) ** 0.5) * np.sign(x + 3.0) + 0.01 * x,
    }
    for name, fn in funs.items():
        for rtol in (1e-2, 1e-6):
            root, seq = oracle.brentq(fn, -13.8, 36.8, rtol=rtol)
            assert root == float(g[f"{name}_{rtol:g}_root"])
            assert np.array_equal(np.array(seq), g[f"{name}_{rtol:g}_seq"])  # bit-identical iterate sequence


def test_step_adjust_and_costs(golden):
    g = golden("ilqr_c1")
    mult, *_ = oracle.step_adjust(
        1.0, (g["sa_F2"], g["sa_f2"], g["sa_cov2"]), (g["o0_K"], g["o0_k"], g["o0_covar"]),
        (g["x0_mean"], g["x0_covar"]), (g["C"], g["c"], g["cc"]), (g["F"], g["f"], g["dyn_covar"]),
        (g["prev_K"], g["prev_k"], g["prev_covar"]), (g["x0_mean"], g["x0_covar"]), (g["C"], g["c"], g["cc"]))
    assert abs(mult - float(g["sa_mult"])) < 1e-9
    m = golden("misc")
    ec = oracle.expected_costs(m["C"], m["c"], m["cc"], m["mean"], m["covar"], 4)
    assert relerr(ec, m["expected"]) < TOL
    mean, cov = oracle.state_fit(m["X0"])
    assert relerr(mean, m["sd_mean"]) < TOL and relerr(cov, m["sd_covar"]) < TOL


def test_brentq_restatement_matches_live_scipy_on_random_functions():
    """Property test of the oracle's Brent restatement against the installed scipy (the third-party dependency of
    lqg.py:121): random monotone and non-monotone bracketed functions, random tolerances - identical roots AND
    identical sequences of evaluated abscissae."""
    from hypothesis import given, settings
    from hypothesis import strategies as st
    from scipy.optimize import brentq as scipy_brentq

    @settings(max_examples=150, deadline=None)
    @given(a=st.floats(-5, 5), b=st.floats(0.05, 3), c=st.floats(-2, 2), kind=st.integers(0, 3),
           rtol=st.sampled_from([1e-2, 1e-4, 1e-8, 4 * np.finfo(float).eps]))
    def check(a, b, c, kind, rtol):
        fns = [lambda x: b * (x - a), lambda x: np.tanh(b * (x - a)) + 1e-3 * c * (x - a),
               lambda x: np.expm1(b * (x - a) / 8), lambda x: (x - a) ** 3 + c * 1e-2 * (x - a)]
        fn = fns[kind]
        lo, hi = a - 7.3, a + 11.1
        if not np.sign(fn(lo)) * np.sign(fn(hi)) < 0:
            return
        seen = []

        def traced(x):
            seen.append(x)
            return fn(x)

        try:
            want = scipy_brentq(traced, lo, hi, xtol=2e-12, rtol=rtol, maxiter=100)
        except RuntimeError:  # no convergence in 100 iterations (a triple root): the restatement must say so too
            with pytest.raises(RuntimeError):
                oracle.brentq(fn, lo, hi, rtol=rtol)
            return
        got, trace = oracle.brentq(fn, lo, hi, rtol=rtol)
        assert got == want and trace == seen

    check()
