"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_golden.py``.

Steps
  1. run the reference's own hot-path unit tests (its hard-coded known answers,
     SURVEY.md section 8(c)) and refuse to write fixtures unless they pass;
  2. re-run the same seeded cases through the reference API and store inputs
     and full-precision outputs;
  3. add cases the reference's tests do not pin (ILQR.step/optimize with the
     scipy brentq iterate sequence, sklearn EM iterations with explicit init,
     warm start, step_adjust) from the live reference / sklearn / scipy.
"""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

REF = "/root/reference"
HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, REF)
sys.dont_write_bytecode = True


def run_reference_tests():
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    cmd = [
        sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider",
        "tests/traj_opt_test.py", "tests/dynamics_test.py", "tests/samples_test.py", "tests/utils_test.py",
    ]
    out = subprocess.run(cmd, cwd=REF, env=env, capture_output=True, text=True)
    tail = out.stdout.strip().splitlines()[-1]
    if out.returncode != 0:
        raise SystemExit("reference tests failed here:\n" + out.stdout[-2000:])
    return tail


def lqg_small():
    from donk.costs import loss_l2
    from donk.samples import StateDistribution
    from donk.traj_opt import lqg
    from tests.utils import random_lq_pol, random_spd, random_tvlg

    out = {}
    # tests/traj_opt_test.py:8-63
    T, dX, dU = 5, 3, 2
    rng = np.random.default_rng(0)
    x0 = StateDistribution(mean=rng.normal(size=(dX)), covar=random_spd((dX, dX), rng))
    dyn = random_tvlg(T, dX, dU, rng)
    pol = random_lq_pol(T, dX, dU, rng)
    traj = lqg.forward(dyn, pol, x0)
    out.update(fw_x0_mean=x0.mean, fw_x0_covar=x0.covar, fw_F=dyn.F, fw_f=dyn.f, fw_dyn_covar=dyn.covar,
               fw_K=pol.K, fw_k=pol.k, fw_pol_covar=pol.covar, fw_mean=traj.mean, fw_covar=traj.covar)
    # tests/traj_opt_test.py:65-91
    rng = np.random.default_rng(0)
    dyn = random_tvlg(T, dX, dU, rng)
    _, c, C = loss_l2(
        x=rng.normal(size=(T + 1, dX + dU)),
        t=rng.normal(size=(T + 1, dX + dU)),
        w=np.concatenate([rng.normal(size=(T + 1, dX)) > 0, 1e-2 * np.ones((T + 1, dU))], axis=1),
    )
    pol = lqg.backward(dyn, C, c)
    out.update(bw_F=dyn.F, bw_f=dyn.f, bw_C=C, bw_c=c, bw_K=pol.K, bw_k=pol.k, bw_covar=pol.covar,
               bw_inv_covar=pol.inv_covar)
    pol = lqg.backward(dyn, C, c, gamma=0.9)
    out.update(bwg_K=pol.K, bwg_k=pol.k, bwg_covar=pol.covar)
    # tests/traj_opt_test.py:93-131
    rng = np.random.default_rng(0)
    prev = random_lq_pol(T, 2, 3, rng)
    Ck, ck = lqg.extended_costs_kl(prev)
    out.update(ex_K=prev.K, ex_k=prev.k, ex_covar=prev.covar, ex_inv_covar=prev.inv_covar, ex_C=Ck, ex_c=ck)
    # tests/traj_opt_test.py:133-145
    rng = np.random.default_rng(0)
    X = rng.standard_normal((T, dX))
    pol = random_lq_pol(T, dX, dU, rng)
    prev = random_lq_pol(T, dX, dU, rng)
    kl = lqg.kl_divergence_action(X, pol, prev)
    assert abs(kl - 3.914744335914712) < 1e-12
    out.update(kl_X=X, kl_K=pol.K, kl_k=pol.k, kl_covar=pol.covar, kl_pK=prev.K, kl_pk=prev.k,
               kl_pcovar=prev.covar, kl_pinv=prev.inv_covar, kl_value=kl)
    np.savez_compressed(HERE / "lqg_small.npz", **out)


def fitlr_traj00():
    from donk.dynamics import to_transitions
    from donk.dynamics.linear_dynamics import fit_lr
    from donk.dynamics.prior import GMMPrior

    with np.load(f"{REF}/tests/data/traj_00.npz") as data:
        X = data["X"]
        U = data["U"][:, :-1]
    dyn = fit_lr(X, U, regularization=1e-6)  # tests/dynamics_test.py:21-90
    out = dict(X=X, U=U, F=dyn.F, f=dyn.f, covar=dyn.covar)
    prior = GMMPrior(8, random_state=0)  # tests/dynamics_test.py:102-178
    prior.update(to_transitions(X, U))
    g = prior.gmm
    dynp = fit_lr(X[:3], U[:3], prior, regularization=0)
    out.update(gmm_weights=g.weights_, gmm_means=g.means_, gmm_covariances=g.covariances_,
               gmm_prec_chol=g.precisions_cholesky_, gmm_N=prior.N, gmm_n_iter=g.n_iter_,
               gmm_lower_bound=g.lower_bound_, pF=dynp.F, pf=dynp.f, pcovar=dynp.covar)
    # prior.eval of one timestep (prior_gmm.py:26-41)
    xux = np.c_[X[:3, 7], U[:3, 7], X[:3, 8]]
    niw = prior.eval(xux)
    out.update(ev_xux=xux, ev_mu0=niw.mu0, ev_Phi=niw.Phi, ev_N_mean=niw.N_mean, ev_N_covar=niw.N_covar)
    # the k-means labels sklearn starts from (for the label-initialised EM oracle)
    from sklearn.cluster import KMeans
    from sklearn.utils import check_random_state

    labels = KMeans(n_clusters=8, n_init=1, random_state=check_random_state(0)).fit(to_transitions(X, U)).labels_
    out.update(kmeans_labels=labels)
    # dynamics prior with prior_weight != 1 and regularisation branch
    dynw = fit_lr(X[:6], U[:6], prior, regularization=1e-6, prior_weight=0.5)
    out.update(wF=dynw.F, wf=dynw.f, wcovar=dynw.covar)
    np.savez_compressed(HERE / "fitlr_traj00.npz", **out)


def gmm_em():
    from sklearn.mixture import GaussianMixture

    from donk_b200 import synthetic

    n, d, K = 1500, 10, 3
    X, w0, m0, p0 = synthetic.gmm_pool(n, d, K, seed=7)
    out = dict(X=X, w0=w0, m0=m0, p0=p0)
    import warnings

    for it in (1, 2, 3, 5):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            g = GaussianMixture(K, weights_init=w0, means_init=m0, precisions_init=p0, tol=0, max_iter=it).fit(X)
        out[f"it{it}_weights"] = g.weights_
        out[f"it{it}_means"] = g.means_
        out[f"it{it}_covariances"] = g.covariances_
        out[f"it{it}_prec_chol"] = g.precisions_cholesky_
        out[f"it{it}_lower_bound"] = g.lower_bound_
    # default stop rule + warm start on a second pool (prior_gmm.py:21-24)
    g = GaussianMixture(K, weights_init=w0, means_init=m0, precisions_init=p0).fit(X)
    out.update(d_weights=g.weights_, d_means=g.means_, d_covariances=g.covariances_, d_prec_chol=g.precisions_cholesky_,
               d_lower_bound=g.lower_bound_, d_n_iter=g.n_iter_, d_converged=g.converged_)
    X2 = X[::2] + 0.05
    g.warm_start = True
    g.fit(X2)
    out.update(X2=X2, w_weights=g.weights_, w_means=g.means_, w_covariances=g.covariances_,
               w_lower_bound=g.lower_bound_, w_n_iter=g.n_iter_, w_converged=g.converged_)
    # predict_proba on a few rows (GMMPrior.eval path)
    out.update(pp_rows=X2[:40], pp=g.predict_proba(X2[:40]))
    np.savez_compressed(HERE / "gmm_em.npz", **out)


def _ilqr_case(config, T, dX, dU, N, with_prior, etas):
    from donk.costs import QuadraticCosts
    from donk.dynamics import to_transitions
    from donk.dynamics.linear_dynamics import fit_lr
    from donk.dynamics.prior import GMMPrior
    from donk.policy import LinearGaussianPolicy
    from donk.samples import StateDistribution
    from donk.traj_opt import ILQR

    from donk_b200 import synthetic

    cond = synthetic.condition(config, 0, T, dX, dU, N)
    out = dict(X=cond.X, U=cond.U, prev_K=cond.prev_K, prev_k=cond.prev_k, prev_covar=cond.prev_covar,
               C=cond.C, c=cond.c, cc=cond.cc)
    prior = None
    if with_prior:
        pool = np.concatenate(
            [to_transitions(c.X, c.U) for c in (synthetic.condition(config, i, T, dX, dU, N) for i in range(1, 5))]
            + [to_transitions(cond.X, cond.U)]
        )
        prior = GMMPrior(4, random_state=0)
        prior.update(pool)
        g = prior.gmm
        out.update(gmm_weights=g.weights_, gmm_means=g.means_, gmm_covariances=g.covariances_,
                   gmm_prec_chol=g.precisions_cholesky_, gmm_N=prior.N)
    dyn = fit_lr(cond.X, cond.U, prior=prior)
    x0 = StateDistribution.fit(cond.X[:, 0])
    out.update(F=dyn.F, f=dyn.f, dyn_covar=dyn.covar, x0_mean=x0.mean, x0_covar=x0.covar)
    prev = LinearGaussianPolicy(cond.prev_K, cond.prev_k, cond.prev_covar)
    costs = QuadraticCosts(cond.C, cond.c, cond.cc)
    ilqr = ILQR(dyn, prev, costs, x0)
    out.update(prev_inv_covar=prev.inv_covar, C_kl=ilqr.C_kl, c_kl=ilqr.c_kl)
    out["etas"] = np.array(etas)
    for i, eta in enumerate(etas):
        r = ilqr.step(eta)
        out[f"s{i}_K"] = r.policy.K
        out[f"s{i}_k"] = r.policy.k
        out[f"s{i}_covar"] = r.policy.covar
        out[f"s{i}_inv_covar"] = r.policy.inv_covar
        out[f"s{i}_kl"] = r.kl_div
        out[f"s{i}_cost"] = r.expected_costs
        out[f"s{i}_mean"] = r.trajectory.mean
        if T <= 50:
            out[f"s{i}_tcovar"] = r.trajectory.covar
    # optimize, recording every eta the reference visits (lqg.py:90-123)
    visited = []
    orig = ilqr.step

    def rec(eta):
        visited.append(float(eta))
        return orig(eta)

    ilqr.step = rec
    for j, kl_step in enumerate((1.0, 0.05)):
        visited.clear()
        r = ilqr.optimize(kl_step)
        out[f"o{j}_kl_step"] = kl_step
        out[f"o{j}_visited"] = np.array(visited)
        out[f"o{j}_eta"] = r.eta
        out[f"o{j}_K"] = r.policy.K
        out[f"o{j}_k"] = r.policy.k
        out[f"o{j}_covar"] = r.policy.covar
        out[f"o{j}_kl"] = r.kl_div
        out[f"o{j}_cost"] = r.expected_costs
    return out, (dyn, prev, costs, x0)


def ilqr_cases():
    from donk.traj_opt import lqg

    out, (dyn, prev, costs, x0) = _ilqr_case(1, 50, 6, 2, 20, False, [1e-6, 1e-3, 1.0, 37.5, 1e4, 1e16])
    # step_adjust (lqg.py:309-350) between two consecutive "iterations" of the same condition
    from donk.policy import LinearGaussianPolicy

    new_pol = LinearGaussianPolicy(out["o0_K"], out["o0_k"], out["o0_covar"])
    dyn2 = type(dyn)(dyn.F * 0.98, dyn.f + 0.01, dyn.covar * 1.1)
    out["sa_F2"], out["sa_f2"], out["sa_cov2"] = dyn2.F, dyn2.f, dyn2.covar
    out["sa_mult"] = lqg.step_adjust(1.0, dyn2, new_pol, x0, costs, dyn, prev, x0, costs)
    np.savez_compressed(HERE / "ilqr_c1.npz", **out)

    out, _ = _ilqr_case(2, 100, 14, 7, 50, True, [1e-2, 1.0, 1e2])
    np.savez_compressed(HERE / "ilqr_c2.npz", **out)


def brent_traces():
    """scipy.optimize.brentq iterate sequences on scalar test functions."""
    from scipy import optimize

    out = {}
    funs = {
        "cubic": (lambda x: (x - 1.3) ** 3 + 0.2 * (x - 1.3), -13.8, 36.8),
        "tanh": (lambda x: np.tanh(0.3 * (x - 5.0)) + 0.1, -13.8, 36.8),
        "exp": (lambda x: np.exp(-0.2 * x) - 0.05, -13.8, 36.8),
        "kink": (lambda x: (abs(x + 3.0) ** 0.5) * np.sign(x + 3.0) + 0.01 * x, -13.8, 36.8),
    }
    for name, (fn, a, b) in funs.items():
        for rtol in (1e-2, 1e-6):
            seq = []

            def rec(x, fn=fn, seq=seq):
                seq.append(x)
                return fn(x)

            root = optimize.brentq(rec, a, b, rtol=rtol)
            out[f"{name}_{rtol:g}_root"] = root
            out[f"{name}_{rtol:g}_seq"] = np.array(seq)
    np.savez_compressed(HERE / "brent.npz", **out)


def misc():
    from donk.costs import QuadraticCosts
    from donk.samples import StateDistribution, TrajectoryDistribution
    from tests.utils import random_spd

    rng = np.random.default_rng(3)
    T, dX, dU = 6, 4, 2
    p = dX + dU
    X0 = rng.standard_normal((9, dX))
    sd = StateDistribution.fit(X0)
    C = random_spd((T + 1, p, p), rng)
    c = rng.standard_normal((T + 1, p))
    cc = rng.standard_normal(T + 1)
    mean = rng.standard_normal((T + 1, p))
    mean[-1, dX:] = 0
    covar = random_spd((T + 1, p, p), rng)
    ec = QuadraticCosts(C, c, cc).expected_costs(TrajectoryDistribution(mean, covar, dX, dU))
    np.savez_compressed(HERE / "misc.npz", X0=X0, sd_mean=sd.mean, sd_covar=sd.covar, C=C, c=c, cc=cc, mean=mean,
                        covar=covar, expected=ec)


if __name__ == "__main__":
    import numpy, scipy, sklearn  # noqa: E401

    tail = run_reference_tests()
    os.chdir(REF)  # the reference opens its fixtures by relative path
    lqg_small()
    fitlr_traj00()
    gmm_em()
    ilqr_cases()
    brent_traces()
    misc()
    manifest = dict(
        reference_tests=tail,
        numpy=numpy.__version__, scipy=scipy.__version__, sklearn=sklearn.__version__,
        files=sorted(p.name for p in HERE.glob("*.npz")),
    )
    (HERE / "MANIFEST.json").write_text(json.dumps(manifest, indent=1) + "\n")
    print(manifest)
