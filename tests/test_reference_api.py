"""The drop-in Python API (same names / signatures / errors as the reference) on top of the kernels.

Mirrors the reference's own tests: tests/traj_opt_test.py, tests/dynamics_test.py, tests/samples_test.py,
tests/utils_test.py, tests/costs_test.py:258-301.
"""
import pickle

import numpy as np
import pytest

from donk_testutil import load_golden, relerr, to_np


def test_lqg_module_functions(backend):
    from donk_b200.dynamics import LinearDynamics
    from donk_b200.policy import LinearGaussianPolicy
    from donk_b200.samples import StateDistribution
    from donk_b200.traj_opt import lqg

    g = load_golden("lqg_small")
    dyn = LinearDynamics(g["fw_F"], g["fw_f"], g["fw_dyn_covar"])
    pol = LinearGaussianPolicy(g["fw_K"], g["fw_k"], g["fw_pol_covar"])
    traj = lqg.forward(dyn, pol, StateDistribution(g["fw_x0_mean"], g["fw_x0_covar"]))
    assert traj.mean.shape == (6, 5) and traj.covar.shape == (6, 5, 5)
    assert relerr(traj.mean, g["fw_mean"]) < 1e-9 and relerr(traj.covar, g["fw_covar"]) < 1e-9
    assert np.array_equal(traj.covar, np.swapaxes(traj.covar, 1, 2))  # tests/traj_opt_test.py:63
    assert np.array_equal(traj.X_mean, traj.mean[:, :3]) and np.array_equal(traj.U_covar, traj.covar[:-1, 3:, 3:])
    pol = lqg.backward(LinearDynamics(g["bw_F"], g["bw_f"], g["fw_dyn_covar"]), g["bw_C"], g["bw_c"])
    assert relerr(pol.K, g["bw_K"]) < 1e-9 and relerr(pol.inv_covar, g["bw_inv_covar"]) < 1e-9
    prev = LinearGaussianPolicy(g["ex_K"], g["ex_k"], g["ex_covar"])
    C, c = lqg.extended_costs_kl(prev)  # lazy inv_covar from covar (tvlg.py:88-99)
    assert relerr(C, g["ex_C"]) < 1e-9 and relerr(c, g["ex_c"]) < 1e-9
    kl = lqg.kl_divergence_action(g["kl_X"], LinearGaussianPolicy(g["kl_K"], g["kl_k"], g["kl_covar"]),
                                  LinearGaussianPolicy(g["kl_pK"], g["kl_pk"], g["kl_pcovar"]))
    assert abs(kl - 3.914744335914712) < 1e-12
    with pytest.raises(np.linalg.LinAlgError):  # Quu not positive definite (lqg.py:165)
        lqg.backward(LinearDynamics(g["bw_F"], g["bw_f"], g["fw_dyn_covar"]), -g["bw_C"], g["bw_c"])


def test_ilqr_class(backend):
    from donk_b200.costs import QuadraticCosts
    from donk_b200.dynamics import LinearDynamics
    from donk_b200.policy import LinearGaussianPolicy
    from donk_b200.samples import StateDistribution
    from donk_b200.traj_opt import ILQR
    from donk_b200.traj_opt import lqg

    g = load_golden("ilqr_c1")
    dyn = LinearDynamics(g["F"], g["f"], g["dyn_covar"])
    prev = LinearGaussianPolicy(g["prev_K"], g["prev_k"], g["prev_covar"])
    costs = QuadraticCosts(g["C"], g["c"], g["cc"])
    x0 = StateDistribution(g["x0_mean"], g["x0_covar"])
    ilqr = ILQR(dyn, prev, costs, x0)
    assert relerr(ilqr.C_kl, g["C_kl"]) < 1e-9
    r = ilqr.step(float(g["etas"][2]))
    assert relerr(r.policy.K, g["s2_K"]) < 1e-9 and abs(r.kl_div - g["s2_kl"]) < 1e-9
    assert relerr(r.trajectory.covar, g["s2_tcovar"]) < 1e-9
    assert abs(r.expected_costs - g["s2_cost"]) < 1e-9 * abs(g["s2_cost"])
    assert relerr(r.policy.chol_covar, np.linalg.cholesky(g["s2_covar"])) < 1e-9
    r = ilqr.optimize(1.0)
    assert abs(r.eta - g["o0_eta"]) < 1e-9 * g["o0_eta"] and relerr(r.policy.K, g["o0_K"]) < 1e-8
    surf = ilqr.sample_surface(1e-3, 1e4, N=8)
    assert len(surf) == 8 and surf[0].eta == pytest.approx(1e-3)
    assert abs(surf[0].kl_div - g["s1_kl"]) < 1e-9 * abs(g["s1_kl"])
    with pytest.raises(ValueError, match="max_eta eta to low"):
        ilqr.optimize(1e-30, max_eta=1e-3)
    pickle.dumps(r)  # results stay plain numpy holders (datalogging / shelve compatibility)
    # step_adjust (lqg.py:309-350)
    dyn2 = LinearDynamics(g["sa_F2"], g["sa_f2"], g["sa_cov2"])
    new_pol = LinearGaussianPolicy(g["o0_K"], g["o0_k"], g["o0_covar"])
    mult = lqg.step_adjust(1.0, dyn2, new_pol, x0, costs, dyn, prev, x0, costs)
    assert abs(mult - float(g["sa_mult"])) < 1e-9


def test_fit_lr_api(backend):
    from donk_b200.dynamics import LinearDynamics
    from donk_b200.dynamics.linear_dynamics import fit_lr

    g = load_golden("fitlr_traj00")
    dyn = fit_lr(g["X"], g["U"], regularization=1e-6)
    assert dyn.F.shape == (44, 15, 17) and dyn.f.shape == (44, 15) and dyn.covar.shape == (44, 15, 15)
    assert (dyn.T, dyn.dX, dyn.dU) == (44, 15, 2)
    assert relerr(dyn.F, g["F"]) < 1e-8
    assert np.array_equal(dyn.covar, np.swapaxes(dyn.covar, 1, 2))  # tests/dynamics_test.py:84
    for t in range(44):
        assert np.all(np.linalg.eigvalsh(dyn.covar[t]) >= -1e-12)
    with pytest.raises(ValueError):  # tests/dynamics_test.py:92-100
        fit_lr(np.empty((1, 5, 3)), np.empty((1, 4, 2)), regularization=1e-6)
    d = LinearDynamics(F=np.array([[[1, 0, 1], [0, 1, 1]]]), f=np.array([[0, 1]]),
                       covar=np.array([[[0.1, 0.1], [0.1, 0.1]]]))
    assert np.array_equal(d.predict(x=[1, 1], u=[1], t=0, noise=None), [2, 3])  # tests/dynamics_test.py:10-19
    assert str(LinearDynamics(np.empty((4, 3, 5)), np.empty((4, 3)), np.empty((4, 3, 3)))) == \
        "LinearDynamics[T=4, dX=3, dU=2]"


def test_value_types_and_helpers(backend):
    from donk_b200.costs import QuadraticCosts
    from donk_b200.dynamics.prior import NormalInverseWishart
    from donk_b200.samples import StateDistribution, TrajectoryDistribution, TransitionPool
    from donk_b200.utils import regularize, symmetrize, trace_of_product
    from donk_b200.utils.batched import batched_cholesky, batched_inv_spd

    rng = np.random.default_rng(0)
    X, U = rng.standard_normal((3, 11, 5)), rng.standard_normal((3, 10, 3))
    pool = TransitionPool()
    pool.add(X, U)  # tests/samples_test.py:9-31
    assert np.array_equal(pool.get_transitions(),
                          np.c_[X[:, :-1].reshape(-1, 5), U.reshape(-1, 3), X[:, 1:].reshape(-1, 5)])
    assert np.array_equal(pool.get_transitions(N=5),
                          np.c_[X[2, -6:-1].reshape(-1, 5), U[2, -5:].reshape(-1, 3), X[2, -5:].reshape(-1, 5)])
    assert not pool.get_transitions().flags.writeable
    capped = TransitionPool(capacity=7)
    capped.add(X, U)
    assert capped.get_transitions().shape == (7, 13)
    A = rng.uniform(size=(4, 3, 3))
    B = A.copy()
    assert symmetrize(B) is B and np.array_equal(B, np.swapaxes(B, 1, 2))  # tests/utils_test.py:30-47
    assert regularize(B, 0.5) is B
    assert np.allclose(trace_of_product(A, B), np.trace(A @ B, axis1=1, axis2=2))
    spd = B + 3 * np.eye(3)
    assert relerr(batched_cholesky(spd), np.linalg.cholesky(spd)) < 1e-12
    assert relerr(batched_inv_spd(np.linalg.cholesky(spd)), np.linalg.inv(spd)) < 1e-10
    m = load_golden("misc")
    sd = StateDistribution.fit(m["X0"])
    assert relerr(sd.mean, m["sd_mean"]) < 1e-12 and relerr(sd.covar, m["sd_covar"]) < 1e-12
    ec = QuadraticCosts(m["C"], m["c"], m["cc"]).expected_costs(TrajectoryDistribution(m["mean"], m["covar"], 4, 2))
    assert relerr(ec, m["expected"]) < 1e-12
    prior = NormalInverseWishart.non_informative_prior(1).posterior(np.array([1]), np.array([[0.5]]), 1)
    assert np.allclose(prior.map_mean(), [1]) and np.allclose(prior.map_covar(), [[0.5]])  # dynamics_test.py:222-232


def test_traj_opt_algorithm_two_iterations(backend):
    """TrajOptAlgorithm.iteration (traj_opt.py:35-96) vs the same sequence through the oracle."""
    import oracle
    from donk_b200 import synthetic
    from donk_b200.costs import QuadraticCosts
    from donk_b200.policy import LinearGaussianPolicy
    from donk_b200.traj_opt import TrajOptAlgorithm
    from oracle.lqg import ILQRProblem

    T, dX, dU, N = 10, 4, 2, 10
    c0, c1 = synthetic.condition(8, 0, T, dX, dU, N), synthetic.condition(8, 1, T, dX, dU, N)
    costs = QuadraticCosts(c0.C, c0.c, c0.cc)
    algo = TrajOptAlgorithm(kl_step=1.0)
    algo.iteration(c0.X, c0.U, costs, prev_pol=LinearGaussianPolicy(c0.prev_K, c0.prev_k, c0.prev_covar))
    pol1 = algo.pol
    algo.iteration(c1.X, c1.U, costs)
    # oracle, same order of operations
    F0, f0, S0 = oracle.fit_lr(c0.X, c0.U)
    x00 = oracle.state_fit(c0.X[:, 0])
    p0 = ILQRProblem(F0, f0, S0, c0.C, c0.c, c0.cc, c0.prev_K, c0.prev_k, c0.prev_covar,
                     np.linalg.inv(c0.prev_covar), *x00)
    r0, _ = oracle.ilqr_optimize(p0, 1.0)
    assert relerr(pol1.K, r0.K) < 1e-8
    F1, f1, S1 = oracle.fit_lr(c1.X, c1.U)
    x01 = oracle.state_fit(c1.X[:, 0])
    p1 = ILQRProblem(F1, f1, S1, c0.C, c0.c, c0.cc, r0.K, r0.k, r0.covar, r0.inv_covar, *x01)
    r1, _ = oracle.ilqr_optimize(p1, 1.0)
    assert relerr(algo.pol.K, r1.K) < 1e-7
    mult, *_ = oracle.step_adjust(1.0, (F1, f1, S1), (r1.K, r1.k, r1.covar), x01, (c0.C, c0.c, c0.cc),
                                  (F0, f0, S0), (r0.K, r0.k, r0.covar), x00, (c0.C, c0.c, c0.cc))
    assert abs(algo.kl_step_mult - mult) < 1e-6


def test_batched_gps_iteration_matches_per_condition_api(backend):
    """BatchedTrajOpt (full GPS inner iteration, BASELINE config 5 shape) == TrajOptAlgorithm per condition."""
    import torch

    from donk_b200 import synthetic
    from donk_b200.batched import BatchedTrajOpt
    from donk_b200.costs import QuadraticCosts
    from donk_b200.policy import LinearGaussianPolicy
    from donk_b200.traj_opt import TrajOptAlgorithm

    T, dX, dU, N, B = 8, 6, 2, 12, 3
    it0 = [synthetic.condition(30, i, T, dX, dU, N) for i in range(B)]
    it1 = [synthetic.condition(31, i, T, dX, dU, N) for i in range(B)]
    st = lambda xs: np.stack(xs)  # noqa: E731
    algo = BatchedTrajOpt(kl_step=0.7)
    C, c, cc = st([x.C for x in it0]), st([x.c for x in it0]), st([x.cc for x in it0])
    algo.iteration(st([x.X for x in it0]), st([x.U for x in it0]), C, c, cc, st([x.prev_K for x in it0]),
                   st([x.prev_k for x in it0]), st([x.prev_covar for x in it0]))
    algo.iteration(st([x.X for x in it1]), st([x.U for x in it1]), C, c, cc)
    for b in range(B):
        single = TrajOptAlgorithm(kl_step=0.7)
        costs = QuadraticCosts(it0[b].C, it0[b].c, it0[b].cc)
        single.iteration(it0[b].X, it0[b].U, costs,
                         prev_pol=LinearGaussianPolicy(it0[b].prev_K, it0[b].prev_k, it0[b].prev_covar))
        single.iteration(it1[b].X, it1[b].U, costs)
        assert relerr(algo.pol[0][b], single.pol.K) < 1e-9 and relerr(algo.pol[2][b], single.pol.covar) < 1e-9
        assert abs(float(algo.kl_step_mult[b]) - single.kl_step_mult) < 1e-9


def test_gps_iteration_from_host_equals_device_iteration(backend):
    """BatchedTrajOpt.iteration_from_host (chunked upload overlapped with the dynamics fit, cost buffers in turn) gives
    the iteration() results bit for bit over three iterations (the third reads the costs of the second in step_adjust)."""
    import torch

    from donk_b200 import synthetic
    from donk_b200.batched import BatchedTrajOpt

    T, dX, dU, N, B = 6, 6, 2, 10, 5
    its = [[synthetic.condition(40 + k, i, T, dX, dU, N) for i in range(B)] for k in range(3)]
    st = lambda xs: np.stack(xs)  # noqa: E731
    a, b = BatchedTrajOpt(kl_step=0.7), BatchedTrajOpt(kl_step=0.7)
    for k, it in enumerate(its):
        X, U = st([x.X for x in it]), st([x.U for x in it])
        C, c, cc = (st([getattr(x, n) for x in it]) * (1.0 + 0.1 * k) for n in ("C", "c", "cc"))
        if k == 0:
            pol = [st([getattr(x, n) for x in it]) for n in ("prev_K", "prev_k", "prev_covar")]
            a.iteration(X, U, C, c, cc, *pol)
            b.iteration(X, U, C, c, cc, *pol)
            with pytest.raises(Exception):
                BatchedTrajOpt(kl_step=0.7).iteration_from_host(X, U, C, c, cc)
            continue
        a.iteration(X, U, C, c, cc)
        host = [torch.as_tensor(t) for t in (X, U, C, c, cc)]
        if backend.device.type == "cuda":
            host = [t.pin_memory() for t in host]
        out = [torch.empty(t.shape, dtype=torch.float64) for t in a.pol]
        b.iteration_from_host(*host, chunks=3, out_policy=out)
        if backend.device.type == "cuda":
            torch.cuda.synchronize()
        for ta, tb, th in zip(a.pol, b.pol, out):
            assert torch.equal(ta, tb) and np.array_equal(to_np(ta), th.numpy())
        assert torch.equal(a.kl_step_mult, b.kl_step_mult)


def test_device_transition_pool_matches_reference_semantics(backend):
    """batched.DeviceTransitionPool against the host TransitionPool (donk/samples/pool.py:6-56): growing pool, FIFO
    capacity with wrap-around, a batch larger than the capacity, newest-N windows."""
    from donk_b200.batched import DeviceTransitionPool
    from donk_b200.samples import TransitionPool

    rng = np.random.default_rng(0)
    dX, dU, T = 3, 2, 4
    for cap in (None, 23, 8):
        host, dev = TransitionPool(cap), DeviceTransitionPool(cap)
        for N in (2, 3, 1, 4, 2):
            X, U = rng.standard_normal((N, T + 1, dX)), rng.standard_normal((N, T, dU))
            host.add(X, U)
            dev.add(X, U)
            assert len(dev) == host.get_transitions().shape[0]
            assert np.array_equal(to_np(dev.get_transitions()), host.get_transitions())
            for n in (1, 5, 100):
                assert np.array_equal(to_np(dev.get_transitions(n)), host.get_transitions(n))
    with pytest.raises(AssertionError):
        dev.add(rng.standard_normal((2, T + 1, dX + 1)), rng.standard_normal((2, T, dU)))


def test_log_prob_known_answer(backend):
    """tests/dynamics_test.py:180-204: fit_lr under GMMPrior(8, random_state=0) on the training roll-outs of
    tests/data/traj_00.npz, LinearDynamics.log_prob of the three held-out roll-outs (the reference's literal values)."""
    from donk_b200.dynamics import to_transitions
    from donk_b200.dynamics.linear_dynamics import fit_lr
    from donk_b200.dynamics.prior import GMMPrior

    g = load_golden("fitlr_traj00")
    X, U = g["X"], g["U"]
    N, _, dX = X.shape
    _, T, dU = U.shape
    X_train, U_train, X_test, U_test = X[:-3], U[:-3], X[-3:], U[-3:]
    prior = GMMPrior(8, random_state=0, init="sklearn")  # the literal values pin scikit-learn's k-means stream
    prior.update(to_transitions(X_train, U_train).reshape(-1, dX + dU + dX))
    dyn = fit_lr(X_train, U_train, prior, regularization=0)
    log_prob = dyn.log_prob(X_test, U_test)
    assert log_prob.shape == (3, T)
    np.testing.assert_allclose(np.mean(log_prob, axis=-1), [-2139.578291, -2076.051295, -911263.960917])
