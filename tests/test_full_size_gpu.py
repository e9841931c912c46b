"""BASELINE.json's full sizes on the GPU, checked through size-independent properties plus oracle spot checks.

The oracle cannot run 65 536 solves or 10 M EM rows in test time, so at full size the CUDA path is pinned by:
replication invariance (a condition copied to another batch position gives bit-identical results), packing
invariance (E = 16 candidates per launch vs one at a time: different CTA shapes, same numbers), monotonicity of the
dual (KL falls as eta grows), idempotence of `optimize` (a step at the returned eta reproduces the result), linearity
of the EM statistics over row shards, monotonicity of the EM lower bound — and by the oracle on a few units drawn
from the full-size batch.
"""
import numpy as np
import pytest
import torch

import oracle
from donk_testutil import relerr
from oracle.lqg import ILQRProblem

pytestmark = pytest.mark.gpu


@pytest.fixture
def backend():
    """The CUDA library only (the emulated backend of conftest.py is for small cases)."""
    from donk_b200 import _lib

    _lib.install_test_backend(None)
    return _lib.native()


def _ilqr_from_rollouts(ro, T, dX, dU, prior=None):
    from donk_b200 import batched, synthetic

    dev = ro["X"].device
    B, p = ro["X"].shape[0], dX + dU
    F, f, S, info = batched.fit_lr(ro["X"], ro["U"], prior)
    assert int(info.sum()) == 0
    C, c, cc = synthetic.l2_costs(T, dX, dU)
    x0m, x0S = batched.state_fit_from_rollouts(ro["X"])
    il = batched.BatchedILQR(F, f, S, ro["prev_K"], ro["prev_k"],
                             torch.as_tensor(C, device=dev).expand(B, T + 1, p, p).contiguous(),
                             torch.as_tensor(c, device=dev).expand(B, T + 1, p).contiguous(),
                             torch.as_tensor(cc, device=dev).expand(B, T + 1).contiguous(), x0m, x0S,
                             prev_covar=ro["prev_covar"])
    return il, (F, f, S), (C, c, cc), (x0m, x0S)


def _oracle_problem(b, ro, dyn, costs, x0):
    F, f, S = (t[b].cpu().numpy() for t in dyn)
    pc = ro["prev_covar"][b].cpu().numpy()
    return ILQRProblem(F, f, S, *costs, ro["prev_K"][b].cpu().numpy(), ro["prev_k"][b].cpu().numpy(), pc,
                       np.linalg.inv(pc), x0[0][b].cpu().numpy(), x0[1][b].cpu().numpy())


def test_config3_full_size_step_and_optimize(backend):
    """BASELINE configs[2]: 4096 conditions x 16 eta, T=100, dX=20, dU=6."""
    from donk_b200 import synthetic

    dev = backend.device
    B, E, T, dX, dU, N = 4096, 16, 100, 20, 6, 64
    ro = synthetic.rollout_batch_torch(B, T, dX, dU, N, seed=3, device=dev)
    for key in ro:  # replication: the last 8 conditions are copies of the first 8
        ro[key][-8:] = ro[key][:8]
    il, dyn, costs, x0 = _ilqr_from_rollouts(ro, T, dX, dU)
    for t in dyn:  # the fits of the copies are bit-identical too
        assert torch.equal(t[-8:], t[:8])
    eta = torch.as_tensor(np.logspace(-2, 2, E), device=dev).expand(B, E).contiguous()
    r = il.step(eta)
    assert int(r.info.sum()) == 0
    kl, cost, K0 = r.kl_div.clone(), r.expected_costs.clone(), r.K[:8].clone()
    assert torch.isfinite(kl).all() and torch.isfinite(cost).all()
    # replication invariance
    assert torch.equal(kl[-8:], kl[:8]) and torch.equal(cost[-8:], cost[:8]) and torch.equal(r.K[-8:], K0)
    # the dual is monotone: a larger eta weights the KL term more, so KL(eta) falls
    assert bool((kl[:, 1:] <= kl[:, :-1] * (1 + 1e-9) + 1e-12).all())
    # packing invariance: one candidate per launch (one-warp CTAs) gives the same numbers as 16 per launch
    for e in (0, 7, 15):
        r1 = il.step(eta[:, e:e + 1].contiguous())
        assert relerr(r1.kl_div[:, 0], kl[:, e]) < 1e-12 and relerr(r1.expected_costs[:, 0], cost[:, e]) < 1e-12
    # oracle on three conditions drawn from the batch
    for b in (0, 1777, 4095):
        prob = _oracle_problem(b, ro, dyn, costs, x0)
        r = il.step(eta)
        for e in (0, 9, 15):
            ref = oracle.ilqr_step(prob, float(eta[b, e]))
            assert relerr(r.K[b, e], ref.K) < 1e-8 and relerr(r.covar[b, e], ref.covar) < 1e-8
            assert abs(float(kl[b, e]) - ref.kl_div) < 1e-8 * max(1.0, abs(ref.kl_div))
            assert abs(float(cost[b, e]) - ref.expected_costs) < 1e-8 * abs(ref.expected_costs)
    # ILQR.optimize for all 4096 conditions; idempotence: a step at the returned eta reproduces the result
    res, status, n_steps = il.optimize(1.0, raise_on_failure=False)
    assert int((status != 0).sum()) == 0 and 8 <= n_steps <= 40
    again = il.step(res.eta)
    assert torch.equal(again.kl_div, res.kl_div) and torch.equal(again.K, res.K)
    assert bool(((res.kl_div[:, 0] - 1.0).abs() < 0.1).all())  # Brent stops at rtol 1e-2 in log eta
    assert torch.equal(res.eta[-8:], res.eta[:8])
    ref, _ = oracle.ilqr_optimize(_oracle_problem(5, ro, dyn, costs, x0), 1.0)
    assert abs(float(res.eta[5, 0]) - ref.eta) < 1e-9 * ref.eta and relerr(res.K[5, 0], ref.K) < 1e-8


def test_config2_fit_with_prior_batch(backend):
    """BASELINE configs[1]: fit_lr with GMMPrior(4), T=100, dX=14, dU=7, N=50, for 256 conditions at once."""
    from donk_b200 import batched, synthetic
    from oracle.gmm import GMMParams, gmm_eval_prior

    dev = backend.device
    B, T, dX, dU, N, K = 256, 100, 14, 7, 50, 4
    ro = synthetic.rollout_batch_torch(B, T, dX, dU, N, seed=11, device=dev)
    ro["X"][-4:], ro["U"][-4:] = ro["X"][:4], ro["U"][:4]
    pool = batched.to_transitions(ro["X"][:8].reshape(-1, T + 1, dX), ro["U"][:8].reshape(-1, T, dU))
    d = 2 * dX + dU
    rng = np.random.default_rng(0)
    idx = rng.choice(pool.shape[0], K, replace=False)
    gm = batched.DeviceGMM(K).set_params(np.full(K, 1 / K), pool[idx].cpu().numpy(),
                                         precisions=np.broadcast_to(np.eye(d), (K, d, d)).copy())
    gm.fit(pool)
    assert gm.converged and all(b >= a - 1e-9 for a, b in zip(gm.lower_bounds, gm.lower_bounds[1:]))
    F, f, S, info = batched.fit_lr(ro["X"], ro["U"], gm)
    assert int(info.sum()) == 0
    assert torch.equal(F[-4:], F[:4]) and torch.equal(S[-4:], S[:4])
    assert torch.equal(S, S.transpose(-1, -2))       # exactly symmetric (tests/dynamics_test.py:84)
    torch.linalg.cholesky(S)                         # and positive definite
    prm = GMMParams(gm.weights.cpu().numpy(), gm.means.cpu().numpy(), gm.covariances.cpu().numpy(),
                    gm.precisions_cholesky.cpu().numpy())
    for b in (0, 200):
        Fo, fo, co = oracle.fit_lr(ro["X"][b].cpu().numpy(), ro["U"][b].cpu().numpy(),
                                   lambda xux: gmm_eval_prior(prm, gm.N, xux))
        assert relerr(F[b], Fo) < 1e-8 and relerr(f[b], fo) < 1e-8 and relerr(S[b], co) < 1e-8


@pytest.mark.parametrize("dims,N,K,B,T", [((20, 6), 64, 16, 24, 100), ((14, 7), 80, 8, 16, 60), ((16, 8), 40, 8, 20, 50)])
def test_fit_prior_weights_from_the_transition_view(backend, monkeypatch, dims, N, K, B, T):
    """fit_lr with a many-cluster prior at sizes where the cluster weights of all fits come from ONE tensor-path E pass
    over the transition view of the roll-outs (>= 32 768 sample rows; warp kernels with K >= 8 or N > 64, and the
    cooperative kernel): every fit of two conditions against the oracle, and the in-kernel form of the prior (forced)
    against the same numbers where it exists (N <= 64)."""
    from donk_b200 import batched, synthetic
    from oracle.gmm import GMMParams, gmm_eval_prior

    dev = backend.device
    dX, dU = dims
    d = 2 * dX + dU
    ro = synthetic.rollout_batch_torch(B, T, dX, dU, N, seed=13, device=dev)
    assert B * N * T >= 32768
    pool = batched.to_transitions(ro["X"][:4].reshape(-1, T + 1, dX), ro["U"][:4].reshape(-1, T, dU))
    rng = np.random.default_rng(1)
    idx = rng.choice(pool.shape[0], K, replace=False)
    gm = batched.DeviceGMM(K, max_iter=8).set_params(np.full(K, 1 / K), pool[idx].cpu().numpy(),
                                                    precisions=np.broadcast_to(np.eye(d), (K, d, d)).copy())
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gm.fit(pool)
    F, f, S, info = batched.fit_lr(ro["X"], ro["U"], gm)
    assert int(info.sum()) == 0 and torch.equal(S, S.transpose(-1, -2))
    prm = GMMParams(gm.weights.cpu().numpy(), gm.means.cpu().numpy(), gm.covariances.cpu().numpy(),
                    gm.precisions_cholesky.cpu().numpy())
    for b in (0, B - 1):
        Fo, fo, co = oracle.fit_lr(ro["X"][b].cpu().numpy(), ro["U"][b].cpu().numpy(),
                                   lambda xux: gmm_eval_prior(prm, gm.N, xux))
        assert relerr(F[b], Fo) < 1e-8 and relerr(f[b], fo) < 1e-8 and relerr(S[b], co) < 1e-8
    if N <= 64 and dims != (16, 8):
        monkeypatch.setenv("DONK_FIT_PRIOR_INKERNEL", "1")
        F2, f2, S2, info2 = batched.fit_lr(ro["X"], ro["U"], gm)
        assert int(info2.sum()) == 0
        assert relerr(F2, F.cpu().numpy()) < 1e-9 and relerr(S2, S.cpu().numpy()) < 1e-9


def test_config4_em_full_size(backend):
    """BASELINE configs[3]: EM on a 10 M-row pool, d=72, K=16."""
    from donk_b200 import _lib, batched, synthetic
    from oracle.gmm import gmm_e_step, gmm_m_step, params_from_precisions

    dev = backend.device
    n, d, K = 10_000_000, 72, 16
    X, w0, m0, p0 = synthetic.gmm_pool_torch(n, d, K, seed=4000, device=dev)
    gm = batched.DeviceGMM(K, tol=0.0, max_iter=3).set_params(w0, m0, precisions=p0)
    nat = _lib.native()
    o = dict(dtype=torch.float64, device=dev)
    ws = torch.empty(nat.size("gmm_workspace_bytes", n, d, K) // 8 + 1, **o)
    size = nat.size("gmm_stats_size", d, K)
    full, lo, hi = (torch.empty(size, **o) for _ in range(3))
    shift = torch.empty(K, d, **o)

    def stats_of(rows, out):
        nat.call("gmm_em_step_f64", rows.shape[0], d, K, rows, gm.weights, gm.means, gm.precisions_cholesky, None, out,
                 shift, ws, ws.numel() * 8)

    # linearity over row shards (what the multi-GPU all-reduce relies on): stats(all) = stats(first half) + stats(rest)
    stats_of(X, full)
    stats_of(X[:n // 2], lo)
    stats_of(X[n // 2:], hi)
    assert relerr(lo + hi, full) < 1e-11
    assert abs(float(full[:K].sum()) - n) < 1e-6 * n  # responsibilities sum to one per row
    # 200 k-row prefix against the oracle's E and M steps
    m = 200_000
    prm = params_from_precisions(w0, m0, p0)
    lb_ref, log_resp = gmm_e_step(X[:m].cpu().numpy(), prm)
    ref = gmm_m_step(X[:m].cpu().numpy(), np.exp(log_resp))
    g2 = batched.DeviceGMM(K, tol=0.0, max_iter=1).set_params(w0, m0, precisions=p0)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g2.fit(X[:m])
        gm.fit(X)
    assert abs(g2.lower_bound - lb_ref) < 1e-9 * abs(lb_ref)
    assert relerr(g2.means, ref.means) < 1e-8 and relerr(g2.covariances, ref.covariances) < 1e-8
    assert relerr(g2.weights, ref.weights) < 1e-8
    # full size: the EM lower bound never falls, weights stay a distribution, covariances stay symmetric PD
    lbs = gm.lower_bounds
    assert len(lbs) == 3 and all(b >= a - 1e-9 for a, b in zip(lbs, lbs[1:]))
    assert abs(float(gm.weights.sum()) - 1.0) < 1e-12 and gm.N == n
    assert relerr(gm.covariances, gm.covariances.transpose(-1, -2)) < 1e-14
    torch.linalg.cholesky(gm.covariances)


def test_config5_dims_long_horizon(backend):
    """BASELINE configs[4] dimensions (T=500, dX=64, dU=16, N=128) on a few conditions: the CTA-per-problem iLQR kernel
    over the full horizon against the oracle (500 sequential Riccati steps at p = 80 is where rounding accumulates), the
    full `optimize` of one condition, and the dynamics fit."""
    from donk_b200 import synthetic

    dev = backend.device
    B, T, dX, dU, N = 6, 500, 64, 16, 128
    ro = synthetic.rollout_batch_torch(B, T, dX, dU, N, seed=5, device=dev)
    for key in ro:
        ro[key][-1:] = ro[key][:1]
    il, dyn, costs, x0 = _ilqr_from_rollouts(ro, T, dX, dU)
    F, f, S = dyn
    assert torch.equal(S, S.transpose(-1, -2)) and torch.equal(F[-1], F[0])
    eta = torch.as_tensor([[0.1, 10.0]], device=dev).expand(B, 2).contiguous()
    r = il.step(eta)
    assert int(r.info.sum()) == 0 and torch.isfinite(r.kl_div).all()
    assert torch.equal(r.kl_div[-1], r.kl_div[0]) and bool((r.kl_div[:, 1] < r.kl_div[:, 0]).all())
    Fo, fo, co = oracle.fit_lr(ro["X"][1].cpu().numpy()[:, :6], ro["U"][1].cpu().numpy()[:, :5])
    assert relerr(F[1, :5], Fo) < 1e-8 and relerr(S[1, :5], co) < 1e-8
    # full horizon against the oracle: every policy output, KL and expected cost of two (condition, eta) problems
    for b, e in ((1, 0), (4, 1)):
        ref = oracle.ilqr_step(_oracle_problem(b, ro, dyn, costs, x0), float(eta[b, e]))
        assert relerr(r.K[b, e], ref.K) < 1e-8 and relerr(r.k[b, e], ref.k) < 1e-8
        assert relerr(r.covar[b, e], ref.covar) < 1e-8 and relerr(r.inv_covar[b, e], ref.inv_covar) < 1e-8
        assert abs(float(r.kl_div[b, e]) - ref.kl_div) < 1e-8 * max(1.0, abs(ref.kl_div))
        assert abs(float(r.expected_costs[b, e]) - ref.expected_costs) < 1e-8 * abs(ref.expected_costs)
    # the cooperative kernel and the generic kernel agree over the whole horizon (different summation orders)
    import os

    K_coop = r.K.clone()
    os.environ["DONK_ILQR_GENERIC"] = "1"
    try:
        rg = il.step(eta)
    finally:
        del os.environ["DONK_ILQR_GENERIC"]
    assert relerr(rg.K, K_coop.cpu().numpy()) < 1e-9
    # optimize (Brent over eta, E = 1 launches) for one condition against the oracle's search
    res, status, n_steps = il.optimize(1.0)
    assert int(status.abs().sum()) == 0
    ref, visited = oracle.ilqr_optimize(_oracle_problem(2, ro, dyn, costs, x0), 1.0)
    assert abs(float(res.eta[2, 0]) - ref.eta) < 1e-8 * ref.eta and relerr(res.K[2, 0], ref.K) < 1e-8
