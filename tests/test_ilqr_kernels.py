"""Kernel logic of the iLQR step (donk_b200/csrc/ilqr.cuh) under host emulation vs the oracle."""
import numpy as np
import pytest
import torch

import oracle
from donk_testutil import load_golden, relerr, set_reverse, to_np
from oracle.lqg import ILQRProblem

TOL = 1e-9


def _problem(g):
    return ILQRProblem(g["F"], g["f"], g["dyn_covar"], g["C"], g["c"], g["cc"], g["prev_K"], g["prev_k"],
                       g["prev_covar"], g["prev_inv_covar"], g["x0_mean"], g["x0_covar"])


def _batched(g, reps=1):
    from donk_b200.batched import BatchedILQR

    r = lambda a: np.broadcast_to(a, (reps,) + a.shape).copy()  # noqa: E731
    return BatchedILQR(r(g["F"]), r(g["f"]), r(g["dyn_covar"]), r(g["prev_K"]), r(g["prev_k"]), r(g["C"]), r(g["c"]),
                       r(g["cc"]), r(g["x0_mean"]), r(g["x0_covar"]), prev_covar=r(g["prev_covar"]))


@pytest.mark.parametrize("reverse", [0, 1])
def test_backward_forward_small(backend, reverse):
    from donk_b200 import batched

    set_reverse(backend, reverse)
    g = load_golden("lqg_small")
    K, k, cov, inv, info = batched.lqr_backward(g["bw_F"][None], g["bw_f"][None], g["bw_C"][None], g["bw_c"][None])
    assert int(info.sum()) == 0
    assert relerr(K[0], g["bw_K"]) < TOL and relerr(k[0], g["bw_k"]) < TOL
    assert relerr(cov[0], g["bw_covar"]) < TOL and relerr(inv[0], g["bw_inv_covar"]) < TOL
    K, k, cov, inv, info = batched.lqr_backward(g["bw_F"][None], g["bw_f"][None], g["bw_C"][None], g["bw_c"][None], gamma=0.9)
    assert relerr(K[0], g["bwg_K"]) < TOL and relerr(cov[0], g["bwg_covar"]) < TOL
    mean, covar = batched.lqr_forward(g["fw_F"][None], g["fw_f"][None], g["fw_dyn_covar"][None], g["fw_K"][None],
                                      g["fw_k"][None], g["fw_pol_covar"][None], g["fw_x0_mean"][None], g["fw_x0_covar"][None])
    assert relerr(mean[0], g["fw_mean"]) < TOL and relerr(covar[0], g["fw_covar"]) < TOL
    c = covar[0].cpu().numpy()
    assert np.array_equal(c, np.swapaxes(c, 1, 2))


def test_helpers(backend):
    from donk_b200 import batched

    g = load_golden("lqg_small")
    C, c = batched.extended_costs_kl(g["ex_K"], g["ex_k"], g["ex_inv_covar"])
    assert relerr(C, g["ex_C"]) < TOL and relerr(c, g["ex_c"]) < TOL
    inv, chol, hld = batched.spd_factor(torch.as_tensor(g["ex_covar"]), want_chol=True)
    assert relerr(inv, g["ex_inv_covar"]) < TOL
    assert relerr(chol, np.linalg.cholesky(g["ex_covar"])) < TOL
    assert relerr(hld, np.log(np.diagonal(np.linalg.cholesky(g["ex_covar"]), axis1=1, axis2=2)).sum(-1)) < TOL
    kl = batched.kl_divergence_action(g["kl_X"][None], g["kl_K"][None], g["kl_k"][None], g["kl_covar"][None],
                                      g["kl_pK"][None], g["kl_pk"][None], g["kl_pcovar"][None])
    assert abs(float(kl[0]) - 3.914744335914712) < 1e-12
    with pytest.raises(np.linalg.LinAlgError):
        batched.spd_factor(torch.as_tensor(-np.eye(3)[None]))


@pytest.mark.parametrize("case,reverse", [("ilqr_c1", 0), ("ilqr_c1", 1), ("ilqr_c2", 0)])
def test_step_matches_reference(backend, case, reverse):
    set_reverse(backend, reverse)
    g = load_golden(case)
    il = _batched(g, reps=3)
    assert relerr(il.C_kl[1], g["C_kl"]) < TOL and relerr(il.c_kl[2], g["c_kl"]) < TOL
    etas = g["etas"]
    r = il.step(torch.as_tensor(etas).expand(3, len(etas)), want_traj=True)
    assert int(r.info.sum()) == 0
    for i in range(len(etas)):
        for b in (0, 2):
            assert relerr(r.K[b, i], g[f"s{i}_K"]) < TOL and relerr(r.k[b, i], g[f"s{i}_k"]) < TOL
            assert relerr(r.covar[b, i], g[f"s{i}_covar"]) < TOL
            assert relerr(r.inv_covar[b, i], g[f"s{i}_inv_covar"]) < TOL
            assert relerr(r.traj_mean[b, i], g[f"s{i}_mean"]) < TOL
            assert abs(float(r.kl_div[b, i]) - g[f"s{i}_kl"]) < 1e-9 * max(1.0, abs(g[f"s{i}_kl"]))
            assert abs(float(r.expected_costs[b, i]) - g[f"s{i}_cost"]) < 1e-9 * abs(g[f"s{i}_cost"])
            if f"s{i}_tcovar" in g:
                assert relerr(r.traj_covar[b, i], g[f"s{i}_tcovar"]) < TOL


def test_optimize_follows_reference_brent_path(backend):
    g = load_golden("ilqr_c1")
    il = _batched(g, reps=2)
    for j in (0, 1):
        res, status, n_steps = il.optimize(float(g[f"o{j}_kl_step"]))
        assert n_steps == len(g[f"o{j}_visited"])  # same number of ILQR.step evaluations as the reference
        assert status.tolist() == [0, 0]
        assert abs(float(res.eta[0, 0]) - g[f"o{j}_eta"]) < 1e-9 * g[f"o{j}_eta"]
        assert relerr(res.K[1, 0], g[f"o{j}_K"]) < 1e-8 and relerr(res.k[1, 0], g[f"o{j}_k"]) < 1e-8
        assert abs(float(res.kl_div[0, 0]) - g[f"o{j}_kl"]) < 1e-8
    # kl_step so large that the constraint is inactive at min_eta (lqg.py:108-109)
    res, status, n_steps = il.optimize(1e9)
    assert status.tolist() == [1, 1] and float(res.eta[0, 0]) == 1e-6
    with pytest.raises(ValueError):
        il.optimize(1e-30, max_eta=1e-3)


def test_multi_condition_packing_and_active_mask(backend):
    """E=1 packs several conditions per CTA; masked conditions keep their outputs untouched."""
    from donk_b200 import synthetic
    from donk_b200.batched import BatchedILQR

    T, dX, dU, N = 12, 5, 3, 12
    conds = [synthetic.condition(9, i, T, dX, dU, N) for i in range(11)]
    fits = [oracle.fit_lr(c.X, c.U) for c in conds]
    x0 = [oracle.state_fit(c.X[:, 0]) for c in conds]
    st = lambda xs: np.stack(xs)  # noqa: E731
    il = BatchedILQR(st([f[0] for f in fits]), st([f[1] for f in fits]), st([f[2] for f in fits]),
                     st([c.prev_K for c in conds]), st([c.prev_k for c in conds]), st([c.C for c in conds]),
                     st([c.c for c in conds]), st([c.cc for c in conds]), st([m for m, _ in x0]),
                     st([S for _, S in x0]), prev_covar=st([c.prev_covar for c in conds]))
    eta = torch.as_tensor(np.logspace(-1, 2, 11))
    active = torch.ones(11, dtype=torch.int32)
    active[[3, 7]] = 0
    active = active.to(backend.device)
    r = il.step(eta, active=None)
    kl_all = r.kl_div[:, 0].clone()
    r.kl_div.fill_(-1.0)
    r = il.step(eta, active=active)
    assert float(r.kl_div[3, 0]) == -1.0 and float(r.kl_div[7, 0]) == -1.0
    for b, c in enumerate(conds):
        inv = np.linalg.inv(c.prev_covar)
        prob = ILQRProblem(*fits[b], c.C, c.c, c.cc, c.prev_K, c.prev_k, c.prev_covar, inv, *x0[b])
        ref = oracle.ilqr_step(prob, float(eta[b]))
        assert abs(float(kl_all[b]) - ref.kl_div) < 1e-9 * max(1, abs(ref.kl_div))
        if b not in (3, 7):
            assert relerr(r.K[b, 0], ref.K) < TOL
            assert abs(float(r.expected_costs[b, 0]) - ref.expected_costs) < 1e-9 * abs(ref.expected_costs)


@pytest.mark.parametrize("dims,reverse", [((20, 6), 0), ((20, 6), 1), ((6, 2), 0), ((14, 7), 1)])
def test_warp_kernel_dims_vs_oracle(backend, dims, reverse):
    """The compile-time-dimension warp kernel (ilqr_warp.cuh): several eta per condition, several conditions,
    a masked condition, and E=1 packing (several conditions per CTA)."""
    from donk_b200 import synthetic
    from donk_b200.batched import BatchedILQR

    set_reverse(backend, reverse)
    dX, dU = dims
    T, N, B = 7, dX + dU + 6, 5
    conds = [synthetic.condition(12, i, T, dX, dU, N) for i in range(B)]
    fits = [oracle.fit_lr(c.X, c.U) for c in conds]
    x0 = [oracle.state_fit(c.X[:, 0]) for c in conds]
    st = lambda xs: np.stack(xs)  # noqa: E731
    il = BatchedILQR(st([f[0] for f in fits]), st([f[1] for f in fits]), st([f[2] for f in fits]),
                     st([c.prev_K for c in conds]), st([c.prev_k for c in conds]), st([c.C for c in conds]),
                     st([c.c for c in conds]), st([c.cc for c in conds]), st([m for m, _ in x0]),
                     st([S for _, S in x0]), prev_covar=st([c.prev_covar for c in conds]))
    etas = np.array([1e-3, 0.3, 2.0, 50.0, 1e5])
    r = il.step(torch.as_tensor(etas).expand(B, len(etas)), want_traj=True)
    assert int(r.info.sum()) == 0
    probs = []
    for b, c in enumerate(conds):
        probs.append(ILQRProblem(*fits[b], c.C, c.c, c.cc, c.prev_K, c.prev_k, c.prev_covar,
                                 np.linalg.inv(c.prev_covar), *x0[b]))
    for b in (0, B - 1):
        for e, eta in enumerate(etas):
            ref = oracle.ilqr_step(probs[b], float(eta))
            assert relerr(r.K[b, e], ref.K) < TOL and relerr(r.k[b, e], ref.k) < TOL
            assert relerr(r.covar[b, e], ref.covar) < TOL and relerr(r.inv_covar[b, e], ref.inv_covar) < TOL
            assert relerr(r.traj_mean[b, e], ref.mean) < TOL and relerr(r.traj_covar[b, e], ref.traj_covar) < TOL
            assert abs(float(r.kl_div[b, e]) - ref.kl_div) < 1e-9 * max(1.0, abs(ref.kl_div))
            assert abs(float(r.expected_costs[b, e]) - ref.expected_costs) < 1e-9 * abs(ref.expected_costs)
    # E = 1: conditions packed per CTA, one masked out
    active = torch.ones(B, dtype=torch.int32)
    active[2] = 0
    r1 = il.step(torch.as_tensor(etas), active=active.to(backend.device))
    for b in (0, 1, 3, 4):
        ref = oracle.ilqr_step(probs[b], float(etas[b]))
        assert relerr(r1.K[b, 0], ref.K) < TOL
        assert abs(float(r1.kl_div[b, 0]) - ref.kl_div) < 1e-9 * max(1.0, abs(ref.kl_div))


def _coop_case(dX, dU, T, N, B, cfg=5):
    from donk_b200 import synthetic
    from donk_b200.batched import BatchedILQR

    conds = [synthetic.condition(cfg, i, T, dX, dU, N) for i in range(B)]
    fits = [oracle.fit_lr(c.X, c.U) for c in conds]
    x0 = [oracle.state_fit(c.X[:, 0]) for c in conds]
    st = lambda xs: np.stack(xs)  # noqa: E731
    il = BatchedILQR(st([f[0] for f in fits]), st([f[1] for f in fits]), st([f[2] for f in fits]),
                     st([c.prev_K for c in conds]), st([c.prev_k for c in conds]), st([c.C for c in conds]),
                     st([c.c for c in conds]), st([c.cc for c in conds]), st([m for m, _ in x0]),
                     st([S for _, S in x0]), prev_covar=st([c.prev_covar for c in conds]))
    probs = [ILQRProblem(*fits[b], c.C, c.c, c.cc, c.prev_K, c.prev_k, c.prev_covar, np.linalg.inv(c.prev_covar), *x0[b])
             for b, c in enumerate(conds)]
    return il, probs


def _check_step(r, b, e, ref, traj=False):
    assert relerr(r.K[b, e], ref.K) < TOL and relerr(r.k[b, e], ref.k) < TOL
    assert relerr(r.covar[b, e], ref.covar) < TOL and relerr(r.inv_covar[b, e], ref.inv_covar) < TOL
    assert abs(float(r.kl_div[b, e]) - ref.kl_div) < 1e-9 * max(1.0, abs(ref.kl_div))
    assert abs(float(r.expected_costs[b, e]) - ref.expected_costs) < 1e-9 * abs(ref.expected_costs)
    if traj:
        assert relerr(r.traj_mean[b, e], ref.mean) < TOL and relerr(r.traj_covar[b, e], ref.traj_covar) < TOL
        c = to_np(r.traj_covar[b, e])
        assert np.array_equal(c, np.swapaxes(c, 1, 2))


@pytest.mark.parametrize("dims,G,reverse", [((16, 8), 8, 0), ((16, 8), 3, 1), ((16, 8), 16, 0), ((16, 8), 1, 0)])
def test_coop_kernel_small_dims_vs_oracle(backend, monkeypatch, dims, G, reverse):
    """The CTA-per-problem kernel (ilqr_coop.cuh) at small instantiations: every output of ILQR.step incl. the
    trajectory, several eta per condition, a masked condition, different warp counts (tile schedules)."""
    set_reverse(backend, reverse)
    monkeypatch.setenv("DONK_COOP_WARPS", str(G))
    dX, dU = dims
    T, B = 6, 3
    il, probs = _coop_case(dX, dU, T, dX + dU + 9, B, cfg=15)
    etas = np.array([1e-3, 0.7, 40.0, 1e6])
    r = il.step(torch.as_tensor(etas).expand(B, len(etas)), want_traj=True)
    assert int(r.info.sum()) == 0
    for b in (0, B - 1):
        for e, eta in enumerate(etas):
            _check_step(r, b, e, oracle.ilqr_step(probs[b], float(eta)), traj=True)
    active = torch.ones(B, dtype=torch.int32)
    active[1] = 0
    r1 = il.step(torch.as_tensor(etas[:B]), active=active.to(backend.device))  # the FULL instantiation (no trajectory)
    for b in (0, 2):
        _check_step(r1, b, 0, oracle.ilqr_step(probs[b], float(etas[b])))
    # the plain LQR entry points run through the same kernel (partial flags)
    from donk_b200 import batched

    p0 = probs[0]
    K, k, cov, inv, info = batched.lqr_backward(p0.F[None], p0.f[None], p0.C[None], p0.c[None], gamma=0.9)
    ref = oracle.backward(p0.F, p0.f, p0.C, p0.c, gamma=0.9)
    assert relerr(K[0], ref[0]) < TOL and relerr(k[0], ref[1]) < TOL and relerr(cov[0], ref[2]) < TOL


def test_coop_kernel_config5_dims(backend):
    """dX=64, dU=16 (BASELINE config 5) on the CTA-per-problem kernel, short horizon (the T=500 comparison is the
    -m gpu test in test_full_size_gpu.py)."""
    il, probs = _coop_case(64, 16, 3, 100, 2)
    etas = np.array([0.5, 20.0])
    r = il.step(torch.as_tensor(etas).expand(2, 2))
    assert int(r.info.sum()) == 0
    for b in range(2):
        for e, eta in enumerate(etas):
            _check_step(r, b, e, oracle.ilqr_step(probs[b], float(eta)))


def test_generic_kernel_config5_dims(backend, monkeypatch):
    """dX=64, dU=16 on the generic CTA kernel (what any dimension without an instantiation runs)."""
    monkeypatch.setenv("DONK_ILQR_GENERIC", "1")
    il, probs = _coop_case(64, 16, 3, 100, 2)
    etas = np.array([0.5, 20.0])
    r = il.step(torch.as_tensor(etas).expand(2, 2))
    assert int(r.info.sum()) == 0
    for b in range(2):
        for e, eta in enumerate(etas):
            _check_step(r, b, e, oracle.ilqr_step(probs[b], float(eta)))


@pytest.mark.parametrize("dims", [(20, 6), (5, 3)])
def test_step_f32_tolerance(backend, dims):
    """Single-precision step (donk_ilqr_step_f32) against the fp64 oracle: the stated fp32 tolerance."""
    from donk_b200 import synthetic
    from donk_b200.batched import BatchedILQR

    dX, dU = dims
    T, N, B = 10, dX + dU + 10, 2
    conds = [synthetic.condition(14, i, T, dX, dU, N) for i in range(B)]
    fits = [oracle.fit_lr(c.X, c.U) for c in conds]
    x0 = [oracle.state_fit(c.X[:, 0]) for c in conds]
    st = lambda xs: np.stack(xs)  # noqa: E731
    il = BatchedILQR(st([f[0] for f in fits]), st([f[1] for f in fits]), st([f[2] for f in fits]),
                     st([c.prev_K for c in conds]), st([c.prev_k for c in conds]), st([c.C for c in conds]),
                     st([c.c for c in conds]), st([c.cc for c in conds]), st([m for m, _ in x0]),
                     st([S for _, S in x0]), prev_covar=st([c.prev_covar for c in conds]))
    etas = np.array([1e-2, 1.0, 1e2])
    r = il.step_f32(torch.as_tensor(etas).expand(B, 3))
    assert r.K.dtype == torch.float32 and int(r.info.sum()) == 0
    for b, c in enumerate(conds):
        prob = ILQRProblem(*fits[b], c.C, c.c, c.cc, c.prev_K, c.prev_k, c.prev_covar,
                           np.linalg.inv(c.prev_covar), *x0[b])
        for e, eta in enumerate(etas):
            ref = oracle.ilqr_step(prob, float(eta))
            assert relerr(r.K[b, e].double(), ref.K) < 2e-3
            assert abs(float(r.kl_div[b, e]) - ref.kl_div) < 5e-3 * max(1.0, abs(ref.kl_div))
            assert abs(float(r.expected_costs[b, e]) - ref.expected_costs) < 5e-3 * abs(ref.expected_costs)


@pytest.mark.parametrize("B,E,T", [(1, 1, 1), (3, 17, 2), (5, 3, 4), (2, 9, 3)])
def test_warp_kernel_edge_shapes(backend, B, E, T):
    """Ragged launches of the everything-enabled (FULL) warp kernel at (6,2): one problem with a single time step,
    E that is not a multiple of the eta candidates per CTA (17 = 8 + 8 + 1), fewer candidates than a CTA holds."""
    from donk_b200 import synthetic
    from donk_b200.batched import BatchedILQR

    dX, dU, N = 6, 2, 12
    conds = [synthetic.condition(17, i, T, dX, dU, N) for i in range(B)]
    fits = [oracle.fit_lr(c.X, c.U) for c in conds]
    x0 = [oracle.state_fit(c.X[:, 0]) for c in conds]
    st = lambda xs: np.stack(xs)  # noqa: E731
    il = BatchedILQR(st([f[0] for f in fits]), st([f[1] for f in fits]), st([f[2] for f in fits]),
                     st([c.prev_K for c in conds]), st([c.prev_k for c in conds]), st([c.C for c in conds]),
                     st([c.c for c in conds]), st([c.cc for c in conds]), st([m for m, _ in x0]),
                     st([S for _, S in x0]), prev_covar=st([c.prev_covar for c in conds]))
    etas = np.logspace(-2, 3, E)
    r = il.step(torch.as_tensor(etas).expand(B, E))
    assert int(r.info.sum()) == 0
    for b, c in enumerate(conds):
        prob = ILQRProblem(*fits[b], c.C, c.c, c.cc, c.prev_K, c.prev_k, c.prev_covar, np.linalg.inv(c.prev_covar), *x0[b])
        for e in range(E):
            ref = oracle.ilqr_step(prob, float(etas[e]))
            assert relerr(r.K[b, e], ref.K) < TOL and relerr(r.covar[b, e], ref.covar) < TOL
            assert abs(float(r.kl_div[b, e]) - ref.kl_div) < 1e-9 * max(1.0, abs(ref.kl_div))
            assert abs(float(r.expected_costs[b, e]) - ref.expected_costs) < 1e-9 * abs(ref.expected_costs)


def test_warp_kernel_reports_quu_not_positive_definite(backend):
    """lqg.py:165: scipy's solve(assume_a="pos") raises LinAlgError; the kernel flags the problem in ``info`` with
    1 + t of the first failing step (the backward pass starts at t = T-1) and the healthy problems stay correct."""
    from donk_b200 import synthetic
    from donk_b200.batched import BatchedILQR

    dX, dU, T, N = 6, 2, 5, 12
    conds = [synthetic.condition(19, i, T, dX, dU, N) for i in range(3)]
    fits = [oracle.fit_lr(c.X, c.U) for c in conds]
    x0 = [oracle.state_fit(c.X[:, 0]) for c in conds]
    st = lambda xs: np.stack(xs)  # noqa: E731
    C = st([c.C for c in conds])
    C[1] = -C[1]  # negative-definite costs for condition 1: Quu = C_uu / eta + C_kl,uu + ... fails for small eta
    il = BatchedILQR(st([f[0] for f in fits]), st([f[1] for f in fits]), st([f[2] for f in fits]),
                     st([c.prev_K for c in conds]), st([c.prev_k for c in conds]), C,
                     st([c.c for c in conds]), st([c.cc for c in conds]), st([m for m, _ in x0]),
                     st([S for _, S in x0]), prev_covar=st([c.prev_covar for c in conds]))
    etas = np.array([1e-3, 1.0])
    r = il.step(torch.as_tensor(etas).expand(3, 2))
    info = r.info.cpu().numpy()
    assert info[0].tolist() == [0, 0] and info[2].tolist() == [0, 0]
    assert info[1, 0] == T  # eta = 1e-3: -C_uu / eta dominates at the very first step (t = T-1)
    for b in (0, 2):
        c = conds[b]
        prob = ILQRProblem(*fits[b], c.C, c.c, c.cc, c.prev_K, c.prev_k, c.prev_covar, np.linalg.inv(c.prev_covar), *x0[b])
        ref = oracle.ilqr_step(prob, 1.0)
        assert relerr(r.K[b, 1], ref.K) < TOL and abs(float(r.kl_div[b, 1]) - ref.kl_div) < 1e-9 * max(1.0, abs(ref.kl_div))


def test_host_pipelined_step_equals_device_step(backend):
    """HostPipelinedILQR (chunked double-buffered upload, ragged last chunk, inputs changed in place between steps) returns
    exactly what BatchedILQR.step returns for the same conditions."""
    from donk_b200.batched import BatchedILQR, HostPipelinedILQR

    g = load_golden("ilqr_c1")
    B, E = 5, 3
    rng = np.random.default_rng(3)
    r = lambda a: np.broadcast_to(a, (B,) + a.shape).copy()  # noqa: E731
    host = {k: torch.as_tensor(r(g[k])) for k in HostPipelinedILQR.KEYS}
    host["f"] += torch.as_tensor(0.01 * rng.standard_normal(tuple(host["f"].shape)))   # conditions differ
    if backend.device.type == "cuda":
        host = {k: v.pin_memory() for k, v in host.items()}
    eta = torch.as_tensor(np.tile(np.array([0.5, 1.0, 4.0]), (B, 1)))
    pipe = HostPipelinedILQR(host, E, chunks=3)  # chunks of 2, 2 and 1 conditions
    for step in range(2):
        kl, cost, info = pipe.step(eta)
        il = BatchedILQR(host["F"], host["f"], host["dyn_covar"], host["prev_K"], host["prev_k"], host["C"], host["c"],
                         host["cc"], host["x0_mean"], host["x0_covar"], prev_covar=host["prev_covar"])
        ref = il.step(eta)
        assert int(info.sum()) == 0
        assert np.array_equal(kl.numpy(), to_np(ref.kl_div)) and np.array_equal(cost.numpy(), to_np(ref.expected_costs))
        host["f"] += 0.02  # a producer refills the host buffers in place: the next step must see the new inputs
