"""Timings of the secondary paths on the GPU: ILQR.optimize (config 3), generic kernels at config-5 / config-2 dims."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from donk_b200 import _lib, batched, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out

def build(B, T, dX, dU, N, prior=None):
    ro = synthetic.rollout_batch_torch(B, T, dX, dU, N, seed=7, device=dev)
    p = dX + dU
    ms, (F, f, S, info) = timed(lambda: batched.fit_lr(ro["X"], ro["U"], prior))
    assert int(info.sum()) == 0
    C, c, cc = synthetic.l2_costs(T, dX, dU)
    x0m, x0S = batched.state_fit_from_rollouts(ro["X"])
    il = batched.BatchedILQR(F, f, S, ro["prev_K"], ro["prev_k"],
                             torch.as_tensor(C, device=dev).expand(B, T + 1, p, p).contiguous(),
                             torch.as_tensor(c, device=dev).expand(B, T + 1, p).contiguous(),
                             torch.as_tensor(cc, device=dev).expand(B, T + 1).contiguous(), x0m, x0S, prev_covar=ro["prev_covar"])
    return ro, il, ms

import os as _os
if _os.environ.get("PROF_ALT_LIB"):  # experiments: time an alternative build of the library (profiles/build_alt.py)
    _lib.LIB_PATH = (Path(__file__).resolve().parent.parent / _os.environ["PROF_ALT_LIB"]).resolve()
what = sys.argv[1]
if what == "optimize":
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    ro, il, fit_ms = build(B, 100, 20, 6, 64)
    ms, (res, status, n_steps) = timed(lambda: il.optimize(1.0, raise_on_failure=False), 2)
    print(f"optimize config3 B={B}: {ms:.1f} ms, {n_steps} batched steps, {B / (ms * 1e-3):.0f} optimize/s, status counts",
          torch.bincount(status, minlength=4).tolist(), "fit_ms", fit_ms)
elif what == "config5":
    B, T = int(sys.argv[2]), int(sys.argv[3])
    ro, il, fit_ms = build(B, T, 64, 16, 128)
    print(f"fit_lr (64,16) N=128 B={B} T={T}: {fit_ms:.1f} ms = {B * T / (fit_ms * 1e-3):.0f} fits/s")
    eta = torch.ones(B, 1, dtype=torch.float64, device=dev)
    ms, r = timed(lambda: il.step(eta), 2)
    print(f"ILQR.step (64,16) B={B} T={T} E=1: {ms:.1f} ms = {B / (ms * 1e-3):.1f} solves/s, info {int(r.info.sum())}")
    ms, (res, status, n_steps) = timed(lambda: il.optimize(1.0, raise_on_failure=False), 1)
    print(f"ILQR.optimize (64,16) B={B}: {ms:.1f} ms, {n_steps} steps, status", torch.bincount(status, minlength=4).tolist())
elif what == "config5fit":
    B, T = int(sys.argv[2]), int(sys.argv[3])
    dX, dU, N, K = 64, 16, 128, 16
    d = 2 * dX + dU
    ro = synthetic.rollout_batch_torch(B, T, dX, dU, N, seed=7, device=dev)
    ms, (F, f, S, info) = timed(lambda: batched.fit_lr(ro["X"], ro["U"]))
    print(f"fit_lr (64,16) N=128 no prior B={B} T={T}: {ms:.1f} ms = {B * T / (ms * 1e-3):.0f} fits/s, info counts", torch.bincount(info.flatten(), minlength=3).tolist())
    pool = batched.to_transitions(ro["X"][:min(B, 32)].reshape(-1, T + 1, dX), ro["U"][:min(B, 32)].reshape(-1, T, dU))
    rng = np.random.default_rng(0)
    idx = rng.choice(pool.shape[0], K, replace=False)
    gm = batched.DeviceGMM(K, max_iter=5).set_params(np.full(K, 1 / K), pool[idx].cpu().numpy(), precisions=np.broadcast_to(np.eye(d), (K, d, d)).copy())
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter(); gm.fit(pool); torch.cuda.synchronize()
    print(f"GMM fit on {pool.shape[0]} rows d={d} K={K}: {gm.n_iter} iterations, {1e3 * (time.perf_counter() - t0):.1f} ms")
    ms, (F, f, S, info) = timed(lambda: batched.fit_lr(ro["X"], ro["U"], gm), 2)
    print(f"fit_lr (64,16) N=128 K=16 prior B={B} T={T}: {ms:.1f} ms = {B * T / (ms * 1e-3):.0f} fits/s, info counts", torch.bincount(info.flatten(), minlength=3).tolist())
elif what == "config2":
    B = int(sys.argv[2])
    T, dX, dU, N, K = 100, 14, 7, 50, 4
    ro = synthetic.rollout_batch_torch(B, T, dX, dU, N, seed=7, device=dev)
    pool = batched.to_transitions(ro["X"][:8].reshape(-1, T + 1, dX), ro["U"][:8].reshape(-1, T, dU))
    d = 2 * dX + dU
    rng = np.random.default_rng(0)
    idx = rng.choice(pool.shape[0], K, replace=False)
    gm = batched.DeviceGMM(K).set_params(np.full(K, 1 / K), pool[idx].cpu().numpy(), precisions=np.broadcast_to(np.eye(d), (K, d, d)).copy())
    t0 = time.perf_counter(); gm.fit(pool); torch.cuda.synchronize()
    print(f"GMM fit on {pool.shape[0]} rows d={d} K={K}: {gm.n_iter} iterations, {1e3 * (time.perf_counter() - t0):.1f} ms, converged {gm.converged}")
    ms, (F, f, S, info) = timed(lambda: batched.fit_lr(ro["X"], ro["U"], gm))
    print(f"fit_lr + GMM prior (14,7) N=50 K=4, B={B}: {ms:.1f} ms = {B * T / (ms * 1e-3):.0f} fits/s, info {int(info.sum())}")
    ms, _ = timed(lambda: batched.fit_lr(ro["X"], ro["U"]))
    print(f"fit_lr no prior  (14,7) N=50, B={B}: {ms:.1f} ms = {B * T / (ms * 1e-3):.0f} fits/s")
elif what == "kmeans":
    n, d, K = int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 72, int(sys.argv[4]) if len(sys.argv) > 4 else 16
    X, _, _, _ = synthetic.gmm_pool_torch(n, d, K, seed=5, device=dev)
    km = batched.DeviceKMeans(K, random_state=0)
    c0 = X[:K].contiguous()
    nat = _lib.native()
    ws = torch.empty(nat.size("kmeans_workspace_bytes", d, K) // 8 + 1, dtype=torch.float64, device=dev)
    sums = torch.empty(K * d + K, dtype=torch.float64, device=dev)
    ms_a, _ = timed(lambda: km._pass(X, c0, labels=True))
    ms_u, _ = timed(lambda: km._pass(X, c0, sums=sums, ws=ws))
    gb = n * d * 8 / 1e9
    print(f"kmeans n={n} d={d} K={K}: labels-only pass {ms_a:.2f} ms ({gb / ms_a * 1e3:.0f} GB/s), Lloyd pass (labels+sums) {ms_u:.2f} ms ({gb / ms_u * 1e3:.0f} GB/s)")
    t0 = time.perf_counter(); c = km._seed(X); torch.cuda.synchronize()
    print(f"  greedy k-means++ seeding: {1e3 * (time.perf_counter() - t0):.1f} ms")
    t0 = time.perf_counter(); km.fit(X); torch.cuda.synchronize()
    print(f"  DeviceKMeans.fit: {km.n_iter} Lloyd iterations, {1e3 * (time.perf_counter() - t0):.1f} ms, counts", torch.bincount(km.labels.long(), minlength=K).tolist())
elif what == "config1":
    # BASELINE configs[0]: T=50, dX=6, dU=2, N=20 - one condition (latency) and a 4096-condition batch (throughput)
    for B in (1, 4096):
        ro, il, fit_ms = build(B, 50, 6, 2, 20)
        ms, (res, status, n_steps) = timed(lambda: il.optimize(1.0, raise_on_failure=False), 5)
        eta = torch.ones(B, 1, dtype=torch.float64, device=dev)
        ms_s, _ = timed(lambda: il.step(eta), 20)
        print(f"config1 B={B}: fit_lr {fit_ms * 1e3:.0f} us, ILQR.step {ms_s * 1e3:.0f} us, ILQR.optimize {ms * 1e3:.0f} us "
              f"({n_steps} steps) = {B / (ms * 1e-3):.0f} optimize/s, status", torch.bincount(status, minlength=4).tolist())
elif what == "config3prior":
    # config-3 dimensions with a K-cluster prior (d = 46): what a GPS iteration with GMMPrior costs in fit_lr
    B, K = int(sys.argv[2]), int(sys.argv[3])
    T, dX, dU, N = 100, 20, 6, 64
    ro = synthetic.rollout_batch_torch(B, T, dX, dU, N, seed=7, device=dev)
    pool = batched.to_transitions(ro["X"][:16].reshape(-1, T + 1, dX), ro["U"][:16].reshape(-1, T, dU))
    d = 2 * dX + dU
    rng = np.random.default_rng(0)
    idx = rng.choice(pool.shape[0], K, replace=False)
    gm = batched.DeviceGMM(K).set_params(np.full(K, 1 / K), pool[idx].cpu().numpy(), precisions=np.broadcast_to(np.eye(d), (K, d, d)).copy())
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gm.fit(pool)
    ms, (F, f, S, info) = timed(lambda: batched.fit_lr(ro["X"], ro["U"], gm))
    print(f"fit_lr + GMM prior (20,6) N=64 K={K}, B={B}: {ms:.1f} ms = {B * T / (ms * 1e-3):.0f} fits/s, info {int(info.sum())}")
    ms, _ = timed(lambda: batched.fit_lr(ro["X"], ro["U"]))
    print(f"fit_lr no prior  (20,6) N=64, B={B}: {ms:.1f} ms = {B * T / (ms * 1e-3):.0f} fits/s")
elif what == "epass":
    # E pass (predict_proba) alone: n rows, d features, K clusters
    n, d, K = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    X, w0, m0, p0 = synthetic.gmm_pool_torch(n, d, K, seed=5, device=dev)
    gm = batched.DeviceGMM(K).set_params(w0, m0, precisions=p0)
    ms, resp = timed(lambda: gm.predict_proba(X), 3)
    fl = n * K * d * (d + 1)
    print(f"E pass n={n} d={d} K={K}: {ms:.2f} ms = {fl / ms / 1e9:.2f} TFLOP/s (triangular count)")
elif what == "tc":
    # tensor-core E pass: split copy, E pass alone, full EM iteration (tf32x3 E + fp64 M) vs the fp64 iteration
    n, d, K = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    X, w0, m0, p0 = synthetic.gmm_pool_torch(n, d, K, seed=5, device=dev)
    nat = _lib.native()
    gm = batched.DeviceGMM(K, tol=0.0, max_iter=1, precision="tf32x3").set_params(w0, m0, precisions=p0)
    ms_prep, tc = timed(lambda: gm._tc_setup(X), 2)
    ws = torch.empty(nat.size("gmm_workspace_bytes", n, d, K) // 8 + 1, dtype=torch.float64, device=dev)
    def estep():
        nat.call("gmm_tc_estep_f32", n, d, K, tc["planes"], tc["center"], gm.weights, gm.means, gm.precisions_cholesky,
                 ws, ws[n * K:], tc["work"], tc["work"].numel() * 8)
    ms_e, _ = timed(estep, 5)
    fl = n * K * d * (d + 1)
    print(f"tc E pass n={n} d={d} K={K}: split copy {ms_prep:.2f} ms (once per fit), E pass {ms_e:.2f} ms = "
          f"{fl / ms_e / 1e9:.1f} TFLOP/s algorithmic (triangular count), {3 * 2 * n * K * round(d / 8 + 0.49) * 8 * round(d / 8 + 0.49) * 8 / ms_e / 1e9:.0f} TFLOP/s issued TF32")
    ms64, r64 = timed(lambda: gm.predict_proba(X), 3)
    r32 = ws[:n * K].view(n, K)
    print(f"  fp64 E pass {ms64:.2f} ms; max |resp_tc - resp_fp64| = {float((r32 - r64).abs().max()):.2e}")
    stats = torch.empty(nat.size("gmm_stats_size", d, K), dtype=torch.float64, device=dev)
    ref = None
    for prec in ("fp64", "tf32x3-e", "tf32x3"):
        g = batched.DeviceGMM(K, tol=0.0, max_iter=1, precision=prec.split("-")[0]).set_params(w0, m0, precisions=p0)
        g.tc_m_pass = not prec.endswith("-e")
        g._lb = torch.zeros(1, dtype=torch.float64, device=dev); g._info = torch.zeros(K, dtype=torch.int32, device=dev)
        g._tc = g._tc_setup(X, m_pass=True) if prec != "fp64" else None
        g.em_iteration(X, float(n), ws, stats)
        res = (g.means.clone(), g.covariances.clone())
        if ref is None: ref = res
        err = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(res, ref)]
        g.set_params(w0, m0, precisions=p0)
        ms_it, _ = timed(lambda: g.em_iteration(X, float(n), ws, stats), 5)
        print(f"  EM iteration ({prec}): {ms_it:.2f} ms, lower bound {float(g._lb):.10f}; after one iteration vs fp64: means {err[0]:.2e} covariances {err[1]:.2e}")
        if prec == "tf32x3":
            tc = g._tc
            def mstep():
                nat.call("gmm_tc_mstep_f32", n, d, K, tc["mplanes"], tc["center"], ws, ws[n * K:], stats, g._shift, tc["mwork"], tc["mwork"].numel() * 8)
            ms_m, _ = timed(mstep, 5)
            print(f"  tc M pass alone (respT + mstep + reduce): {ms_m:.2f} ms = {3 * 2 * n * K * d * round((d + 1) / 16 + 0.49) * 16 / ms_m / 1e9:.0f} TFLOP/s issued TF32")
            import os
            Xc = X - tc["center"]
            R = ws[:n * K].view(n, K)
            S64 = torch.stack([Xc.T @ (Xc * R[:, k:k + 1]) for k in range(K)])
            nk64 = R.sum(0)
            for fl in (4, 16, 64, 256, 1024):
                os.environ["DONK_GMM_TC_FLUSH"] = str(fl)
                ms_f, _ = timed(mstep, 3)
                S32 = stats[K * (1 + d):-1].view(K, d, d)
                err = ((S32 - S64).abs().amax(dim=(1, 2)) / S64.abs().amax(dim=(1, 2))).max()
                print(f"    drain every {fl:4d} tiles: {ms_f:.2f} ms, max rel err of the scatter vs fp64 {float(err):.2e}, nk err {float(((stats[:K] - nk64).abs() / nk64).max()):.1e}")
            os.environ.pop("DONK_GMM_TC_FLUSH")
