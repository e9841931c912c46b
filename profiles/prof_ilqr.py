"""Small driver for ncu captures of the kernels (one launch each after a warm-up launch).

  python profiles/prof_ilqr.py ilqr   [B]   -> BatchedILQR.step, B conditions x 16 eta, T=100, dX=20, dU=6
  python profiles/prof_ilqr.py fit    [B]   -> fit_lr, N=64
  python profiles/prof_ilqr.py em     [n]   -> one EM iteration, d=72, K=16
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from donk_b200 import _lib, batched, synthetic  # noqa: E402
import os as _os
if _os.environ.get("PROF_ALT_LIB"):  # experiments: time an alternative build of the library
    _lib.LIB_PATH = (Path(__file__).resolve().parent.parent / _os.environ["PROF_ALT_LIB"]).resolve()

what = sys.argv[1] if len(sys.argv) > 1 else "ilqr"
dev = torch.device("cuda", 0)
T, dX, dU, N, E = 100, 20, 6, 64, 16
if what in ("ilqr", "fit"):
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 592
    ro = synthetic.rollout_batch_torch(B, T, dX, dU, N, seed=1, device=dev)
    for _ in range(2):
        F, f, S, info = batched.fit_lr(ro["X"], ro["U"])
    torch.cuda.synchronize()
    if what == "fit" and len(sys.argv) > 3:
        import os
        n = int(sys.argv[3])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            F, f, S, info = batched.fit_lr(ro["X"], ro["U"])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print("cfg", {k: v for k, v in os.environ.items() if k.startswith("DONK_")}, "B", B, "ms_per_launch", ms,
              "fits_per_s", B * T / (ms * 1e-3), "info", int(info.sum()))
    if what == "ilqr":
        p = dX + dU
        C, c, cc = synthetic.l2_costs(T, dX, dU)
        x0m, x0S = batched.state_fit_from_rollouts(ro["X"])
        il = batched.BatchedILQR(F, f, S, ro["prev_K"], ro["prev_k"],
                                 torch.as_tensor(C, device=dev).expand(B, T + 1, p, p).contiguous(),
                                 torch.as_tensor(c, device=dev).expand(B, T + 1, p).contiguous(),
                                 torch.as_tensor(cc, device=dev).expand(B, T + 1).contiguous(), x0m, x0S,
                                 prev_covar=ro["prev_covar"])
        eta = torch.as_tensor(np.logspace(-2, 2, E), device=dev).expand(B, E).contiguous()
        for _ in range(2):
            r = il.step(eta)
        torch.cuda.synchronize()
        if len(sys.argv) > 3:  # timing mode: ms per step over n launches
            n = int(sys.argv[3])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                r = il.step(eta)
            e1.record()
            torch.cuda.synchronize()
            import os
            print("cfg", {k: v for k, v in os.environ.items() if k.startswith("DONK_")}, "B", B,
                  "ms_per_step", e0.elapsed_time(e1) / n, "solves_per_s", B * E / (e0.elapsed_time(e1) / n * 1e-3))
        print("kl", float(r.kl_div[0, 0]), "info", int(r.info.sum()))
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
    d, K = 72, 16
    X, w0, m0, p0 = synthetic.gmm_pool_torch(n, d, K, seed=4000, device=dev)
    gm = batched.DeviceGMM(K, tol=0.0, max_iter=2).set_params(w0, m0, precisions=p0)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gm.fit(X)
    torch.cuda.synchronize()
    print("lb", gm.lower_bound)
    if len(sys.argv) > 3:
        import os
        nat = _lib.native()
        o = dict(dtype=torch.float64, device=dev)
        ws = torch.empty(nat.size("gmm_workspace_bytes", n, d, K) // 8 + 1, **o)
        stats = torch.empty(nat.size("gmm_stats_size", d, K), **o)
        gm._lb = torch.zeros(1, **o); gm._info = torch.zeros(K, dtype=torch.int32, device=dev)
        it = int(sys.argv[3])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it):
            gm.em_iteration(X, float(n), ws, stats)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / it
        print("cfg", {k: v for k, v in os.environ.items() if k.startswith("DONK_")}, "n", n, "ms_per_iter", ms,
              "rows_per_s", n / (ms * 1e-3), "lb", float(gm._lb))
