"""Time BatchedILQR.step with an alternative build of the library (experiments): python profiles/prof_altlib.py LIB B n"""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from donk_b200 import _lib, batched, synthetic  # noqa: E402
_lib.LIB_PATH = (ROOT / sys.argv[1]).resolve()
B, n = int(sys.argv[2]), int(sys.argv[3])
E = int(sys.argv[4]) if len(sys.argv) > 4 else 16
dev = torch.device("cuda", 0)
T, dX, dU, N = 100, 20, 6, 64
ro = synthetic.rollout_batch_torch(B, T, dX, dU, N, seed=1, device=dev)
F, f, S, info = batched.fit_lr(ro["X"], ro["U"])
p = dX + dU
C, c, cc = synthetic.l2_costs(T, dX, dU)
x0m, x0S = batched.state_fit_from_rollouts(ro["X"])
il = batched.BatchedILQR(F, f, S, ro["prev_K"], ro["prev_k"], torch.as_tensor(C, device=dev).expand(B, T + 1, p, p).contiguous(),
                         torch.as_tensor(c, device=dev).expand(B, T + 1, p).contiguous(),
                         torch.as_tensor(cc, device=dev).expand(B, T + 1).contiguous(), x0m, x0S, prev_covar=ro["prev_covar"])
eta = torch.as_tensor(np.logspace(-2, 2, E), device=dev).expand(B, E).contiguous()
for _ in range(2):
    r = il.step(eta)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    r = il.step(eta)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"{sys.argv[1]} B={B} E={E}: {ms:.2f} ms per step, {B * E / ms * 1e3:.0f} solves/s, kl[0,0]={float(r.kl_div[0, 0]):.6f} info={int(r.info.sum())}")
