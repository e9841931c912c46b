"""One ILQR.step launch at config-5 dimensions (for ncu captures of ilqr_coop_kernel):  python profiles/run_coop_once.py [B] [T] [E]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from donk_b200 import batched, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
E = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dX, dU, N = 64, 16, 128
ro = synthetic.rollout_batch_torch(B, T, dX, dU, N, seed=1, device=dev)
F, f, S, info = batched.fit_lr(ro["X"], ro["U"])
p = dX + dU
C, c, cc = synthetic.l2_costs(T, dX, dU)
x0m, x0S = batched.state_fit_from_rollouts(ro["X"])
il = batched.BatchedILQR(F, f, S, ro["prev_K"], ro["prev_k"], torch.as_tensor(C, device=dev).expand(B, T + 1, p, p).contiguous(),
                         torch.as_tensor(c, device=dev).expand(B, T + 1, p).contiguous(),
                         torch.as_tensor(cc, device=dev).expand(B, T + 1).contiguous(), x0m, x0S, prev_covar=ro["prev_covar"])
eta = torch.ones(B, E, dtype=torch.float64, device=dev)
for _ in range(2):
    r = il.step(eta)
torch.cuda.synchronize()
print("kl", float(r.kl_div[0, 0]), "info", int(r.info.sum()))
