import sys, os
sys.path.insert(0, '/root/repo')
import torch
from donk_b200 import batched, synthetic
dev = torch.device('cuda', 0)
for (dX, dU, N, B) in [(14, 7, 50, 512), (14, 7, 64, 512), (20, 6, 50, 512), (6, 2, 20, 2048), (20, 6, 64, 512)]:
    ro = synthetic.rollout_batch_torch(B, 100, dX, dU, N, seed=7, device=dev)
    for _ in range(2): out = batched.fit_lr(ro["X"], ro["U"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = batched.fit_lr(ro["X"], ro["U"])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(os.environ.get("DONK_FIT_GENERIC", "warp"), (dX, dU, N, B), f"{ms:.2f} ms", f"{B*100/(ms*1e-3):.0f} fits/s", int(out[3].sum()))
