"""Where does the k-means++ seeding spend its time (10 M x 72, K = 16)?  CUDA-event timing of each piece of one round."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from donk_b200 import batched, synthetic
dev = torch.device("cuda", 0)
n, d, K = 10_000_000, 72, 16
X, _, _, _ = synthetic.gmm_pool_torch(n, d, K, seed=5, device=dev)
km = batched.DeviceKMeans(K, random_state=0)
g = torch.Generator(device=dev); g.manual_seed(0)
_, closest, _ = km._pass(X, X[:1].contiguous(), mind2=True)
def t(name, fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) / reps:8.3f} ms")
    return r
cum = t("cumsum", lambda: torch.cumsum(closest, 0))
u = torch.rand(4, dtype=torch.float64, device=dev, generator=g) * cum[-1]
cand = t("searchsorted", lambda: torch.searchsorted(cum, u).clamp_(max=n - 1))
Xc = t("X[cand]", lambda: X[cand])
D = t("pass dist=True (4 candidates)", lambda: km._pass(X, Xc, dist=True)[2])
t("minimum", lambda: torch.minimum(D, closest[:, None], out=D))
t("D.sum(0)", lambda: D.sum(0))
t("ones @ D", lambda: torch.ones(1, n, dtype=torch.float64, device=dev) @ D)
best = D.sum(0).argmin().reshape(1)
t("index_select col", lambda: D.index_select(1, best)[:, 0].contiguous())
t("pass mind2 (1 centre)", lambda: km._pass(X, Xc[:1].contiguous(), mind2=True)[1])
t("whole _seed", lambda: km._seed(X), reps=2)
