"""Per-phase cycle breakdown of ilqr_warp_kernel (profiling build of the library with -DDONK_PHASE_PROF).

  nvcc ... -DDONK_PHASE_PROF -o profiles/_build/libdonk_prof.so donk_b200/csrc/*.cu   (done by this script
  when the file is missing and nvcc exists);  python profiles/prof_phases.py [B]
Each warp adds clock64() deltas between phase boundaries to a global counter; the table is cycles per warp and step.
Numbers from this build are diagnostic only (the counters perturb the kernel a little), never bench values.
"""
import ctypes, subprocess, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
LIB = ROOT / "profiles" / "_build" / "libdonk_prof.so"
if not LIB.exists():
    LIB.parent.mkdir(exist_ok=True)
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                    "-shared", "-DDONK_PHASE_PROF", "-o", str(LIB)] + [str(f) for f in sorted((ROOT / "donk_b200/csrc").glob("*.cu"))], check=True)
from donk_b200 import _lib, batched, synthetic  # noqa: E402
_lib.LIB_PATH = LIB
nat = _lib.native()
nat.lib.donk_phase_prof_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
dev = torch.device("cuda", 0)
T, dX, dU, N, E = 100, 20, 6, 64, 16
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
ro = synthetic.rollout_batch_torch(B, T, dX, dU, N, seed=1, device=dev)
F, f, S, info = batched.fit_lr(ro["X"], ro["U"])
p = dX + dU
C, c, cc = synthetic.l2_costs(T, dX, dU)
x0m, x0S = batched.state_fit_from_rollouts(ro["X"])
il = batched.BatchedILQR(F, f, S, ro["prev_K"], ro["prev_k"], torch.as_tensor(C, device=dev).expand(B, T + 1, p, p).contiguous(),
                         torch.as_tensor(c, device=dev).expand(B, T + 1, p).contiguous(),
                         torch.as_tensor(cc, device=dev).expand(B, T + 1).contiguous(), x0m, x0S, prev_covar=ro["prev_covar"])
eta = torch.as_tensor(np.logspace(-2, 2, E), device=dev).expand(B, E).contiguous()
il.step(eta); torch.cuda.synchronize()
nat.lib.donk_phase_prof_read(None, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); il.step(eta); e1.record(); torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 32)()
nat.lib.donk_phase_prof_read(out, 0)
names = {0: "B1  W=V[F|f] (+C loads issued)", 1: "B2  Q=C/eta+C_kl+F'W", 2: "bwd cta_sync+stage issue", 3: "B34 chol/gains", 4: "B5  V update",
         5: "bwd cp.async wait+cta_sync", 6: "F1  K[S|mu] (+policy wait, C loads issued)", 7: "F2  Suu", 8: "cost + tcov/tmean",
         9: "fwd mid wait+cta_sync", 10: "KL product + F3 (F[S|mu])'", 11: "F4  F S F' + D, KL tail, policy stage",
         12: "prologues", 13: "fwd end cta_sync+stage issue"}
warps = B * E / 8  # every 8th CTA records (PHASE_MARK)
tot = sum(out)
print(f"B={B} E={E}: {e0.elapsed_time(e1):.2f} ms (profiling build); cycles per warp and step:")
for i in range(14):
    per = out[i] / warps / T
    print(f"  {i:2d} {names[i]:40s} {per:9.0f}  {100 * out[i] / tot:5.1f} %")
print(f"  total per warp and step pair: {tot / warps / T:.0f} cycles")
