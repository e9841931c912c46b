"""Per-phase cycle breakdown of ilqr_coop_kernel (profiling build of the library with -DDONK_PHASE_PROF).

The profiling library is cross-compiled where nvcc exists (profiles/_build/libdonk_prof.so, shipped to the GPU box with the
snapshot):  python profiles/prof_phases_coop.py --build   then on the GPU:  python profiles/prof_phases_coop.py [B] [T]
Thread 0 of every CTA adds clock64() deltas between phase boundaries (after the barrier that ends a phase) to a global
counter; the table is cycles per CTA and step.  Diagnostic only (the counters perturb the kernel), never bench values.
"""
import ctypes, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
LIB = ROOT / "profiles" / "_build" / "libdonk_prof.so"
if "--build" in sys.argv or not LIB.exists():
    LIB.parent.mkdir(exist_ok=True)
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-DDONK_PHASE_PROF"]
    units = sorted((ROOT / "donk_b200/csrc").glob("*.cu"))
    objs = [LIB.parent / (u.stem + ".prof.o") for u in units]
    with ThreadPoolExecutor(8) as pool:
        list(pool.map(lambda uo: subprocess.run(["nvcc"] + flags + ["-c", "-o", str(uo[1]), str(uo[0])], check=True), zip(units, objs)))
    subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB)] + [str(o) for o in objs], check=True)
    if "--build" in sys.argv:
        sys.exit(0)
import numpy as np, torch  # noqa: E402
from donk_b200 import _lib, batched, synthetic  # noqa: E402
_lib.LIB_PATH = LIB
nat = _lib.native()
nat.lib.donk_coop_phase_prof_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
dev = torch.device("cuda", 0)
args = [a for a in sys.argv[1:] if not a.startswith("--")]
B = int(args[0]) if len(args) > 0 else 128
T = int(args[1]) if len(args) > 1 else 200
dX, dU, N, E = 64, 16, 128, 1
ro = synthetic.rollout_batch_torch(B, T, dX, dU, N, seed=1, device=dev)
F, f, S, info = batched.fit_lr(ro["X"], ro["U"])
p = dX + dU
C, c, cc = synthetic.l2_costs(T, dX, dU)
x0m, x0S = batched.state_fit_from_rollouts(ro["X"])
il = batched.BatchedILQR(F, f, S, ro["prev_K"], ro["prev_k"], torch.as_tensor(C, device=dev).expand(B, T + 1, p, p).contiguous(),
                         torch.as_tensor(c, device=dev).expand(B, T + 1, p).contiguous(),
                         torch.as_tensor(cc, device=dev).expand(B, T + 1).contiguous(), x0m, x0S, prev_covar=ro["prev_covar"])
eta = torch.ones(B, E, dtype=torch.float64, device=dev)
il.step(eta); torch.cuda.synchronize()
nat.lib.donk_coop_phase_prof_read(None, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); il.step(eta); e1.record(); torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 32)()
nat.lib.donk_coop_phase_prof_read(out, 0)
names = {0: "B1  W=V[F|f] + barrier", 1: "B2  warp 0: Quu/q_u fragments", 2: "B2  warp 0: chol, L^-1, Quu^-1", 3: "B2  wait for the other warps",
         4: "B4  gains + barrier", 5: "B5  V update + wait + barrier", 6: "F1  K[S|mu] + barrier", 7: "F2  Suu, KL delta/trace + barrier",
         8: "F3  (F[S|mu])', cost + barrier", 9: "F4  F S F' + D + wait + barrier", 12: "bwd prologue", 13: "fwd prologue",
         14: "B1  step prologue (stage F, prefetch)", 15: "B1  warp 0 own tile", 16: "F1  step prologue (stage, prefetch, C loads)",
         17: "F1  warp 0 own tiles", 18: "F3  warp 0 cost part", 19: "F3  warp 0 own tile"}
names[0] = "B1  barrier wait"; names[6] = "F1  barrier wait"; names[8] = "F3  reduce + barrier wait"
ctas = B * E
tot = sum(out)
print(f"B={B} E={E} T={T}: {e0.elapsed_time(e1):.2f} ms (profiling build); cycles per CTA and step:")
for i in sorted(names):
    print(f"  {i:2d} {names[i]:40s} {out[i] / ctas / T:9.0f}  {100 * out[i] / tot:5.1f} %")
print(f"  total per CTA and step pair: {tot / ctas / T:.0f} cycles")
