#!/usr/bin/env python
"""Benchmark of the donk.ai trajectory-optimisation hot path on B200 (one JSON line on stdout).

Headline workload (BASELINE.json configs[2], the largest single-GPU configuration of the metric
"iLQR backward-pass solves/sec"): batched ILQR.step over 4096 conditions x 16 eta candidates per GPU,
T=100, dX=20, dU=6, fp64.  One *solve* = one (condition, eta) problem: backward Riccati recursion +
forward Gaussian pass + KL + expected cost (donk/traj_opt/lqg.py:55-75).  One *step* = one launch of
the fused kernel over all 65 536 solves of a rank.  The dynamics are produced on the device by the
fit_lr kernel from synthetic roll-outs.

Every other BASELINE.json configuration is measured in the same run and reported under "extra", each with
its own roofline object and (N=1) the CPU figure of the same work:
  config1_latency  configs[0]  fit_lr + ILQR.optimize of ONE condition (latency) and of a 4096 batch
  fit_lr_prior     configs[1]  fit_lr under a GMMPrior(4) NIW prior, T=100, dX=14, dU=7, N=50
  optimize         configs[2]  ILQR.optimize(kl_step=1) for 4096 conditions (one CUDA-graph launch)
  ilqr_strong      configs[2]  the headline step with 4096 conditions IN TOTAL (strong scaling)
  gmm_em           configs[3]  one EM iteration on a 10 M x 72 pool sharded over the GPUs (all-reduce inside)
  gmm_em_tf32x3    configs[3]  the same iteration at the stated fp32 tolerance on the tcgen05 tensor cores
  kmeans_pass      configs[3]  the Lloyd pass of the k-means initialisation over the same pool (HBM-bound)
  gps_iteration    configs[4]  BatchedTrajOpt.iteration (fit_lr + 16-cluster prior + ILQR.optimize + step_adjust)
                               at T=500, dX=64, dU=16, N=128, 128 conditions per GPU

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N>1: launched by torchrun, one rank per GPU, conditions sharded (weak scaling, no data-path
collective for ILQR; the EM extra all-reduces its statistics over NCCL and checks them against a single-rank pass).
--impl reference: the reference's own CPU implementation on all host cores, on a bounded sample of the same
workload: the UNMODIFIED reference from baseline/_ref (pure Python; `pip install --no-deps --target baseline/_ref`)
through donk.traj_opt.ILQR.step when it is there (kind "reference"), else the numpy restatement in oracle/ ("port").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CONFIGS = {
    1: dict(T=50, dX=6, dU=2, N=20),
    2: dict(T=100, dX=14, dU=7, N=50, K=4, B=4096),
    3: dict(T=100, dX=20, dU=6, N=64, B=4096, E=16),
    4: dict(n=10_000_000, d=72, K=16),
    5: dict(T=500, dX=64, dU=16, N=128, B=128, K=16, n_pool=2_000_000),
}
WORKLOAD = CONFIGS[3]


def eta_grid():
    import numpy as np

    return np.logspace(-2, 2, 16)  # the sample_surface form of SURVEY.md 8(d) config 3


# ------------------------------------------------------------------------------------------------
# Algorithmic work per unit (SURVEY.md section 8(d) "Algorithmic work per unit"; mul + add = 2 flops)
# ------------------------------------------------------------------------------------------------
def ilqr_flops_per_solve(T, dX, dU):
    """One ILQR.step: backward + forward + KL + expected cost."""
    p = dX + dU
    back = 2 * dX * dX * p + 2 * dX * p * p + 2 * dX * dX + 2 * dX * p + (7 / 3) * dU ** 3 + 2 * dU * dU * dX + 2 * dX * dX * dU
    fwd = 2 * dU * dX * dX + 2 * dU * dU * dX + 2 * dX * p * p + 2 * dX * dX * p
    klc = 2 * dU * dX + 4 * dU * dU + dU ** 3 / 3 + 4 * p * p
    return T * (back + fwd + klc)


def forward_flops(T, dX, dU):
    """One forward pass + expected cost (the three evaluations of step_adjust, lqg.py:328-334)."""
    p = dX + dU
    return T * (2 * dU * dX * dX + 2 * dU * dU * dX + 2 * dX * p * p + 2 * dX * dX * p + 4 * p * p)


def ilqr_bytes_per_launch(B, E, T, dX, dU):
    """Algorithmic HBM bytes of one launch: every input once per condition, policy scratch written by the
    backward pass and read back by the forward pass, results."""
    p = dX + dU
    per_cond = T * (dX * p + dX + dX * dX) + (T + 1) * (p * p + p + 1) + T * (p * p + p) \
        + T * (dU * dX + dU + dU * dU + 1) + dX + dX * dX
    pol = T * (dU * dX + dU + 2 * dU * dU)
    return 8 * (B * per_cond + B * E * (2 * pol - T * dU * dU + 3))


def fit_flops_per_fit(dX, dU, N, K=0):
    """One (condition, t) regression: N d (d+1) for the moments, 2 p^3 for the eigenvalue test and the factor,
    the solve and the Schur complement; with a K-cluster prior the responsibilities of the N samples
    (triangular products) and the mixture covariance."""
    p, d = dX + dU, 2 * dX + dU
    base = N * d * (d + 1) + 2 * p ** 3 + 2 * p * p * dX + 2 * dX * p * p + 2 * dX * dX * p
    prior = N * K * (d * (d + 1) + 2 * d) + 2 * K * d * d if K else 0
    return base + prior


def fit_bytes_per_fit(dX, dU, N):
    """Unique input bytes (x_{t+1} is the next step's x_t) + outputs."""
    p = dX + dU
    return 8 * (N * p + dX * p + dX + dX * dX)


def em_flops_per_row(d, K):
    return K * (2 * d * (d + 1) + 4 * d)


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (baseline/_ref) or its numpy restatement (oracle/), one process per
# host core, OMP_NUM_THREADS=1 set before numpy is imported in the workers (spawned, not forked)
# ------------------------------------------------------------------------------------------------
REF_DIR = ROOT / "baseline" / "_ref"


def cpu_kind():
    return "reference" if (REF_DIR / "donk" / "traj_opt" / "lqg.py").exists() else "port"


class _CpuArm:
    """The CPU implementation of the path behind one small interface (built inside the worker processes)."""

    def __init__(self):
        import numpy as np

        self.np = np
        self.kind = cpu_kind()
        if self.kind == "reference":
            if str(REF_DIR) not in sys.path:
                sys.path.insert(0, str(REF_DIR))
            import donk.traj_opt  # noqa: F401  (the unmodified reference)
        else:
            import oracle  # noqa: F401

    # -- prior ---------------------------------------------------------------------------------------
    def prior(self, params):
        if params is None:
            return None
        w, means, cov, n_pool = params
        if self.kind == "reference":
            from donk.dynamics.prior import GMMPrior
            from sklearn.mixture._gaussian_mixture import _compute_precision_cholesky

            pr = GMMPrior(len(w))
            pr.gmm.weights_, pr.gmm.means_, pr.gmm.covariances_ = w, means, cov
            pr.gmm.precisions_cholesky_ = _compute_precision_cholesky(cov, "full")
            pr.N = n_pool
            return pr
        from oracle.gmm import GMMParams, gmm_eval_prior, precision_cholesky

        prm = GMMParams(w, means, cov, precision_cholesky(cov))
        return lambda xux: gmm_eval_prior(prm, n_pool, xux)

    # -- pieces of the path --------------------------------------------------------------------------
    def fit_lr(self, X, U, prior=None):
        if self.kind == "reference":
            from donk.dynamics.linear_dynamics import fit_lr

            return fit_lr(X, U, prior=prior)
        import oracle

        return oracle.fit_lr(X, U, prior)

    def problem(self, c, dyn):
        """The ILQR object of one condition (constructor work included: extended_costs_kl)."""
        np = self.np
        if self.kind == "reference":
            from donk.costs import QuadraticCosts
            from donk.policy import LinearGaussianPolicy
            from donk.samples import StateDistribution
            from donk.traj_opt import ILQR

            return ILQR(dyn, LinearGaussianPolicy(c.prev_K, c.prev_k, c.prev_covar), QuadraticCosts(c.C, c.c, c.cc),
                        StateDistribution.fit(c.X[:, 0]))
        import oracle
        from oracle.lqg import ILQRProblem

        m, S0 = oracle.state_fit(c.X[:, 0])
        return ILQRProblem(dyn[0], dyn[1], dyn[2], c.C, c.c, c.cc, c.prev_K, c.prev_k, c.prev_covar,
                           np.linalg.inv(c.prev_covar), m, S0)

    def step(self, prob, eta):
        if self.kind == "reference":
            return prob.step(eta)
        import oracle

        return oracle.ilqr_step(prob, eta)

    def optimize(self, prob, kl_step):
        if self.kind == "reference":
            return prob.optimize(kl_step)
        import oracle

        return oracle.ilqr_optimize(prob, kl_step)[0]

    def gps_iterations(self, c, prior, n_iter):
        """TrajOptAlgorithm.iteration (traj_opt.py:35-96) n_iter times on the same roll-outs; seconds of the last."""
        if self.kind == "reference":
            from donk.costs import QuadraticCosts
            from donk.policy import LinearGaussianPolicy
            from donk.traj_opt.traj_opt import TrajOptAlgorithm

            algo = TrajOptAlgorithm(kl_step=1.0)
            costs = QuadraticCosts(c.C, c.c, c.cc)
            pol = LinearGaussianPolicy(c.prev_K, c.prev_k, c.prev_covar)
            for _ in range(n_iter):
                t0 = time.perf_counter()
                algo.iteration(c.X, c.U, costs, prev_pol=pol, prior=prior)
                dt = time.perf_counter() - t0
                pol = None
            return dt
        import oracle
        from oracle.lqg import ILQRProblem

        np = self.np
        prev, mult, pol = None, 1.0, (c.prev_K, c.prev_k, c.prev_covar)
        x0 = oracle.state_fit(c.X[:, 0])
        cst = (c.C, c.c, c.cc)
        for _ in range(n_iter):
            t0 = time.perf_counter()
            dyn = oracle.fit_lr(c.X, c.U, prior)
            prob = ILQRProblem(dyn[0], dyn[1], dyn[2], c.C, c.c, c.cc, pol[0], pol[1], pol[2], np.linalg.inv(pol[2]),
                               x0[0], x0[1])
            r = oracle.ilqr_optimize(prob, mult)[0]
            new = (r.K, r.k, r.covar)
            if prev is not None:
                mult = oracle.step_adjust(mult, dyn, new, x0, cst, prev, pol, x0, cst)[0]
            prev, pol = dyn, new
            dt = time.perf_counter() - t0
        return dt


_ARM = None


def _cpu_job(job):
    """One unit of CPU work; returns (units, seconds inside the timed region)."""
    global _ARM
    if _ARM is None:
        _ARM = _CpuArm()
    arm = _ARM
    from donk_b200 import synthetic

    kind, cfg, seed, extra = job
    w = CONFIGS[cfg]
    T, dX, dU, N = w["T"], w["dX"], w["dU"], w["N"]
    c = synthetic.condition(cfg, seed, T, dX, dU, N)
    if kind == "step":  # ILQR.step for every eta of the grid; problem set-up is outside the metric
        prob = arm.problem(c, arm.fit_lr(c.X, c.U))
        t0 = time.perf_counter()
        for eta in extra:
            arm.step(prob, float(eta))
        return len(extra), time.perf_counter() - t0
    if kind == "optimize":
        prob = arm.problem(c, arm.fit_lr(c.X, c.U))
        t0 = time.perf_counter()
        arm.optimize(prob, 1.0)
        return 1, time.perf_counter() - t0
    if kind == "fit_optimize":  # the whole per-condition path of config 1
        t0 = time.perf_counter()
        arm.optimize(arm.problem(c, arm.fit_lr(c.X, c.U)), 1.0)
        return 1, time.perf_counter() - t0
    if kind == "fit_prior":
        prior = arm.prior(extra)
        t0 = time.perf_counter()
        arm.fit_lr(c.X, c.U, prior)
        return T, time.perf_counter() - t0
    if kind == "gps":
        prior = arm.prior(extra)
        return 1, arm.gps_iterations(c, prior, 2)
    raise ValueError(kind)


class CpuPool:
    """Worker processes for the CPU arm: spawned with the BLAS/OpenMP pools pinned to one thread each BEFORE numpy is
    imported there, one process per host core."""

    def __init__(self, cores=None):
        import multiprocessing as mp

        self.cores = cores or os.cpu_count() or 1
        saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
        for k in saved:
            os.environ[k] = "1"
        try:
            self.pool = mp.get_context("spawn").Pool(self.cores)
            self.pool.map(_cpu_warm, range(self.cores))  # import numpy/scipy/the arm in every worker
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v

    def run(self, jobs):
        """-> (units, busy seconds summed over workers).  Throughput of the pool = units * cores / busy: every core
        runs timed regions back to back (set-up between them is outside the metric, as on the GPU side)."""
        out = self.pool.map(_cpu_job, jobs, chunksize=max(1, len(jobs) // self.cores))
        return sum(u for u, _ in out), sum(s for _, s in out)

    def close(self):
        self.pool.close()
        self.pool.join()


def _cpu_warm(_):
    global _ARM
    if _ARM is None:
        _ARM = _CpuArm()
    return os.getpid()


def cpu_ilqr_samples(pool, conds_per_core, n_samples, seed0=0):
    """n_samples bounded samples of the headline workload; returns the list of solves/s and the last busy seconds."""
    etas = list(eta_grid())
    rates, busy = [], 0.0
    for s in range(n_samples):
        jobs = [("step", 3, seed0 + s * 100000 + i, etas) for i in range(pool.cores * conds_per_core)]
        units, busy = pool.run(jobs)
        rates.append(units * pool.cores / busy)
    return rates, busy


def run_reference(args):
    import numpy as np

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pool = CpuPool()
    try:
        for _ in range(args.warmup):
            cpu_ilqr_samples(pool, 1, 1)
        rates, busy_total = [], 0.0
        t0 = time.perf_counter()
        for s in range(args.steps):
            r, busy = cpu_ilqr_samples(pool, 4, 1, seed0=s * 1000)
            rates += r
            busy_total += busy
        wall = time.perf_counter() - t0
    finally:
        pool.close()
    value = float(np.median(rates))
    w = WORKLOAD
    line = {
        "impl": "reference", "metric": "ilqr_solves_per_sec", "value": value, "unit": "solves/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": pool.cores, "kind": cpu_kind(),
                         "sample": f"per step {4 * pool.cores} conditions x 16 eta (T={w['T']},dX={w['dX']},dU={w['dU']}) through "
                                   + ("donk.traj_opt.ILQR.step of the unmodified reference (baseline/_ref)"
                                      if cpu_kind() == "reference" else "oracle.ilqr_step (numpy restatement)")
                                   + f", one single-threaded process per core; median of {len(rates)} samples; "
                                     f"rate = solves x cores / summed busy seconds",
                         "samples": rates},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(n_gpus):
    w = WORKLOAD
    return {
        "workload": f"batched ILQR.step: {w['B']} conditions x {w['E']} eta per GPU, T={w['T']}, dX={w['dX']}, "
                    f"dU={w['dU']} (BASELINE.json configs[2]); dynamics from fit_lr on N={w['N']} synthetic roll-outs",
        "conditions_per_gpu": w["B"], "eta_candidates": w["E"], "T": w["T"], "dX": w["dX"], "dU": w["dU"],
        "solve": "one (condition, eta): backward + forward + KL + expected cost",
        "parallelism": f"conditions sharded over {n_gpus} GPU(s), no collective",
        "l2": f"inputs per launch ({ilqr_bytes_per_launch(w['B'], w['E'], w['T'], w['dX'], w['dU']) / 1e9:.1f} GB "
              f"algorithmic traffic) exceed the 126 MB L2; no flush between steps",
    }


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the GPU work of this script runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        import numpy as np

        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 6]
        os.unlink(self.f.name)
        sm, pw, mx, reasons = [], [], None, set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                pw.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": mx, "reasons": ["no samples"]}
        load = [s for s, w_ in zip(sm, pw) if w_ >= 0.6 * max(pw)]  # samples taken under load
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "samples_under_load": len(load), "power_w_max": max(pw)}


def bind_to_gpu_numa_node(local_rank):
    """Run this rank (and allocate its pinned host buffers) on the CPUs next to its GPU: cudaHostAlloc places pages on
    the calling thread's NUMA node, and H2D copies from the far socket share the inter-socket link with every other
    rank.  Returns a description for the JSON line."""
    try:
        import torch

        props = torch.cuda.get_device_properties(local_rank)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        cpulist = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
        node = open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            return {"numa_node": node, "cpus_bound": len(use), "of_allowed": len(allowed)}
        return {"numa_node": node, "cpus_bound": None, "note": "GPU-local CPUs cover the whole allowed set (or none of it)"}
    except Exception as ex:  # informational only
        return {"error": repr(ex)[:120]}


def run_ours(args):
    import ctypes

    import numpy as np
    import torch
    import torch.distributed as dist

    from donk_b200 import _lib, batched, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dry = args.dry_run_emu  # tests only: tiny sizes on the host emulation of the kernels, to check this script
    ETA_GRID = eta_grid()
    if dry:
        sys.path.insert(0, str(ROOT / "tests"))
        from emu.build_emu import backend as emu_backend

        _lib.install_test_backend(emu_backend())
        CONFIGS[1].update(T=5, dX=6, dU=2, N=12)
        CONFIGS[2].update(T=3, dX=14, dU=7, N=30, K=2, B=2)
        CONFIGS[3].update(T=6, dX=4, dU=2, N=8, B=5, E=16)
        CONFIGS[4].update(n=600, d=10, K=3)
        CONFIGS[5].update(T=3, dX=16, dU=8, N=40, B=2, K=2, n_pool=500)
        dev = torch.device("cpu")
        numa = None
        if world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("gloo")
    else:
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        numa = bind_to_gpu_numa_node(local) if world > 1 else None
        if world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # NCCL's own log lines (version banner, NCCL_DEBUG output) go to stderr: stdout carries the one JSON line
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            dist.init_process_group("nccl", device_id=dev)
    nat = _lib.native()
    w = WORKLOAD
    B, E, T, dX, dU, N = w["B"], w["E"], w["T"], w["dX"], w["dU"], w["N"]
    p = dX + dU
    f64 = dict(dtype=torch.float64, device=dev)

    def sync():
        if not dry:
            torch.cuda.synchronize()

    def barrier():
        sync()
        if world > 1:
            dist.barrier()
        sync()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], **f64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        if dry:
            t0 = time.perf_counter()
            for _ in range(steps):
                fn()
            return max_over_ranks(1e3 * (time.perf_counter() - t0))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()  # torch's current stream == the stream the library launches on
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    def timed_median(fn, reps, warmup, inner=1):
        """Per-call time as the median over `reps` samples of `inner` back-to-back calls each (extras: a single allocator
        hiccup moves a short mean by 10-15 %; back-to-back calls keep multi-launch sequences pipelined)."""
        timed(fn, 0, warmup)
        return float(np.median([timed(fn, inner, 0) / inner for _ in range(reps)]))

    def free():
        if not dry:
            torch.cuda.empty_cache()

    def costs_for(Bc, Tc, dXc, dUc):
        """quadratic_cost_approximation_l2(t = 0, w = [1]*dX + [1e-2]*dU) for every condition, built on the device."""
        pc = dXc + dUc
        wv = torch.cat([torch.ones(dXc, **f64), 1e-2 * torch.ones(dUc, **f64)]).expand(Bc, Tc + 1, pc).contiguous()
        return batched.quadratic_cost_approximation_l2(torch.zeros(Bc, Tc + 1, pc, **f64), wv)

    clocks = ClockSampler(local)
    holder = {}
    extra = {}

    # ---- FP64 pipe probe: the roofline denominator of the FP64-bound kernels ------------------------------
    scratch = torch.zeros(8, **f64)
    peaks = {"dfma": 0.0, "dmma": 0.0}
    for mode, name in ((0, "dfma"), (1, "dmma")):
        flops = ctypes.c_double(0)
        for it in range(0 if dry else 4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            nat.call("fp64_probe", mode, 148 * 8, 20000, scratch, ctypes.byref(flops))
            e1.record()
            torch.cuda.synchronize()
            if it:
                peaks[name] = max(peaks[name], flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    fp64_peak = max(peaks.values()) or None
    hbm_peak = json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    peak_src = (f"in-run FP64 probe on this GPU (dfma {peaks['dfma']:.1f}, dmma {peaks['dmma']:.1f} TFLOP/s); "
                f"MEASURED_PEAKS.json holds no FP64 figure")

    def roof(flops, nbytes, ms, kernel):
        """Roofline object of a launch (or launch sequence) that does `flops` / moves `nbytes` per rank in `ms`."""
        tf = flops / (ms * 1e-3) / 1e12
        gbs = nbytes / (ms * 1e-3) / 1e9
        fp_frac = tf / fp64_peak if fp64_peak else None
        hbm_frac = gbs / hbm_peak
        bound = "fp64" if (fp_frac or 0) >= hbm_frac else "hbm"
        return {"bound": bound, "achieved": tf if bound == "fp64" else gbs, "peak": fp64_peak if bound == "fp64" else hbm_peak,
                "unit": "TFLOP/s" if bound == "fp64" else "GB/s", "frac": fp_frac if bound == "fp64" else hbm_frac,
                "traffic": None, "kernel": kernel, "fp64_tflops": tf, "fp64_frac": fp_frac, "hbm_gbs": gbs,
                "hbm_frac": hbm_frac, "algorithmic_flops": flops, "algorithmic_bytes": nbytes}

    # ---- synthetic inputs on the device (untimed) ------------------------------------------------
    ro = synthetic.rollout_batch_torch(B, T, dX, dU, N, seed=3000 + rank, device=dev)
    C_np, c_np, cc_np = synthetic.l2_costs(T, dX, dU)
    Cb, cb, ccb = costs_for(B, T, dX, dU)
    assert np.array_equal(Cb[0].cpu().numpy(), C_np)
    eta = torch.as_tensor(ETA_GRID, device=dev).expand(B, E).contiguous()

    # ---- extra: fit_lr (also produces the dynamics the headline consumes) ----------------------------
    def fit_once():
        holder["fit"] = batched.fit_lr(ro["X"], ro["U"])

    nfit = max(2, args.steps // 2)
    fit_ms = timed_median(fit_once, nfit, 2)
    F, f, dyn_covar, info = holder["fit"]
    assert int(info.sum().item()) == 0, "fit_lr: matrix not positive definite"
    extra["fit_lr"] = {"timestep_fits_per_sec": B * T * world / (fit_ms * 1e-3), "ms_per_launch": fit_ms,
                       "config": f"B={B} conditions/GPU, T={T}, dX={dX}, dU={dU}, N={N}, no prior",
                       "roofline": roof(fit_flops_per_fit(dX, dU, N) * B * T, fit_bytes_per_fit(dX, dU, N) * B * T, fit_ms,
                                        f"fit_warp_kernel<{dX},{dU}>")}
    x0m, x0S = batched.state_fit_from_rollouts(ro["X"])
    il = batched.BatchedILQR(F, f, dyn_covar, ro["prev_K"], ro["prev_k"], Cb, cb, ccb, x0m, x0S,
                             prev_covar=ro["prev_covar"])
    del ro["X"], ro["U"]
    free()

    # ---- headline: device-resident ILQR.step -----------------------------------------------------------
    def step_once():
        holder["res"] = il.step(eta)

    step_once()
    sync()
    assert int(holder["res"].info.sum().item()) == 0, "ILQR.step: Quu not positive definite"
    timed(step_once, 0, args.warmup)
    l0 = nat.launch_count()
    ms = timed(step_once, args.steps, 0)
    launches = nat.launch_count() - l0
    ms_per_step = ms / args.steps
    solves_per_step = B * E * world
    value = solves_per_step / (ms_per_step * 1e-3)

    # ---- parity spot check against the oracle (2 conditions x 3 eta) -----------------------------------
    parity = None
    if rank == 0:
        import oracle
        from oracle.lqg import ILQRProblem

        res = holder["res"]
        worst = 0.0
        for b in (0, B - 1):
            g = lambda t: t[b].cpu().numpy()  # noqa: E731
            prob = ILQRProblem(g(F), g(f), g(dyn_covar), C_np, c_np, cc_np, g(il.prev_K), g(il.prev_k), g(il.prev_covar),
                               g(il.prev_inv_covar), g(x0m), g(x0S))
            for e in (0, 7, 15):
                ref = oracle.ilqr_step(prob, float(ETA_GRID[e]))
                worst = max(worst, np.max(np.abs(res.K[b, e].cpu().numpy() - ref.K)) / np.max(np.abs(ref.K)),
                            abs(float(res.kl_div[b, e]) - ref.kl_div) / max(1.0, abs(ref.kl_div)),
                            abs(float(res.expected_costs[b, e]) - ref.expected_costs) / abs(ref.expected_costs))
        parity = worst
    dev_res = (holder["res"].kl_div.cpu(), holder["res"].expected_costs.cpu())

    # ---- extra: ILQR.optimize (config 3), one CUDA-graph launch per call ---------------------------------
    try:
        def opt_once():
            holder["opt"] = il.optimize(1.0, raise_on_failure=False)

        n_opt = 2 if not dry else 1
        opt_ms = timed(opt_once, n_opt, 1) / n_opt
        _, status, n_steps = holder["opt"]
        extra["optimize"] = {
            "optimisations_per_sec": B * world / (opt_ms * 1e-3), "ms_per_call": opt_ms, "batched_steps": n_steps,
            "status_counts": torch.bincount(status, minlength=5).tolist(),
            "config": f"ILQR.optimize(kl_step=1) for {B} conditions/GPU (T={T}, dX={dX}, dU={dU}); one graph launch, "
                      f"{n_steps} step kernels with E=1 + controller kernels, one 12-byte device->host read",
            "roofline": roof(ilqr_flops_per_solve(T, dX, dU) * B * n_steps, ilqr_bytes_per_launch(B, 1, T, dX, dU) * n_steps,
                             opt_ms, f"ilqr_warp_kernel<{dX},{dU}> at E=1 inside the optimize graph")}
    except Exception as ex:
        extra["optimize"] = {"error": repr(ex)[:300]}

    # ---- e2e: host buffers in, results out, through the public batched API ------------------------------
    # (chunked double-buffered pipeline: the H2D copy of chunk c+1 overlaps the solve of chunk c)
    pin = (lambda t: t) if dry else (lambda t: t.pin_memory())
    host = {k: pin(v.cpu()) for k, v in dict(F=F, f=f, dyn_covar=dyn_covar, prev_K=il.prev_K, prev_k=il.prev_k,
                                             prev_covar=il.prev_covar, C=Cb, c=cb, cc=ccb, x0_mean=x0m,
                                             x0_covar=x0S).items()}
    eta_host = pin(eta.cpu())
    holder.pop("res", None)
    holder.pop("opt", None)
    del il
    free()
    pipe = batched.HostPipelinedILQR(host, E, chunks=args.e2e_chunks)
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes

    def e2e_once():
        holder["e2e"] = pipe.step(eta_host)

    e2e_steps = max(1, min(args.steps, 5))
    e2e_ms = timed(e2e_once, e2e_steps, 1) / e2e_steps
    e2e_value = solves_per_step / (e2e_ms * 1e-3)
    kl_h, cost_h, info_h = holder["e2e"]
    assert int(info_h.sum()) == 0
    assert torch.allclose(kl_h, dev_res[0], rtol=1e-12, atol=1e-12) and torch.allclose(cost_h, dev_res[1], rtol=1e-12)
    clk = clocks.stop()
    del pipe
    holder.clear()
    free()

    flops_per_launch = ilqr_flops_per_solve(T, dX, dU) * B * E
    achieved_tf = flops_per_launch / (ms_per_step * 1e-3) / 1e12
    bytes_per_launch = ilqr_bytes_per_launch(B, E, T, dX, dU)
    traffic = None  # DRAM bytes per launch from the committed ncu --set full capture (scaled by solves per launch)
    tp = ROOT / "profiles" / "ilqr_traffic.json"
    if tp.exists() and (dX, dU, T) == (20, 6, 100):
        traffic = json.load(open(tp))["dram_bytes_per_solve"] * B * E

    # ---- extra: the headline step with 4096 conditions IN TOTAL (strong scaling) ---------------------------
    try:
        Bs = max(1, B // world)
        sl = slice(0, Bs)
        il_s = batched.BatchedILQR(*(host[k][sl].to(dev) for k in ("F", "f", "dyn_covar", "prev_K", "prev_k", "C", "c", "cc",
                                                                    "x0_mean", "x0_covar")),
                                   prev_covar=host["prev_covar"][sl].to(dev))
        eta_s = eta[:Bs].contiguous()
        s_ms = timed(lambda: il_s.step(eta_s), max(3, args.steps // 2), 2) / max(3, args.steps // 2)
        extra["ilqr_strong"] = {"solves_per_sec": Bs * world * E / (s_ms * 1e-3), "ms_per_step": s_ms, "scaling": "strong",
                                "config": f"{Bs * world} conditions in total ({Bs}/GPU) x {E} eta, T={T}, dX={dX}, dU={dU}",
                                "roofline": roof(ilqr_flops_per_solve(T, dX, dU) * Bs * E,
                                                 ilqr_bytes_per_launch(Bs, E, T, dX, dU), s_ms, f"ilqr_warp_kernel<{dX},{dU}>")}
        del il_s
    except Exception as ex:
        extra["ilqr_strong"] = {"error": repr(ex)[:300]}
    del host
    free()

    # ---- extra: config 1 (T=50, dX=6, dU=2, N=20): single-condition latency and a 4096 batch ------------------
    try:
        c1 = CONFIGS[1]
        out1 = {}
        for Bc in ((1, 4096) if not dry else (1, 3)):
            r1 = synthetic.rollout_batch_torch(Bc, c1["T"], c1["dX"], c1["dU"], c1["N"], seed=1000 + rank, device=dev)
            C1, c1v, cc1 = costs_for(Bc, c1["T"], c1["dX"], c1["dU"])

            def whole():
                Fq, fq, Sq, _ = batched.fit_lr(r1["X"], r1["U"])
                m0, S0 = batched.state_fit_from_rollouts(r1["X"])
                ilq = batched.BatchedILQR(Fq, fq, Sq, r1["prev_K"], r1["prev_k"], C1, c1v, cc1, m0, S0,
                                          prev_covar=r1["prev_covar"])
                holder["c1"] = ilq.optimize(1.0, raise_on_failure=False)

            n1 = 5 if not dry else 1
            ms1 = timed(whole, n1, 2 if not dry else 0) / n1
            n_steps1 = holder["c1"][2]
            fl = (fit_flops_per_fit(c1["dX"], c1["dU"], c1["N"]) * c1["T"]
                  + ilqr_flops_per_solve(c1["T"], c1["dX"], c1["dU"]) * n_steps1) * Bc
            by = (fit_bytes_per_fit(c1["dX"], c1["dU"], c1["N"]) * c1["T"]
                  + ilqr_bytes_per_launch(1, 1, c1["T"], c1["dX"], c1["dU"]) * n_steps1) * Bc
            out1[f"B{Bc}"] = {"ms_fit_lr_plus_optimize": ms1, "conditions_per_sec": Bc * world / (ms1 * 1e-3),
                              "batched_steps": n_steps1,
                              "roofline": roof(fl, by, ms1, "fit_warp_kernel<6,2> + ilqr_warp_kernel<6,2> (optimize graph)")}
            del r1
        out1["config"] = (f"state fit + fit_lr + ILQR construction + ILQR.optimize(kl_step=1), T={c1['T']}, dX={c1['dX']}, "
                          f"dU={c1['dU']}, N={c1['N']} (BASELINE.json configs[0]); B=1 is latency-bound "
                          f"({c1['T']} dependent time steps x steps), its roofline fraction is reported for completeness")
        extra["config1_latency"] = out1
        holder.clear()
    except Exception as ex:
        extra["config1_latency"] = {"error": repr(ex)[:300]}
    free()

    # ---- extra: config 2 (T=100, dX=14, dU=7, N=50) fit_lr under a GMMPrior(4) prior ---------------------------
    prior2 = None
    try:
        c2 = CONFIGS[2]
        B2 = c2["B"]
        r2 = synthetic.rollout_batch_torch(B2, c2["T"], c2["dX"], c2["dU"], c2["N"], seed=2000 + rank, device=dev)
        prior2 = synthetic.gmm_prior(c2["dX"], c2["dU"], c2["K"], seed=21) + (20000 + c2["N"] * c2["T"],)
        gm2 = batched.DeviceGMM(c2["K"]).set_params(prior2[0], prior2[1], covariances=prior2[2],
                                                    precisions_cholesky=synthetic.precision_cholesky(prior2[2]))
        gm2.N = prior2[3]

        def fit2():
            holder["f2"] = batched.fit_lr(r2["X"], r2["U"], gm2)

        ms2 = timed_median(fit2, 5, 1)
        assert int(holder["f2"][3].sum().item()) == 0
        if rank == 0:  # parity of condition 0 against the oracle
            import oracle
            from oracle.gmm import GMMParams, gmm_eval_prior, precision_cholesky

            prm = GMMParams(prior2[0], prior2[1], prior2[2], precision_cholesky(prior2[2]))
            Tq = min(c2["T"], 3)
            Fo, fo, co = oracle.fit_lr(r2["X"][0, :, :Tq + 1].cpu().numpy(), r2["U"][0, :, :Tq].cpu().numpy(),
                                       lambda xux: gmm_eval_prior(prm, prior2[3], xux))
            par2 = max(float(np.max(np.abs(holder["f2"][0][0, :Tq].cpu().numpy() - Fo)) / np.max(np.abs(Fo))),
                       float(np.max(np.abs(holder["f2"][2][0, :Tq].cpu().numpy() - co)) / np.max(np.abs(co))))
        else:
            par2 = None
        extra["fit_lr_prior"] = {
            "timestep_fits_per_sec": B2 * c2["T"] * world / (ms2 * 1e-3), "ms_per_launch": ms2,
            "parity_max_rel_err_vs_oracle": par2,
            "config": f"B={B2} conditions/GPU, T={c2['T']}, dX={c2['dX']}, dU={c2['dU']}, N={c2['N']}, GMMPrior({c2['K']}) "
                      f"NIW prior evaluated inside the fit (BASELINE.json configs[1])",
            "roofline": roof(fit_flops_per_fit(c2["dX"], c2["dU"], c2["N"], c2["K"]) * B2 * c2["T"],
                             fit_bytes_per_fit(c2["dX"], c2["dU"], c2["N"]) * B2 * c2["T"], ms2,
                             f"fit_warp_kernel<{c2['dX']},{c2['dU']}> with prior")}
        del r2, gm2
        holder.clear()
    except Exception as ex:
        extra["fit_lr_prior"] = {"error": repr(ex)[:300]}
    free()

    # ---- extra: GMM EM iteration on a sharded transition pool (BASELINE.json configs[3]) --------------------
    em_strong = None
    try:
        c4 = CONFIGS[4]
        n_rows, d, K = (c4["n"] if dry else args.em_rows) // world, c4["d"], c4["K"]
        Xp, w0, m0, p0 = synthetic.gmm_pool_torch(n_rows, d, K, seed=4000 + rank, device=dev)
        gm = batched.DeviceGMM(K, tol=0.0, max_iter=1).set_params(w0, m0, precisions=p0)
        ws = torch.empty(nat.size("gmm_workspace_bytes", n_rows, d, K) // 8 + 1, **f64)
        stats = torch.empty(nat.size("gmm_stats_size", d, K), **f64)
        gm._lb = torch.zeros(1, **f64)
        gm._info = torch.zeros(K, dtype=torch.int32, device=dev)
        em_ms = timed_median(lambda: gm.em_iteration(Xp, float(n_rows * world), ws, stats), 3, 2, inner=2)
        extra["gmm_em"] = {"transitions_per_sec_per_iteration": n_rows * world / (em_ms * 1e-3), "ms_per_iteration": em_ms,
                           "config": f"{n_rows * world} rows total ({n_rows}/GPU), d={d}, K={K}, one E+M iteration "
                                     f"incl. all-reduce and M finalise",
                           "roofline": roof(em_flops_per_row(d, K) * n_rows, n_rows * d * 8, em_ms,
                                            "gmm_e_mma_kernel + gmm_m_mma_kernel (mma.sync.m8n8k4.f64)")}
        em_strong = {"rows_total": n_rows * world, "n_gpus": world, "ms_per_iteration": em_ms, "scaling": "strong",
                     "collective": "all-reduce of the packed statistics (NCCL), once per iteration"}
        # the Lloyd pass that initialises the first GMMPrior.update over the same pool (the one HBM-bound kernel of the path)
        try:
            c0 = Xp[:K].contiguous()
            kws = torch.empty(nat.size("kmeans_workspace_bytes", d, K) // 8 + 1, **f64)
            ksums = torch.empty(K * d + K, **f64)
            km_ms = timed(lambda: batched.DeviceKMeans._pass(Xp, c0, labels=True, sums=ksums, ws=kws), 5, 2) / 5
            kl_ms = timed(lambda: batched.DeviceKMeans._pass(Xp, c0, labels=True), 5, 2) / 5
            kbytes = n_rows * d * 8
            extra["kmeans_pass"] = {
                "ms_per_lloyd_pass": km_ms, "ms_labels_only": kl_ms, "rows_per_sec": n_rows * world / (km_ms * 1e-3),
                "config": f"{n_rows * world} rows total ({n_rows}/GPU), d={d}, K={K}: labels + per-cluster sums in one read of the pool "
                          f"(kmeans_pass_mma_kernel: bulk-copied tiles, distances and one-hot sums on DMMA)",
                "roofline": {"bound": "hbm", "achieved": kbytes / (km_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": kbytes / (km_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                             "kernel": "kmeans_pass_mma_kernel", "algorithmic_bytes": kbytes,
                             "labels_only_gbs": kbytes / (kl_ms * 1e-3) / 1e9,
                             "labels_only_frac": kbytes / (kl_ms * 1e-3) / 1e9 / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs"}}
            del kws, ksums, c0
        except Exception as ex:
            extra["kmeans_pass"] = {"error": repr(ex)[:300]}
        # the same iteration at the stated fp32 tolerance: E pass and scatter statistics on tcgen05 (3xTF32 split products,
        # TMEM accumulators, fp64 partial sums / all-reduce / finalisation); fp64 stays the default and the parity path
        try:
            if dry:
                raise RuntimeError("host emulation has no tensor-core path")
            ref64 = batched.DeviceGMM(K, tol=0.0, max_iter=1).set_params(w0, m0, precisions=p0)
            ref64._lb, ref64._info = torch.zeros(1, **f64), torch.zeros(K, dtype=torch.int32, device=dev)
            ref64.em_iteration(Xp, float(n_rows * world), ws, stats)
            gt = batched.DeviceGMM(K, tol=0.0, max_iter=1, precision="tf32x3").set_params(w0, m0, precisions=p0)
            gt._lb, gt._info = torch.zeros(1, **f64), torch.zeros(K, dtype=torch.int32, device=dev)
            prep_ms = timed(lambda: setattr(gt, "_tc", gt._tc_setup(Xp, m_pass=True)), 1, 0)
            gt.em_iteration(Xp, float(n_rows * world), ws, stats)
            rel = lambda a, b: float(((a - b).abs().max() / b.abs().max()).item())  # noqa: E731
            tc_err = {"means": rel(gt.means, ref64.means), "covariances": rel(gt.covariances, ref64.covariances),
                      "weights": rel(gt.weights, ref64.weights), "lower_bound": rel(gt._lb, ref64._lb)}
            gt.set_params(w0, m0, precisions=p0)
            tc_ms = timed_median(lambda: gt.em_iteration(Xp, float(n_rows * world), ws, stats), 3, 2, inner=4)
            peaks_js = json.load(open(ROOT / "MEASURED_PEAKS.json")) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
            tf32_peak = peaks_js.get("bf16_tflops", 2250.0) / 2  # TF32 dense = half the bf16 rate; burst figure (timed alone)
            algo = em_flops_per_row(d, K) * n_rows
            d8, d16 = -(-d // 8) * 8, -(-(d + 1) // 16) * 16
            issued = 3 * 2 * n_rows * K * (d8 * d8 + d * d16)      # three TF32 products per fp32-accurate product, padded tiles
            # HBM: E reads the two TF32 planes (8d B/row) and writes fp64 responsibilities; M reads its planes and them
            tc_bytes = n_rows * (2 * 8 * d + 2 * 8 * K)
            extra["gmm_em_tf32x3"] = {
                "transitions_per_sec_per_iteration": n_rows * world / (tc_ms * 1e-3), "ms_per_iteration": tc_ms,
                "speedup_vs_fp64_iteration": em_ms / tc_ms, "split_copy_ms_once_per_fit": prep_ms,
                "max_rel_err_vs_fp64_after_one_iteration": tc_err,
                "stated_tolerance": "means 1e-4, covariances 1e-3, lower bound 1e-5 relative to the fp64 fit, responsibilities 2e-4 "
                                    "absolute (tests/test_fit_gmm_kernels.py); fp64 is the default and the 1e-8 parity path",
                "config": f"DeviceGMM(precision='tf32x3'): {n_rows * world} rows total, d={d}, K={K}; E pass "
                          f"(gmm_tc_estep_kernel) and scatter statistics (gmm_tc_mstep_kernel) on tcgen05.mma kind::tf32, "
                          f"operands by cp.async.bulk, accumulators in TMEM",
                # achieved = ALGORITHMIC flops (the fp64 iteration's count) per second; the tensor pipe itself executes
                # `issued`: x3 for the split products, full instead of triangular P_k, tiles padded to the MMA shape
                "roofline": {"bound": "tensor", "achieved": algo / (tc_ms * 1e-3) / 1e12, "peak": tf32_peak,
                             "unit": "TFLOP/s", "frac": algo / (tc_ms * 1e-3) / 1e12 / tf32_peak, "traffic": None,
                             "issued_tf32_tflops": issued / (tc_ms * 1e-3) / 1e12,
                             "issued_frac": issued / (tc_ms * 1e-3) / 1e12 / tf32_peak,
                             "kernel": "gmm_tc_estep_kernel + gmm_tc_mstep_kernel (tcgen05.mma.kind::tf32, 3 products per fp32 product)",
                             "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (TF32 dense rate is half of bf16), burst",
                             "issued_tf32_flops": issued, "algorithmic_flops": algo,
                             "algorithmic_tflops": algo / (tc_ms * 1e-3) / 1e12,
                             "hbm_gbs": tc_bytes / (tc_ms * 1e-3) / 1e9, "hbm_frac": tc_bytes / (tc_ms * 1e-3) / 1e9 / hbm_peak,
                             "algorithmic_bytes": tc_bytes}}
            del gt, ref64
        except Exception as ex:
            extra["gmm_em_tf32x3"] = {"error": repr(ex)[:300]}
        del Xp, ws, gm
        free()
        # multi-rank parity: ONE pool, rows dealt to the ranks; the all-reduced result must equal rank 0's single pass
        if world > 1:
            n_par = 1_000_000 if not dry else 400
            Xall, w0, m0, p0 = synthetic.gmm_pool_torch(n_par, d, K, seed=4100, device=dev)  # same rows on every rank
            mine = Xall[rank::world].contiguous()

            def one_iter(X, group_on):
                g = batched.DeviceGMM(K, tol=0.0, max_iter=1).set_params(w0, m0, precisions=p0)
                g._mode = 2  # the tensor-path kernels the timed run uses
                wsx = torch.empty(nat.size("gmm_workspace_bytes", X.shape[0], d, K) // 8 + 1, **f64)
                st = torch.empty(nat.size("gmm_stats_size", d, K), **f64)
                g._lb = torch.zeros(1, **f64)
                g._info = torch.zeros(K, dtype=torch.int32, device=dev)
                if group_on:
                    g.em_iteration(X, float(n_par), wsx, st)
                else:
                    saved = batched.DeviceGMM.__dict__["_allreduce"]  # the staticmethod object itself, not the unwrapped function
                    batched.DeviceGMM._allreduce = staticmethod(lambda t, group: None)
                    try:
                        g.em_iteration(X, float(n_par), wsx, st)
                    finally:
                        batched.DeviceGMM._allreduce = saved
                return g

            gs = one_iter(mine, True)
            g1 = one_iter(Xall, False)
            rel = lambda a, b: float(((a - b).abs().max() / b.abs().max()).item())  # noqa: E731
            err = max(rel(gs.means, g1.means), rel(gs.covariances, g1.covariances), rel(gs.weights, g1.weights),
                      rel(gs._lb, g1._lb))
            extra["gmm_em"]["multi_rank_parity"] = {
                "max_rel_err": max_over_ranks(err), "bound": 1e-11, "rows": n_par,
                "what": f"means/covariances/weights/lower bound after one EM iteration with the rows dealt to {world} ranks "
                        f"(NCCL all-reduce) vs a single-rank pass over all rows, worst rank"}
            assert extra["gmm_em"]["multi_rank_parity"]["max_rel_err"] < 1e-11, extra["gmm_em"]["multi_rank_parity"]
            del Xall, mine, gs, g1
    except AssertionError:
        raise
    except Exception as ex:  # the extra must not take the headline down
        extra["gmm_em"] = {"error": repr(ex)[:300]}
    free()

    # ---- extra: config 5, a full GPS iteration per condition (fit_lr + prior + ILQR.optimize + step_adjust) -----------
    prior5 = None
    try:
        c5 = CONFIGS[5]
        B5, T5, dX5, dU5, N5, K5 = c5["B"], c5["T"], c5["dX"], c5["dU"], c5["N"], c5["K"]
        d5 = 2 * dX5 + dU5
        r5 = synthetic.rollout_batch_torch(B5, T5, dX5, dU5, N5, seed=5000 + rank, device=dev)
        C5, c5v, cc5 = costs_for(B5, T5, dX5, dU5)
        prior5 = synthetic.gmm_prior(dX5, dU5, K5, seed=51) + (c5["n_pool"],)
        gm5 = batched.DeviceGMM(K5).set_params(prior5[0], prior5[1], covariances=prior5[2],
                                               precisions_cholesky=synthetic.precision_cholesky(prior5[2]))
        gm5.N = prior5[3]
        algo = batched.BatchedTrajOpt(kl_step=1.0)
        algo.iteration(r5["X"], r5["U"], C5, c5v, cc5, r5["prev_K"], r5["prev_k"], r5["prev_covar"], prior=gm5)

        def gps_once():
            holder["g5"] = algo.iteration(r5["X"], r5["U"], C5, c5v, cc5, prior=gm5)

        n5 = 3 if not dry else 1
        if not dry:
            gps_once()  # second warm-up: the first iteration that has a predecessor runs step_adjust for the first time
        l5 = nat.launch_count()
        # iterations timed one by one, median reported: an iteration is ~0.45 s, and a single allocator or clock hiccup in a
        # two-iteration mean moved the figure by 15 % between otherwise identical runs
        ms5_all = [timed(gps_once, 1, 0) for _ in range(n5)]
        ms5 = float(np.median(ms5_all))
        l5 = (nat.launch_count() - l5) // n5
        n_steps5 = algo.last["n_steps"]
        fl5 = B5 * (T5 * fit_flops_per_fit(dX5, dU5, N5, K5) + n_steps5 * ilqr_flops_per_solve(T5, dX5, dU5)
                    + 3 * forward_flops(T5, dX5, dU5))
        by5 = B5 * (T5 * fit_bytes_per_fit(dX5, dU5, N5) + n_steps5 * ilqr_bytes_per_launch(1, 1, T5, dX5, dU5))
        # parts, timed separately
        fit5_ms = timed(lambda: batched.fit_lr(r5["X"], r5["U"], gm5), 2, 0) / 2
        g5 = {"conditions_per_sec": B5 * world / (ms5 * 1e-3), "ms_per_iteration": ms5, "batched_steps": n_steps5,
              "library_launches_per_iteration": int(l5),
              "status_counts": torch.bincount(algo.last["status"], minlength=5).tolist(),
              "kl_step_mult_range": [float(algo.kl_step_mult.min()), float(algo.kl_step_mult.max())],
              "ms_per_iteration_samples": ms5_all, "parts_ms": {"fit_lr_with_prior": fit5_ms},
              "config": f"BatchedTrajOpt.iteration: {B5} conditions/GPU ({B5 * world} in total), T={T5}, dX={dX5}, dU={dU5}, "
                        f"N={N5}, {K5}-cluster GMM prior (BASELINE.json configs[4]); steady-state iteration incl. step_adjust",
              "roofline": roof(fl5, by5, ms5, f"fit_coop_kernel<{dX5},{dU5}> + gmm_e_mma_kernel + ilqr_coop_kernel<{dX5},{dU5}>")}
        # e2e form: roll-outs, costs and previous policy in pinned host memory; policy back to the host
        hX, hU = pin(r5["X"].cpu()), pin(r5["U"].cpu())
        hC, hc, hcc = pin(C5.cpu()), pin(c5v.cpu()), pin(cc5.cpu())
        pol_host = [pin(torch.empty_like(t, device="cpu")) for t in algo.pol]
        del r5

        def gps_e2e():
            # chunked upload on a copy stream, overlapped with the state fit / fit_lr of the chunks already on the device
            algo.iteration_from_host(hX, hU, hC, hc, hcc, prior=gm5, chunks=4, out_policy=pol_host)
            sync()

        e5 = float(np.median([timed(gps_e2e, 1, 0) for _ in range(n5 + (0 if dry else 1))][0 if dry else 1:]))
        h2d5 = sum(t.numel() * 8 for t in (hX, hU, hC, hc, hcc))
        g5["e2e"] = {"conditions_per_sec": B5 * world / (e5 * 1e-3), "ms_per_iteration": e5, "h2d_bytes_per_iteration": h2d5,
                     "d2h_bytes_per_iteration": sum(t.numel() * 8 for t in pol_host),
                     "upload_exposed_ms": e5 - ms5, "chunks": 4, "numa": numa,
                     "what": "BatchedTrajOpt.iteration_from_host: roll-outs + costs uploaded from pinned host memory in chunks "
                             "of conditions (copy stream) while fit_lr runs on the chunks that have landed, the rest of the "
                             "iteration on the device, the new policy downloaded"}
        extra["gps_iteration"] = g5
        del hX, hU, hC, hc, hcc
        algo._host_buf = algo._host_costs = None
        holder.clear()
        free()
        # the prior update of the same GPS iteration: EM iterations over the (sharded) pool of config-5 transitions
        n5p = max(K5 * 4, c5["n_pool"] // world)
        Xp, w0, m0, p0 = synthetic.gmm_pool_torch(n5p, d5, K5, seed=5100 + rank, device=dev)
        gmu = batched.DeviceGMM(K5, tol=0.0, max_iter=1).set_params(w0, m0, precisions=p0)
        ws = torch.empty(nat.size("gmm_workspace_bytes", n5p, d5, K5) // 8 + 1, **f64)
        stats = torch.empty(nat.size("gmm_stats_size", d5, K5), **f64)
        gmu._lb = torch.zeros(1, **f64)
        gmu._info = torch.zeros(K5, dtype=torch.int32, device=dev)
        emu_ms = timed_median(lambda: gmu.em_iteration(Xp, float(n5p * world), ws, stats), 3, 1, inner=2)
        extra["gps_iteration"]["prior_update"] = {
            "ms_per_em_iteration": emu_ms, "rows_total": n5p * world, "d": d5, "K": K5,
            "roofline": roof(em_flops_per_row(d5, K5) * n5p, n5p * d5 * 8, emu_ms, "gmm_e_mma_kernel + gmm_m_mma_kernel")}
        del Xp, ws, gmu
    except Exception as ex:
        extra.setdefault("gps_iteration", {})["error"] = repr(ex)[:300]
    free()

    # ---- CPU baseline (rank 0, N=1 only): the reference on all host cores, bounded samples --------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and not dry:
        pool = CpuPool()
        try:
            cores = pool.cores
            rates, busy = cpu_ilqr_samples(pool, 6, 3)
            cpu = {"value": float(np.median(rates)), "unit": "solves/s", "cores": cores, "kind": cpu_kind(), "samples": rates,
                   "sample": f"3 samples of {6 * cores} conditions x 16 eta (T={T},dX={dX},dU={dU}) through "
                             + ("donk.traj_opt.ILQR.step of the unmodified reference (baseline/_ref)" if cpu_kind() == "reference"
                                else "oracle.ilqr_step")
                             + f", one single-threaded process per core, median; last sample {busy:.0f} core-seconds; "
                               f"rate = solves x cores / summed busy seconds"}

            def leg(name, jobs, unit):
                try:
                    units, busy_s = pool.run(jobs)
                    return {"value": units * cores / busy_s, "unit": unit, "cores": cores, "kind": cpu_kind(),
                            "core_seconds": busy_s, "sample": f"{len(jobs)} conditions"}
                except Exception as ex:
                    return {"error": repr(ex)[:200]}

            if "error" not in extra.get("optimize", {}):
                extra["optimize"]["cpu"] = leg("optimize", [("optimize", 3, i, None) for i in range(2 * cores)], "optimisations/s")
            if "error" not in extra.get("config1_latency", {}):
                units, busy_s = pool.run([("fit_optimize", 1, i, None) for i in range(4 * cores)])
                extra["config1_latency"]["cpu"] = {"ms_fit_lr_plus_optimize_one_core": 1e3 * busy_s / units,
                                                   "conditions_per_sec_all_cores": units * cores / busy_s, "cores": cores,
                                                   "kind": cpu_kind()}
            if prior2 is not None and "error" not in extra.get("fit_lr_prior", {}):
                extra["fit_lr_prior"]["cpu"] = leg("fit_prior", [("fit_prior", 2, i, prior2) for i in range(2 * cores)],
                                                   "timestep fits/s")
            if prior5 is not None and "error" not in extra.get("gps_iteration", {}):
                extra["gps_iteration"]["cpu"] = leg("gps", [("gps", 5, i, prior5) for i in range(cores)], "conditions/s")
        finally:
            pool.close()

    if rank == 0:
        line = {
            "metric": "ilqr_solves_per_sec", "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "h2d_gbs_per_rank": h2d / (e2e_ms * 1e-3) / 1e9, "numa": numa},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp64_peak if fp64_peak else None, "traffic": traffic,
                         "kernel": f"ilqr_warp_kernel<double,{dX},{dU}> (products on mma.sync.m8n8k4.f64)",
                         "peak_source": peak_src,
                         "algorithmic_flops_per_launch": flops_per_launch,
                         "hbm": {"achieved": bytes_per_launch / (ms_per_step * 1e-3) / 1e9, "peak": hbm_peak,
                                 "unit": "GB/s", "algorithmic_bytes_per_launch": bytes_per_launch,
                                 "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)"}},
            "cpu_baseline": cpu,
            "parity_max_rel_err_vs_oracle": parity,
            "em_strong_scaling": em_strong,
            "extra": extra,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def _reserve_stdout():
    """stdout carries exactly one JSON line: keep the real stdout aside and point fd 1 at stderr, so that nothing a
    library prints (the NCCL version banner is written to fd 1 from C) can land next to the result."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    _reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--em-rows", type=int, default=10_000_000, help="rows of the GMM-EM extra (total over GPUs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="condition chunks of the host-buffer pipeline")
    ap.add_argument("--dry-run-emu", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
