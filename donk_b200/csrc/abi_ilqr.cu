// donk_b200 — C ABI, part: iLQR step (ILQR.step / backward / forward, donk/traj_opt/lqg.py:55-244).
#include "abi_common.cuh"
#include "ilqr.cuh"
#include "ilqr_warp.cuh"

using namespace donk;
using namespace donk_abi;

namespace {

// ---------------------------------------------------------------------------------------------
// iLQR
// ---------------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(512, 1) ilqr_step_kernel(const IlqrArgs<R> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx ctx{(int)threadIdx.x, (int)blockDim.x};
  ilqr_cta<R>(ctx, a, (int)blockIdx.x, reinterpret_cast<R*>(smem_raw));
}
// Same body compiled for <= 320 threads with a register budget that lets several CTAs share an SM.
template <typename R>
__global__ void __launch_bounds__(320, 2) ilqr_step_kernel_occ2(const IlqrArgs<R> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx ctx{(int)threadIdx.x, (int)blockDim.x};
  ilqr_cta<R>(ctx, a, (int)blockIdx.x, reinterpret_cast<R*>(smem_raw));
}

// Warp-per-problem fast path with compile-time dimensions (ilqr_warp.cuh).
template <typename R, int DX, int DU, bool FULL>
__global__ void __launch_bounds__(512, 1) ilqr_warp_kernel_big(const IlqrArgs<R> a, int ES) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WCtx w{(int)threadIdx.x, (int)blockDim.x, (int)threadIdx.x >> 5, (int)threadIdx.x & 31, (int)blockDim.x >> 5};
  ilqr_warp_cta<R, DX, DU, FULL>(w, a, (int)blockIdx.x, ES, reinterpret_cast<R*>(smem_raw));
}
// register budget of the warp kernel: 8-warp CTAs x 2 per SM = 128 registers per thread (occupancy experiments build
// other shapes with -DDONK_WARP_LB_THREADS / -DDONK_WARP_LB_CTAS, profiles/build_alt.py)
#ifndef DONK_WARP_LB_THREADS
#define DONK_WARP_LB_THREADS 256
#define DONK_WARP_LB_CTAS 2
#endif
template <typename R, int DX, int DU, bool FULL>
__global__ void __launch_bounds__(DONK_WARP_LB_THREADS, DONK_WARP_LB_CTAS) ilqr_warp_kernel(const IlqrArgs<R> a, int ES) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WCtx w{(int)threadIdx.x, (int)blockDim.x, (int)threadIdx.x >> 5, (int)threadIdx.x & 31, (int)blockDim.x >> 5};
  ilqr_warp_cta<R, DX, DU, FULL>(w, a, (int)blockIdx.x, ES, reinterpret_cast<R*>(smem_raw));
}

template <typename R, int DX, int DU, bool FULL>
int ilqr_warp_launch_v(const IlqrArgs<R>& a, int G, int ES, int grid, size_t bytes, void* stream) {
  int rc;
  if (32 * G <= DONK_WARP_LB_THREADS) {
    rc = allow_smem(ilqr_warp_kernel<R, DX, DU, FULL>, bytes);
    if (rc != DONK_OK) return rc;
    ilqr_warp_kernel<R, DX, DU, FULL><<<grid, 32 * G, bytes, (cudaStream_t)stream>>>(a, ES);
  } else {
    rc = allow_smem(ilqr_warp_kernel_big<R, DX, DU, FULL>, bytes);
    if (rc != DONK_OK) return rc;
    ilqr_warp_kernel_big<R, DX, DU, FULL><<<grid, 32 * G, bytes, (cudaStream_t)stream>>>(a, ES);
  }
  return DONK_OK;
}

template <typename R, int DX, int DU>
int ilqr_warp_launch(const IlqrArgs<R>& a, void* stream) {
  int G, ES, grid;
  size_t bytes;
  int rc = ilqr_warp_plan<R, DX, DU>(a.B, a.E, smem_optin_cap(), env_int("DONK_ILQR_WARPS"), G, ES, grid, bytes);
  if (rc != DONK_OK) return rc;
  bytes += (size_t)env_int("DONK_ILQR_SMEM_PAD_KB") * 1024;  // occupancy experiments (profiles/prof_phases.py)
  if (bytes > smem_optin_cap()) return DONK_SMEM_LIMIT;
  // the everything-enabled call (ILQR.step / .optimize) runs the instantiation with the options folded at compile time
  rc = ilqr_warp_is_full<R, DX, DU>(a) ? ilqr_warp_launch_v<R, DX, DU, true>(a, G, ES, grid, bytes, stream)
                                      : ilqr_warp_launch_v<R, DX, DU, false>(a, G, ES, grid, bytes, stream);
  if (rc != DONK_OK) return rc;
  LAUNCHED("ilqr_warp_kernel");
  return DONK_OK;
}


template <typename R>
int ilqr_step_launch(IlqrArgs<R> a, void* stream) {
  if (a.B < 0 || a.E < 1 || a.T < 1 || a.dX < 1 || a.dU < 1) return DONK_BAD_SHAPE;
  if ((a.flags & ILQR_KL) && (a.flags & (ILQR_BACKWARD | ILQR_FORWARD)) != (ILQR_BACKWARD | ILQR_FORWARD))
    return DONK_BAD_SHAPE;
  if (a.B == 0) return DONK_OK;
  IlqrPlan pl;
  int rc;
  if (!env_int("DONK_ILQR_GENERIC")) {
#define DONK_TRY_WARP(DXV, DUV) \
  if (a.dX == DXV && a.dU == DUV) return ilqr_warp_launch<R, DXV, DUV>(a, stream);
    DONK_WARP_DIMS(DONK_TRY_WARP)
#undef DONK_TRY_WARP
    if constexpr (sizeof(R) == 8) {  // large state dimensions: one problem per CTA, products on DMMA tiles (ilqr_coop.cuh)
      const int rc_c = ilqr_coop_try_launch(reinterpret_cast<const IlqrArgs<double>&>(a), stream);
      if (rc_c >= 0 && rc_c != DONK_SMEM_LIMIT) return rc_c;
    }
  }
  rc = ilqr_plan(a.B, a.E, a.dX, a.dU, sizeof(R), smem_optin_cap(), env_int("DONK_ILQR_SLOTS"),
                 env_int("DONK_ILQR_THREADS"), pl);
  if (rc != DONK_OK) {
    snprintf(g_err, sizeof(g_err), "ilqr_step: dX=%d dU=%d does not fit %zu bytes of shared memory", a.dX, a.dU,
             smem_optin_cap());
    return rc;
  }
  a.S = pl.S;
  a.ES = pl.ES;
  if (pl.threads <= 320 && !env_int("DONK_ILQR_OCC1")) {
    rc = allow_smem(ilqr_step_kernel_occ2<R>, pl.smem);
    if (rc != DONK_OK) return rc;
    ilqr_step_kernel_occ2<R><<<pl.grid, pl.threads, pl.smem, (cudaStream_t)stream>>>(a);
  } else {
    rc = allow_smem(ilqr_step_kernel<R>, pl.smem);
    if (rc != DONK_OK) return rc;
    ilqr_step_kernel<R><<<pl.grid, pl.threads, pl.smem, (cudaStream_t)stream>>>(a);
  }
  LAUNCHED("ilqr_step_kernel");
  return DONK_OK;
}

}  // namespace

namespace donk_abi {
// the step launch for other translation units (abi_optimize.cu captures it into the graph of ILQR.optimize)
int ilqr_step_launch_f64(const IlqrArgs<double>& a, void* stream) { return ilqr_step_launch<double>(a, stream); }
}  // namespace donk_abi

extern "C" {

#define ILQR_STEP_IMPL(SUFFIX, R)                                                                             \
  int donk_ilqr_step_##SUFFIX(int B, int E, int T, int dX, int dU, int flags, const R* F, const R* f,         \
                              const R* dyn_covar, const R* C, const R* c, const R* cc, const R* C_kl,         \
                              const R* c_kl, const R* prevK, const R* prevk, const R* prevLam,                \
                              const R* prev_hld, const R* x0_mean, const R* x0_covar, const R* eta, R gamma,  \
                              R fwd_reg, const int* active, R* K, R* k, R* pol_covar, R* pol_inv_covar,       \
                              R* kl, R* cost, R* traj_mean, R* traj_covar, int* info, void* stream) {         \
    IlqrArgs<R> a;                                                                                            \
    memset(&a, 0, sizeof(a));                                                                                 \
    a.B = B; a.E = E; a.T = T; a.dX = dX; a.dU = dU; a.flags = flags;                                         \
    a.F = F; a.f = f; a.dyn_covar = dyn_covar; a.C = C; a.c = c; a.cc = cc; a.C_kl = C_kl; a.c_kl = c_kl;     \
    a.prevK = prevK; a.prevk = prevk; a.prevLam = prevLam; a.prev_hld = prev_hld;                             \
    a.x0m = x0_mean; a.x0S = x0_covar; a.eta = eta; a.gamma = gamma; a.fwd_reg = fwd_reg;                     \
    a.K = K; a.k = k; a.pcov = pol_covar; a.pinv = pol_inv_covar; a.kl = kl; a.cost = cost;                   \
    a.tmean = traj_mean; a.tcov = traj_covar; a.info = info; a.active = active;                               \
    a.aligned16 = (((uintptr_t)F | (uintptr_t)C_kl | (uintptr_t)dyn_covar | (uintptr_t)C | (uintptr_t)K) & 15) == 0;                      \
    return ilqr_step_launch<R>(a, stream);                                                                    \
  }
ILQR_STEP_IMPL(f64, double)
ILQR_STEP_IMPL(f32, float)

#if defined(DONK_PHASE_PROF)
// profiling builds only (profiles/prof_phases.py): read / reset the per-phase cycle counters of the warp kernel
int donk_phase_prof_read(unsigned long long* out32, int reset) {
  if (out32) cudaMemcpyFromSymbol(out32, donk::donk_phase_cycles, 32 * sizeof(unsigned long long));
  if (reset) {
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(donk::donk_phase_cycles, z, sizeof(z));
  }
  return DONK_OK;
}
#endif

}  // extern "C"
