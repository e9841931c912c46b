// donk_b200 — CUDA library: __global__ wrappers, launch configuration and the C ABI
// declared in include/donk_b200.h.  Built for sm_100a only (no other code path).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "../../include/donk_b200.h"
#include "fitlr.cuh"
#include "fitlr_warp.cuh"
#include "gmm.cuh"
#include "gmm_mma.cuh"
#include "ilqr.cuh"
#include "kmeans.cuh"
#include <cstdint>
#include "ilqr_warp.cuh"

using namespace donk;

// (dX, dU) pairs with warp-kernel instantiations (ILQR step, fit_lr): the BASELINE.json configurations that
// fit one warp per problem; other dimensions run the generic CTA kernels.
#define DONK_WARP_DIMS(X) X(20, 6) X(14, 7) X(6, 2)

namespace {

thread_local char g_err[256] = "";
std::atomic<unsigned long long> g_launches{0};

int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return DONK_CUDA_ERROR;
}
#define CU(call)                                           \
  do {                                                     \
    cudaError_t e_ = (call);                               \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call);    \
  } while (0)
#define LAUNCHED(name)                                     \
  do {                                                     \
    g_launches.fetch_add(1, std::memory_order_relaxed);    \
    cudaError_t e_ = cudaGetLastError();                   \
    if (e_ != cudaSuccess) return cuda_fail(e_, name);     \
  } while (0)

int env_int(const char* name) {
  const char* v = getenv(name);
  return v ? atoi(v) : 0;
}

size_t smem_optin_cap() {
  static size_t cap = 0;
  if (!cap) {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cap = v > 0 ? (size_t)v : 48 * 1024;
  }
  return cap;
}

template <typename KernelT>
int allow_smem(KernelT kernel, size_t bytes) {
  if (bytes > 48 * 1024) CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  // ask for the full shared-memory carve-out so that several CTAs fit one SM
  CU(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  return DONK_OK;
}

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
template <typename R, int DX, int DU, bool FULL>
__global__ void __launch_bounds__(256, 2) ilqr_warp_kernel(const IlqrArgs<R> a, int ES) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WCtx w{(int)threadIdx.x, (int)blockDim.x, (int)threadIdx.x >> 5, (int)threadIdx.x & 31, (int)blockDim.x >> 5};
  ilqr_warp_cta<R, DX, DU, FULL>(w, a, (int)blockIdx.x, ES, reinterpret_cast<R*>(smem_raw));
}

template <typename R, int DX, int DU, bool FULL>
int ilqr_warp_launch_v(const IlqrArgs<R>& a, int G, int ES, int grid, size_t bytes, void* stream) {
  int rc;
  if (G <= 8) {
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

template <typename R>
__global__ void __launch_bounds__(256) extended_costs_kernel(int n, int npairs, int dX, int dU, const R* K, const R* k,
                                                             const R* Lam, R* C, R* c) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx ctx{(int)threadIdx.x, (int)blockDim.x};
  extended_costs_cta<R>(ctx, n, npairs, (int)blockIdx.x, dX, dU, K, k, Lam, C, c, reinterpret_cast<R*>(smem_raw));
}

template <typename R>
__global__ void spd_factor_kernel(int n, int d, const R* A, R* Ainv, R* chol, R* hld, int* info) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* ws = reinterpret_cast<R*>(smem_raw) + (size_t)threadIdx.x * 2 * d * d;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n) return;
  const size_t o = i * d * d;
  const int bad = spd_inverse_item<R>(d, A + o, Ainv ? Ainv + o : nullptr, chol ? chol + o : nullptr,
                                      hld ? hld + i : nullptr, ws);
  if (info) info[i] = bad;
}

template <typename R>
__global__ void kl_action_kernel(int T, int dX, int dU, const R* X, const R* K, const R* k, const R* covar,
                                 const R* hld, const R* pK, const R* pk, const R* pLam, const R* phld, R* kl) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* terms = reinterpret_cast<R*>(smem_raw);           // T
  R* delta = terms + T + (size_t)threadIdx.x * dU;     // blockDim * dU
  const size_t b = blockIdx.x;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const size_t bt = b * T + t;
    terms[t] = kl_term_item<R>(dX, dU, X + bt * dX, K + bt * dU * dX, k + bt * dU, covar + bt * dU * dU, hld[bt],
                               pK + bt * dU * dX, pk + bt * dU, pLam + bt * dU * dU, phld[bt], delta);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    R acc = R(0);
    for (int t = 0; t < T; ++t) acc += terms[t];
    kl[b] = acc / R(T);
  }
}

template <typename R>
__global__ void brent_kernel(int B, int phase, R xa, R xb, const R* fa, const R* fb, const R* fnew, R kl_step, R xtol,
                             R rtol, int maxiter, R* state, R* eta, int* active, int* status, int* n_active) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  BrentState<R>* st = reinterpret_cast<BrentState<R>*>(state) + b;
  if (active[b] == 0) return;  // not searched / already finished
  int act = 0;
  R e = eta[b];
  brent_item<R>(*st, phase, xa, xb, phase == 0 ? fa[b] - kl_step : R(0), phase == 0 ? fb[b] - kl_step : R(0),
                phase == 0 ? R(0) : fnew[b] - kl_step, xtol, rtol, maxiter, &e, &act);
  eta[b] = e;
  active[b] = act;
  status[b] = (int)st->status;
  if (act) atomicAdd(n_active, 1);
}


// ---------------------------------------------------------------------------------------------
// fit_lr, to_transitions, StateDistribution.fit
// ---------------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(512, 1) fit_lr_kernel(const FitArgs<R> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx ctx{(int)threadIdx.x, (int)blockDim.x};
  fit_lr_cta<R>(ctx, a, (int)blockIdx.x, reinterpret_cast<R*>(smem_raw));
}

template <typename R, int DX, int DU>
__global__ void __launch_bounds__(128, 3) fit_warp_kernel(const FitArgs<R> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WCtx w{(int)threadIdx.x, (int)blockDim.x, (int)threadIdx.x >> 5, (int)threadIdx.x & 31, (int)blockDim.x >> 5};
  fit_warp_cta<R, DX, DU>(w, a, (int)blockIdx.x, reinterpret_cast<R*>(smem_raw));
}
template <typename R>
__global__ void gmm_prior_consts_kernel(int d, int K, const R* gw, const R* gmeans, const R* gpc, R* out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < K * d + K) gmm_prior_consts_item<R>(d, K, gw, gmeans, gpc, out, idx);
}

template <typename R, int DX, int DU>
int fit_warp_launch(FitArgs<R> a, void* stream) {
  int G;
  size_t bytes;
  const int forced = env_int("DONK_FIT_WARPS");
  int rc = fit_warp_plan<R, DX, DU>(smem_optin_cap() / 3 - 1024, forced, a.N, a.K, G, bytes);
  if (rc != DONK_OK) rc = fit_warp_plan<R, DX, DU>(smem_optin_cap(), forced, a.N, a.K, G, bytes);
  if (rc != DONK_OK) return rc;
  if (G > 4) {  // the kernel is compiled for <= 128 threads
    G = 4;
    bytes = (size_t)4 * (FitWarpDims<DX, DU>::TEAM + FitWarpDims<DX, DU>::prior_tail(a.N, a.K)) * sizeof(R);
  }
  rc = allow_smem(fit_warp_kernel<R, DX, DU>, bytes);
  if (rc != DONK_OK) return rc;
  R* consts = nullptr;
  if (a.K > 0) {  // per-cluster constants of the log-probabilities: stream-ordered scratch, freed after the fit launch
    const int d = 2 * DX + DU, len = a.K * d + a.K;
    if (cudaMallocAsync(reinterpret_cast<void**>(&consts), (size_t)len * sizeof(R), (cudaStream_t)stream) != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "fit_lr: cudaMallocAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
      return DONK_CUDA_ERROR;
    }
    gmm_prior_consts_kernel<R><<<(len + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d, a.K, a.gw, a.gmeans, a.gpc, consts);
    LAUNCHED("gmm_prior_consts_kernel");
    a.gconst = consts;
  }
  const long long nfits = (long long)a.B * a.T;
  fit_warp_kernel<R, DX, DU><<<(unsigned)((nfits + G - 1) / G), 32 * G, bytes, (cudaStream_t)stream>>>(a);
  if (consts) cudaFreeAsync(consts, (cudaStream_t)stream);
  LAUNCHED("fit_warp_kernel");
  return DONK_OK;
}

template <typename R>
__global__ void to_transitions_kernel(size_t total, int T, int dX, int dU, const R* X, const R* U, R* out) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
    out[idx] = transitions_item<R>(T, dX, dU, X, U, idx);
}

template <typename R>
__global__ void state_fit_kernel(size_t total, int N, int dX, const R* X0, size_t ss, size_t cs, R* mean, R* covar) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < total) state_fit_item<R>(N, dX, X0, ss, cs, idx, mean, covar);
}

// ---------------------------------------------------------------------------------------------
// GMM EM
// ---------------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(256) gmm_e_kernel(const GmmArgs<R> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx ctx{(int)threadIdx.x, (int)blockDim.x};
  gmm_e_cta<R>(ctx, a, (int)blockIdx.x, (int)gridDim.x, reinterpret_cast<R*>(smem_raw));
}
template <typename R>
__global__ void __launch_bounds__(384) gmm_m_kernel(const GmmArgs<R> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx ctx{(int)threadIdx.x, (int)blockDim.x};
  gmm_m_cta<R>(ctx, a, (int)blockIdx.x, reinterpret_cast<R*>(smem_raw));
}
// The last block sums the per-tile log-sum-exp totals cooperatively (78 k of them at 10 M rows: one thread walking
// them took 2 ms); every thread adds a strided slice in a fixed order, then a fixed-order tree: deterministic.
template <typename R>
__global__ void __launch_bounds__(128) gmm_reduce_kernel(const GmmArgs<R> a, size_t total, R* stats) {
  if (blockIdx.x + 1 == gridDim.x) {
    __shared__ R part[128];
    R acc = R(0);
    for (int bI = threadIdx.x; bI < a.nblocks; bI += 128) acc += a.lse_part[bI];
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) stats[total - 1] = part[0];
    return;
  }
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx + 1 < total) gmm_reduce_item<R>(a, idx, stats);
}
template <typename R>
__global__ void __launch_bounds__(256) gmm_finalize_kernel(const GmmFinArgs<R> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx ctx{(int)threadIdx.x, (int)blockDim.x};
  gmm_finalize_cta<R>(ctx, a, (int)blockIdx.x, reinterpret_cast<R*>(smem_raw));
}
template <typename R>
__global__ void __launch_bounds__(256) gmm_prec_chol_kernel(int d, const R* prec, R* pc, int* info) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx ctx{(int)threadIdx.x, (int)blockDim.x};
  gmm_prec_chol_cta<R>(ctx, d, prec, pc, info, (int)blockIdx.x, reinterpret_cast<R*>(smem_raw));
}

template <int NB, int RA>
__global__ void __launch_bounds__(256, 1) gmm_e_mma_kernel(const GmmArgs<double> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WCtx w{(int)threadIdx.x, (int)blockDim.x, (int)threadIdx.x >> 5, (int)threadIdx.x & 31, (int)blockDim.x >> 5};
  gmm_e_mma_cta<double, NB, RA>(w, a, (int)blockIdx.x, (int)gridDim.x, reinterpret_cast<double*>(smem_raw));
}
template <int NB, int RA>
__global__ void __launch_bounds__(256, 2) gmm_e_mma_kernel_occ2(const GmmArgs<double> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WCtx w{(int)threadIdx.x, (int)blockDim.x, (int)threadIdx.x >> 5, (int)threadIdx.x & 31, (int)blockDim.x >> 5};
  gmm_e_mma_cta<double, NB, RA>(w, a, (int)blockIdx.x, (int)gridDim.x, reinterpret_cast<double*>(smem_raw));
}
template <int NBW, int KC, int THREADS, int OCC>
__global__ void __launch_bounds__(THREADS, OCC) gmm_m_mma_kernel(const GmmArgs<double> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WCtx w{(int)threadIdx.x, (int)blockDim.x, (int)threadIdx.x >> 5, (int)threadIdx.x & 31, (int)blockDim.x >> 5};
  gmm_m_mma_cta<double, NBW, KC>(w, a, (int)blockIdx.x, reinterpret_cast<double*>(smem_raw));
}

template <int NB, int RA>
int gmm_e_mma_launch(GmmArgs<double>& a, const GmmMmaPlan& pl, void* stream) {
  const int TS = 8 * RA * pl.Ge;
  a.nblocks = (a.n + TS - 1) / TS;
  const size_t smem = gmm_e_mma_smem<NB, RA>(a.K, pl.Ge) * sizeof(double);
  if (smem > smem_optin_cap()) return DONK_SMEM_LIMIT;
  // two CTAs per SM when they fit: the staging / softmax / barrier phases of one overlap the products of the other
  // (113.4 -> 107.0 ms per EM iteration at 10 M x 72, K = 16; DONK_GMM_E_CTAS=1 restores one CTA per SM)
  const bool occ2 = 2 * (smem + 1024) <= (size_t)228 * 1024 && 32 * pl.Ge <= 256 && env_int("DONK_GMM_E_CTAS") != 1;
  int rc = occ2 ? allow_smem(gmm_e_mma_kernel_occ2<NB, RA>, smem) : allow_smem(gmm_e_mma_kernel<NB, RA>, smem);
  if (rc != DONK_OK) return rc;
  int grid = occ2 ? 296 : 148;  // persistent CTAs walking the row tiles
  if (grid > a.nblocks) grid = a.nblocks;
  if (occ2) gmm_e_mma_kernel_occ2<NB, RA><<<grid, 32 * pl.Ge, smem, (cudaStream_t)stream>>>(a);
  else gmm_e_mma_kernel<NB, RA><<<grid, 32 * pl.Ge, smem, (cudaStream_t)stream>>>(a);
  LAUNCHED("gmm_e_mma_kernel");
  return DONK_OK;
}

int gmm_e_mma_dispatch(GmmArgs<double>& a, const GmmMmaPlan& pl, void* stream) {
  switch (pl.NB) {
    case 2: return gmm_e_mma_launch<2, 2>(a, pl, stream);
    case 5: return gmm_e_mma_launch<5, 2>(a, pl, stream);
    case 6: return gmm_e_mma_launch<6, 2>(a, pl, stream);
    case 9: return gmm_e_mma_launch<9, 2>(a, pl, stream);
    default: return gmm_e_mma_launch<18, 1>(a, pl, stream);
  }
}

int gmm_m_mma_dispatch(GmmArgs<double>& a, const GmmMmaPlan& pl, void* stream) {
  const size_t smem = gmm_m_mma_smem(a.d, pl.KC) * sizeof(double);
  if (smem > smem_optin_cap()) return DONK_SMEM_LIMIT;
  const int grid = ((a.K + pl.KC - 1) / pl.KC) * a.nch;
  int rc;
  if (pl.NBW == 3) {  // 16 warps x 3 blocks x 4 clusters, one CTA per SM
    rc = allow_smem(gmm_m_mma_kernel<3, 4, 512, 1>, smem);
    if (rc != DONK_OK) return rc;
    gmm_m_mma_kernel<3, 4, 512, 1><<<grid, 32 * pl.Gm, smem, (cudaStream_t)stream>>>(a);
  } else if (pl.NBW <= 6) {  // 8 warps x 6 blocks x 2 clusters, two CTAs per SM
    rc = allow_smem(gmm_m_mma_kernel<6, 2, 256, 2>, smem);
    if (rc != DONK_OK) return rc;
    gmm_m_mma_kernel<6, 2, 256, 2><<<grid, 32 * pl.Gm, smem, (cudaStream_t)stream>>>(a);
  } else {
    rc = allow_smem(gmm_m_mma_kernel<11, 2, 512, 1>, smem);
    if (rc != DONK_OK) return rc;
    gmm_m_mma_kernel<11, 2, 512, 1><<<grid, 32 * pl.Gm, smem, (cudaStream_t)stream>>>(a);
  }
  LAUNCHED("gmm_m_mma_kernel");
  return DONK_OK;
}

int gmm_e_launch(GmmArgs<double>& a, void* stream) {
  const size_t smem = gmm_e_smem(a.d, a.K) * sizeof(double);
  if (smem > smem_optin_cap()) return DONK_SMEM_LIMIT;
  int rc = allow_smem(gmm_e_kernel<double>, smem);
  if (rc != DONK_OK) return rc;
  int per_sm = (int)((228 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int grid = 148 * per_sm;
  if (grid > a.nblocks) grid = a.nblocks;
  gmm_e_kernel<double><<<grid, 256, smem, (cudaStream_t)stream>>>(a);
  LAUNCHED("gmm_e_kernel");
  return DONK_OK;
}


// ---------------------------------------------------------------------------------------------
// k-means passes (device initialisation of the GMM prior)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 3) kmeans_pass_kernel(int n, int d, int K, const double* X, const double* centers,
                                                             int* labels, double* mind2, double* dist, double* part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx ctx{(int)threadIdx.x, (int)blockDim.x};
  kmeans_pass_cta<double>(ctx, n, d, K, X, centers, labels, mind2, dist, part, (int)blockIdx.x, (int)gridDim.x,
                          reinterpret_cast<double*>(smem_raw));
}
__global__ void kmeans_reduce_kernel(int ncta, int len, const double* part, double* out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < len) kmeans_reduce_item<double>(ncta, len, part, out, idx);
}

// ---------------------------------------------------------------------------------------------
// FP64 pipe probes (roofline denominators for the FP64-bound kernels; MEASURED_PEAKS.json only holds
// HBM and bf16 tensor peaks).  mode 0: dependent-chain-free DFMA on CUDA cores; mode 1: DMMA
// (mma.sync.m8n8k4.f64) to see whether the tensor path offers more FP64 throughput on sm_100a.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_probe_kernel(int mode, int iters, double* out) {
  const double s = 1.0 + 1e-9 * threadIdx.x;
  if (mode == 0) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 0.5 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fma(a[i], s, 1e-3);
    }
    double acc = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += a[i];
    if (acc == 123.456) out[0] = acc;
  } else {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    const double av = s, bv = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(av), "d"(bv));
    }
    double acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += c[i][0] + c[i][1];
    if (acc == 123.456) out[0] = acc;
  }
}

}  // namespace

extern "C" {

const char* donk_version(void) { return "donk_b200 0.1 sm_100a"; }
const char* donk_last_error(void) { return g_err; }
unsigned long long donk_launch_count(void) { return g_launches.load(); }

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
    a.aligned16 = (((uintptr_t)F | (uintptr_t)C_kl | (uintptr_t)dyn_covar) & 15) == 0;                        \
    return ilqr_step_launch<R>(a, stream);                                                                    \
  }
ILQR_STEP_IMPL(f64, double)
ILQR_STEP_IMPL(f32, float)

int donk_extended_costs_kl_f64(int n, int dX, int dU, const double* K, const double* k, const double* Lam,
                               double* C_kl, double* c_kl, void* stream) {
  if (n < 0 || dX < 1 || dU < 1) return DONK_BAD_SHAPE;
  if (n == 0) return DONK_OK;
  const int npairs = 4;
  const size_t smem = extended_costs_smem(dX, dU) * npairs * sizeof(double);
  if (smem > smem_optin_cap()) return DONK_SMEM_LIMIT;
  int rc = allow_smem(extended_costs_kernel<double>, smem);
  if (rc != DONK_OK) return rc;
  extended_costs_kernel<double><<<(unsigned)((n + npairs - 1) / npairs), 256, smem, (cudaStream_t)stream>>>(
      n, npairs, dX, dU, K, k, Lam, C_kl, c_kl);
  LAUNCHED("extended_costs_kernel");
  return DONK_OK;
}

int donk_spd_factor_f64(int n, int d, const double* A, double* Ainv, double* chol, double* half_logdet, int* info,
                        void* stream) {
  if (n < 0 || d < 1) return DONK_BAD_SHAPE;
  if (n == 0) return DONK_OK;
  const size_t per = (size_t)2 * d * d * sizeof(double);
  int threads = 32;
  while (threads > 1 && per * threads > 96 * 1024) threads /= 2;
  if (per * threads > smem_optin_cap()) return DONK_SMEM_LIMIT;
  int rc = allow_smem(spd_factor_kernel<double>, per * threads);
  if (rc != DONK_OK) return rc;
  spd_factor_kernel<double><<<(n + threads - 1) / threads, threads, per * threads, (cudaStream_t)stream>>>(
      n, d, A, Ainv, chol, half_logdet, info);
  LAUNCHED("spd_factor_kernel");
  return DONK_OK;
}

int donk_kl_divergence_action_f64(int B, int T, int dX, int dU, const double* X, const double* K, const double* k,
                                  const double* covar, const double* hld, const double* pK, const double* pk,
                                  const double* pLam, const double* phld, double* kl, void* stream) {
  if (B < 0 || T < 1 || dX < 1 || dU < 1) return DONK_BAD_SHAPE;
  if (B == 0) return DONK_OK;
  const int threads = 64;
  const size_t smem = ((size_t)T + (size_t)threads * dU) * sizeof(double);
  if (smem > smem_optin_cap()) return DONK_SMEM_LIMIT;
  int rc = allow_smem(kl_action_kernel<double>, smem);
  if (rc != DONK_OK) return rc;
  kl_action_kernel<double><<<B, threads, smem, (cudaStream_t)stream>>>(T, dX, dU, X, K, k, covar, hld, pK, pk, pLam,
                                                                        phld, kl);
  LAUNCHED("kl_action_kernel");
  return DONK_OK;
}

int donk_brent_update_f64(int B, int phase, double xa, double xb, const double* fa, const double* fb,
                          const double* fnew, double kl_step, double xtol, double rtol, int maxiter, double* state,
                          double* eta, int* active, int* status, int* n_active, void* stream) {
  if (B < 0) return DONK_BAD_SHAPE;
  if (B == 0) return DONK_OK;
  CU(cudaMemsetAsync(n_active, 0, sizeof(int), (cudaStream_t)stream));
  brent_kernel<double><<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, phase, xa, xb, fa, fb, fnew, kl_step,
                                                                           xtol, rtol, maxiter, state, eta, active,
                                                                           status, n_active);
  LAUNCHED("brent_kernel");
  return DONK_OK;
}

int donk_fit_lr_f64(int B, int N, int T, int dX, int dU, const double* X, const double* U, int gmm_K,
                    const double* gmm_weights, const double* gmm_means, const double* gmm_covariances,
                    const double* gmm_prec_chol, double n_pool, double prior_weight, double regularization,
                    double* F, double* f, double* covar, int* info, void* stream) {
  if (B < 0 || N <= 1 || T < 1 || dX < 1 || dU < 1 || gmm_K < 0) return DONK_BAD_SHAPE;
  if (B == 0) return DONK_OK;
  FitArgs<double> a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.N = N; a.T = T; a.dX = dX; a.dU = dU; a.X = X; a.U = U; a.K = gmm_K;
  a.gw = gmm_weights; a.gmeans = gmm_means; a.gcov = gmm_covariances; a.gpc = gmm_prec_chol;
  a.n_pool = n_pool; a.prior_weight = prior_weight; a.reg = regularization;
  a.F = F; a.f = f; a.covar = covar; a.info = info;
  // warp-per-fit fast path for the instantiated dimensions (with the prior: partial sums for at most 64 samples)
  if ((gmm_K == 0 || N <= 64) && !env_int("DONK_FIT_GENERIC")) {
#define DONK_TRY_FITW(DXV, DUV) \
  if (dX == DXV && dU == DUV) return fit_warp_launch<double, DXV, DUV>(a, stream);
    DONK_WARP_DIMS(DONK_TRY_FITW)
#undef DONK_TRY_FITW
  }
  size_t bytes = 0;
  const int force = env_int("DONK_FIT_CHUNK");
  a.NS = fit_plan(dX, dU, gmm_K, force > 0 ? force : N, sizeof(double), smem_optin_cap(), bytes);
  if (a.NS == 0) {
    snprintf(g_err, sizeof(g_err), "fit_lr: dX=%d dU=%d K=%d does not fit shared memory", dX, dU, gmm_K);
    return DONK_SMEM_LIMIT;
  }
  { FitLayout L(dX, dU, gmm_K, N, a.NS); bytes = (size_t)L.total * sizeof(double); }
  int rc = allow_smem(fit_lr_kernel<double>, bytes);
  if (rc != DONK_OK) return rc;
  const int dp = round_up(2 * dX + dU, 4), nt = dp / 4;
  int threads = round_up(nt * (nt + 1) / 2, 32);
  if (threads < 128) threads = 128;
  if (threads > 512) threads = 512;
  if (env_int("DONK_FIT_THREADS") > 0) threads = env_int("DONK_FIT_THREADS");
  fit_lr_kernel<double><<<B * T, threads, bytes, (cudaStream_t)stream>>>(a);
  LAUNCHED("fit_lr_kernel");
  return DONK_OK;
}

int donk_to_transitions_f64(int N, int T, int dX, int dU, const double* X, const double* U, double* XUX,
                            void* stream) {
  if (N < 0 || T < 1 || dX < 1 || dU < 1) return DONK_BAD_SHAPE;
  const size_t total = (size_t)N * T * (2 * dX + dU);
  if (total == 0) return DONK_OK;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  to_transitions_kernel<double><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(total, T, dX, dU, X, U, XUX);
  LAUNCHED("to_transitions_kernel");
  return DONK_OK;
}

int donk_state_fit_f64(int B, int N, int dX, const double* X0, size_t sample_stride, size_t cond_stride,
                       double* mean, double* covar, void* stream) {
  if (B < 0 || N < 2 || dX < 1) return DONK_BAD_SHAPE;
  const size_t total = (size_t)B * dX * dX;
  if (total == 0) return DONK_OK;
  state_fit_kernel<double><<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      total, N, dX, X0, sample_stride, cond_stride, mean, covar);
  LAUNCHED("state_fit_kernel");
  return DONK_OK;
}

size_t donk_gmm_stats_size(int d, int K) { return gmm_stats_len(d, K); }
size_t donk_gmm_workspace_bytes(int n_local, int d, int K) { return gmm_ws_reals(n_local, d, K) * sizeof(double); }

int donk_gmm_em_step_mode_f64(int n_local, int d, int K, const double* X, const double* weights, const double* means,
                              const double* prec_chol, const double* resp_in, double* stats, double* shift,
                              void* workspace, size_t workspace_bytes, int mode, void* stream) {
  if (n_local < 0 || d < 1 || K < 1 || mode < 0 || mode > 2) return DONK_BAD_SHAPE;
  if (workspace_bytes < donk_gmm_workspace_bytes(n_local, d, K)) return DONK_BAD_SHAPE;
  const size_t total = gmm_stats_len(d, K);
  if (n_local == 0) {
    CU(cudaMemsetAsync(stats, 0, total * sizeof(double), (cudaStream_t)stream));
    return DONK_OK;
  }
  GmmArgs<double> a;
  memset(&a, 0, sizeof(a));
  a.n = n_local; a.d = d; a.K = K; a.X = X; a.w = weights; a.means = means; a.pc = prec_chol;
  a.nblocks = (n_local + GMM_TS - 1) / GMM_TS;
  a.nch = gmm_nch(n_local, d, K);
  double* ws = static_cast<double*>(workspace);
  a.resp = ws;
  a.lse_part = ws + (size_t)n_local * K;
  a.part = a.lse_part + a.nblocks;
  int rc = DONK_OK;
  const GmmMmaPlan pl = gmm_mma_plan(n_local, d, K);
  // Small pools take the scalar kernels: their statistics are centred on every cluster's own current mean, which
  // keeps agreement with scikit-learn at rounding level even when the clusters sit 1e3 widths apart (the reference's
  // own known-answer test on 660 real transitions needs that); the DMMA kernels centre on the mixture mean (shared
  // B fragments), which costs a few digits on such data and nothing measurable on pools with overlapping clusters.
  // mode 1 / 2 (or DONK_GMM_GENERIC / DONK_GMM_MMA) override the size rule.
  const bool want_mma = mode == 2 || (mode == 0 && (n_local >= 65536 || env_int("DONK_GMM_MMA")));
  const bool mma = pl.ok && !env_int("DONK_GMM_GENERIC") && weights != nullptr && want_mma;
  if (resp_in) {
    a.resp = const_cast<double*>(resp_in);
    CU(cudaMemsetAsync(a.lse_part, 0, (size_t)a.nblocks * sizeof(double), (cudaStream_t)stream));
  } else {
    rc = mma ? gmm_e_mma_dispatch(a, pl, stream) : gmm_e_launch(a, stream);
    if (rc != DONK_OK) return rc;
  }
  if (mma) {
    a.nch = pl.nch;
    a.shift_out = shift;
    rc = gmm_m_mma_dispatch(a, pl, stream);
    if (rc != DONK_OK) return rc;
    gmm_reduce_kernel<double><<<(unsigned)((total - 1 + 127) / 128) + 1, 128, 0, (cudaStream_t)stream>>>(a, total, stats);
    LAUNCHED("gmm_reduce_kernel");
    return DONK_OK;
  }
  if (shift) CU(cudaMemcpyAsync(shift, means, (size_t)K * d * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  const int threads = gmm_m_threads(d);
  const size_t smem = gmm_m_smem(d) * sizeof(double);
  if (threads == 0 || smem > smem_optin_cap()) return DONK_SMEM_LIMIT;
  rc = allow_smem(gmm_m_kernel<double>, smem);
  if (rc != DONK_OK) return rc;
  gmm_m_kernel<double><<<K * a.nch, threads, smem, (cudaStream_t)stream>>>(a);
  LAUNCHED("gmm_m_kernel");
  gmm_reduce_kernel<double><<<(unsigned)((total - 1 + 127) / 128) + 1, 128, 0, (cudaStream_t)stream>>>(a, total, stats);
  LAUNCHED("gmm_reduce_kernel");
  return DONK_OK;
}

int donk_gmm_em_step_f64(int n_local, int d, int K, const double* X, const double* weights, const double* means,
                         const double* prec_chol, const double* resp_in, double* stats, double* shift,
                         void* workspace, size_t workspace_bytes, void* stream) {
  return donk_gmm_em_step_mode_f64(n_local, d, K, X, weights, means, prec_chol, resp_in, stats, shift, workspace,
                                   workspace_bytes, 0, stream);
}

int donk_gmm_m_finalize_f64(double n_total, int d, int K, const double* stats, const double* shift, double reg_covar,
                            double* weights, double* means, double* covariances, double* prec_chol,
                            double* lower_bound, int* info, void* stream) {
  if (d < 1 || K < 1) return DONK_BAD_SHAPE;
  GmmFinArgs<double> a;
  a.d = d; a.K = K; a.n_total = n_total; a.reg_covar = reg_covar; a.stats = stats; a.shift = shift;
  a.w = weights; a.means = means; a.cov = covariances; a.pc = prec_chol; a.lower_bound = lower_bound; a.info = info;
  const size_t smem = gmm_fin_smem(d) * sizeof(double);
  if (smem > smem_optin_cap()) return DONK_SMEM_LIMIT;
  int rc = allow_smem(gmm_finalize_kernel<double>, smem);
  if (rc != DONK_OK) return rc;
  gmm_finalize_kernel<double><<<K, 256, smem, (cudaStream_t)stream>>>(a);
  LAUNCHED("gmm_finalize_kernel");
  return DONK_OK;
}

int donk_gmm_predict_proba_f64(int n, int d, int K, const double* X, const double* weights, const double* means,
                               const double* prec_chol, double* resp, void* stream) {
  if (n < 0 || d < 1 || K < 1) return DONK_BAD_SHAPE;
  if (n == 0) return DONK_OK;
  GmmArgs<double> a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.d = d; a.K = K; a.X = X; a.w = weights; a.means = means; a.pc = prec_chol;
  a.nblocks = (n + GMM_TS - 1) / GMM_TS;
  a.resp = resp;
  return gmm_e_launch(a, stream);
}

int donk_gmm_prec_chol_from_precisions_f64(int d, int K, const double* precisions, double* prec_chol, int* info,
                                           void* stream) {
  if (d < 1 || K < 1) return DONK_BAD_SHAPE;
  const size_t smem = gmm_fin_smem(d) * sizeof(double);
  if (smem > smem_optin_cap()) return DONK_SMEM_LIMIT;
  int rc = allow_smem(gmm_prec_chol_kernel<double>, smem);
  if (rc != DONK_OK) return rc;
  gmm_prec_chol_kernel<double><<<K, 256, smem, (cudaStream_t)stream>>>(d, precisions, prec_chol, info);
  LAUNCHED("gmm_prec_chol_kernel");
  return DONK_OK;
}

/* FP64 throughput probe: launches blocks x 256 threads, each doing `iters` rounds of 16 independent DFMA
 * (mode 0: 32 flop/thread/round) or 8 DMMA m8n8k4 (mode 1: 8*512 flop/warp/round).  Returns via *flops the
 * number of floating-point operations the launch performs; time it with events on `stream`. */
int donk_fp64_probe(int mode, int blocks, int iters, double* scratch, double* flops, void* stream) {
  if (blocks < 1 || iters < 1 || mode < 0 || mode > 1) return DONK_BAD_SHAPE;
  fp64_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(mode, iters, scratch);
  LAUNCHED("fp64_probe_kernel");
  if (flops) *flops = mode == 0 ? (double)blocks * 256 * iters * 32.0 : (double)blocks * 8 * iters * 8 * 512.0;
  return DONK_OK;
}

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

enum { KM_PASS_CTAS = 148 * 3 };
size_t donk_kmeans_workspace_bytes(int d, int K) { return (size_t)KM_PASS_CTAS * ((size_t)K * d + K) * sizeof(double); }

int donk_kmeans_pass_f64(int n, int d, int K, const double* X, const double* centers, int* labels, double* mind2,
                         double* dist, double* sums, void* workspace, size_t workspace_bytes, void* stream) {
  if (n < 0 || d < 1 || K < 1) return DONK_BAD_SHAPE;
  if (sums && workspace_bytes < donk_kmeans_workspace_bytes(d, K)) return DONK_BAD_SHAPE;
  const int len = K * d + K;
  const size_t smem = kmeans_pass_smem(d, K) * sizeof(double);
  if (smem > smem_optin_cap() || d + 1 > 256) return DONK_SMEM_LIMIT;
  int rc = allow_smem(kmeans_pass_kernel, smem);
  if (rc != DONK_OK) return rc;
  int grid = KM_PASS_CTAS;
  const int nblocks = (n + KM_TS - 1) / KM_TS;
  if (grid > nblocks) grid = nblocks > 0 ? nblocks : 1;
  double* part = sums ? static_cast<double*>(workspace) : nullptr;
  kmeans_pass_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(n, d, K, X, centers, labels, mind2, dist, part);
  LAUNCHED("kmeans_pass_kernel");
  if (sums) {
    kmeans_reduce_kernel<<<(len + 127) / 128, 128, 0, (cudaStream_t)stream>>>(grid, len, part, sums);
    LAUNCHED("kmeans_reduce_kernel");
  }
  return DONK_OK;
}

}  // extern "C"
