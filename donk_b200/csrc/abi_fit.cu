// donk_b200 — C ABI, part: fit_lr (donk/dynamics/linear_dynamics.py:137-189).
#include "abi_common.cuh"
#include "fitlr.cuh"
#include "fitlr_warp.cuh"

using namespace donk;
using namespace donk_abi;

namespace {

// ---------------------------------------------------------------------------------------------
// fit_lr, to_transitions, StateDistribution.fit
// ---------------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(512, 1) fit_lr_kernel(const FitArgs<R> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx ctx{(int)threadIdx.x, (int)blockDim.x};
  fit_lr_cta<R>(ctx, a, (int)blockIdx.x, reinterpret_cast<R*>(smem_raw));
}

#ifndef DONK_FIT_LB_CTAS
#define DONK_FIT_LB_CTAS 3  // 4-warp CTAs x 3 per SM: 170 registers per thread (experiments: profiles/build_alt.py)
#endif
template <typename R, int DX, int DU>
__global__ void __launch_bounds__(128, DONK_FIT_LB_CTAS) fit_warp_kernel(const FitArgs<R> a) {
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
  // With a prior of many clusters, calls of some size take the cluster weights of all fits from ONE E pass of the EM
  // kernels over the B N T transitions (fit_prior_wts_launch) instead of every warp walking the K precision factors for
  // its own N samples: at (20,6), N = 64, 4096 conditions, K = 16: 108 -> 72 ms; K = 4: 30.8 vs 33.7 ms, so few clusters
  // and small calls keep the single-launch in-kernel form (N <= 64 only).  DONK_FIT_PRIOR_INKERNEL=1 / =0 force either.
  bool pre_wts = false;
  if constexpr (sizeof(R) == 8) {
    if (a.K > 0) {
      const int force_in = env_int("DONK_FIT_PRIOR_INKERNEL", -1);
      pre_wts = a.N > 64 || (force_in < 0 ? (a.K >= 8 && (long long)a.B * a.N * a.T >= 32768) : force_in == 0);
    }
  } else {
    if (a.K > 0 && a.N > 64) return DONK_SMEM_LIMIT;  // single precision: in-kernel prior only
  }
  if ((long long)a.N * (a.T + 1) * DX >= (1LL << 31)) return DONK_SMEM_LIMIT;  // 32-bit row offsets in the moments pass
  int rc = fit_warp_plan<R, DX, DU>(smem_optin_cap() / 3 - 1024, forced, a.N, a.K, G, bytes, pre_wts);
  if (rc != DONK_OK) rc = fit_warp_plan<R, DX, DU>(smem_optin_cap(), forced, a.N, a.K, G, bytes, pre_wts);
  if (rc != DONK_OK) return rc;
  if (G > 4) {  // the kernel is compiled for <= 128 threads
    G = 4;
    bytes = (size_t)4 * (FitWarpDims<DX, DU>::TEAM + FitWarpDims<DX, DU>::prior_tail(a.N, a.K, pre_wts)) * sizeof(R);
  }
  rc = allow_smem(fit_warp_kernel<R, DX, DU>, bytes);
  if (rc != DONK_OK) return rc;
  double* wts_scratch = nullptr;
  if constexpr (sizeof(R) == 8) {
    if (pre_wts) {
      rc = fit_prior_wts_launch(a, stream, &wts_scratch);
      if (rc != DONK_OK) return rc;
    }
  }
  R* consts = nullptr;
  if (a.K > 0 && !pre_wts) {  // per-cluster constants of the log-probabilities: stream-ordered scratch, freed after the fit launch
    const int d = 2 * DX + DU, len = a.K * d + a.K;
    keep_async_pool();
    if (cudaMallocAsync(reinterpret_cast<void**>(&consts), (size_t)len * sizeof(R), (cudaStream_t)stream) != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "fit_lr: cudaMallocAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
      return DONK_CUDA_ERROR;
    }
    gmm_prior_consts_kernel<R><<<(len + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d, a.K, a.gw, a.gmeans, a.gpc, consts);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (cudaError_t e_ = cudaGetLastError(); e_ != cudaSuccess) {
      cudaFreeAsync(consts, (cudaStream_t)stream);
      return cuda_fail(e_, "gmm_prior_consts_kernel");
    }
    a.gconst = consts;
  }
  const long long nfits = (long long)a.B * a.T;
  a.no_prefetch = env_int("DONK_FIT_NO_PREFETCH");
  fit_warp_kernel<R, DX, DU><<<(unsigned)((nfits + G - 1) / G), 32 * G, bytes, (cudaStream_t)stream>>>(a);
  if (consts) cudaFreeAsync(consts, (cudaStream_t)stream);
  if (wts_scratch) cudaFreeAsync(wts_scratch, (cudaStream_t)stream);
  LAUNCHED("fit_warp_kernel");
  return DONK_OK;
}

}  // namespace

template <typename R>
static int fit_lr_launch(FitArgs<R>& a, void* stream) {
  const int B = a.B, N = a.N, T = a.T, dX = a.dX, dU = a.dU, gmm_K = a.K;
  // warp-per-fit fast path for the instantiated dimensions; a plan that does not fit shared memory (many clusters with
  // the in-kernel prior) falls through to the generic kernel
  if (!a.niw_Phi && !env_int("DONK_FIT_GENERIC")) {
#define DONK_TRY_FITW(DXV, DUV)                                          \
  if (dX == DXV && dU == DUV) {                                          \
    const int rc_w = fit_warp_launch<R, DXV, DUV>(a, stream);            \
    if (rc_w != DONK_SMEM_LIMIT) return rc_w;                            \
  }
    DONK_WARP_DIMS(DONK_TRY_FITW)
#undef DONK_TRY_FITW
  }
  // large dimensions: one fit per CTA on DMMA tiles (fitlr_coop.cuh); it flags the fits whose eigenvalue test needs the
  // eigen-solver (info == 2) and the generic kernel below redoes exactly those
  if constexpr (sizeof(R) == 8) {
    if (!env_int("DONK_FIT_GENERIC") && a.info != nullptr) {
      const int rc_c = fit_coop_try_launch(a, stream);
      a.wts_in = nullptr;  // stream-ordered scratch of the cooperative launch, already released
      if (rc_c == DONK_OK) a.redo = 1;
      else if (rc_c >= 0 && rc_c != DONK_SMEM_LIMIT) return rc_c;
    }
  }
  size_t bytes = 0;
  const int force = env_int("DONK_FIT_CHUNK");
  a.NS = fit_plan(dX, dU, gmm_K, force > 0 ? force : N, sizeof(R), smem_optin_cap(), bytes);
  if (a.NS == 0) {
    snprintf(g_err, sizeof(g_err), "fit_lr: dX=%d dU=%d K=%d does not fit shared memory", dX, dU, gmm_K);
    return DONK_SMEM_LIMIT;
  }
  { FitLayout L(dX, dU, gmm_K, N, a.NS); bytes = (size_t)L.total * sizeof(R); }
  int rc = allow_smem(fit_lr_kernel<R>, bytes);
  if (rc != DONK_OK) return rc;
  const int dp = round_up(2 * dX + dU, 4), nt = dp / 4;
  int threads = round_up(nt * (nt + 1) / 2, 32);
  if (threads < 128) threads = 128;
  if (threads > 512) threads = 512;
  if (env_int("DONK_FIT_THREADS") > 0) threads = env_int("DONK_FIT_THREADS");
  fit_lr_kernel<R><<<B * T, threads, bytes, (cudaStream_t)stream>>>(a);
  LAUNCHED("fit_lr_kernel");
  return DONK_OK;
}

extern "C" {

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
  a.aligned16 = (((uintptr_t)X | (uintptr_t)U) & 15) == 0;
  return fit_lr_launch<double>(a, stream);
}

/* single-precision storage and arithmetic (the tensor-path products still accumulate in FP64): same kernels, R = float */
int donk_fit_lr_f32(int B, int N, int T, int dX, int dU, const float* X, const float* U, int gmm_K,
                    const float* gmm_weights, const float* gmm_means, const float* gmm_covariances,
                    const float* gmm_prec_chol, float n_pool, float prior_weight, float regularization,
                    float* F, float* f, float* covar, int* info, void* stream) {
  if (B < 0 || N <= 1 || T < 1 || dX < 1 || dU < 1 || gmm_K < 0) return DONK_BAD_SHAPE;
  if (B == 0) return DONK_OK;
  FitArgs<float> a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.N = N; a.T = T; a.dX = dX; a.dU = dU; a.X = X; a.U = U; a.K = gmm_K;
  a.gw = gmm_weights; a.gmeans = gmm_means; a.gcov = gmm_covariances; a.gpc = gmm_prec_chol;
  a.n_pool = n_pool; a.prior_weight = prior_weight; a.reg = regularization;
  a.F = F; a.f = f; a.covar = covar; a.info = info;
  a.aligned16 = (((uintptr_t)X | (uintptr_t)U) & 15) == 0;
  return fit_lr_launch<float>(a, stream);
}

int donk_fit_lr_niw_f64(int B, int N, int T, int dX, int dU, const double* X, const double* U, const double* niw_mu0,
                        const double* niw_Phi, const double* niw_N_mean, const double* niw_N_covar, double prior_weight,
                        double regularization, double* F, double* f, double* covar, int* info, void* stream) {
  if (B < 0 || N <= 1 || T < 1 || dX < 1 || dU < 1 || !niw_mu0 || !niw_Phi || !niw_N_mean || !niw_N_covar)
    return DONK_BAD_SHAPE;
  if (B == 0) return DONK_OK;
  FitArgs<double> a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.N = N; a.T = T; a.dX = dX; a.dU = dU; a.X = X; a.U = U; a.K = 0;
  a.niw_mu0 = niw_mu0; a.niw_Phi = niw_Phi; a.niw_Nmean = niw_N_mean; a.niw_Ncovar = niw_N_covar;
  a.prior_weight = prior_weight; a.reg = regularization;
  a.F = F; a.f = f; a.covar = covar; a.info = info;
  a.aligned16 = (((uintptr_t)X | (uintptr_t)U) & 15) == 0;
  return fit_lr_launch<double>(a, stream);
}

}  // extern "C"
