// donk_b200 — C ABI, part: GMM EM (GMMPrior.update, donk/dynamics/prior/prior_gmm.py:14-24).
#include "abi_common.cuh"
#include "gmm.cuh"
#include "gmm_mma.cuh"

using namespace donk;
using namespace donk_abi;

namespace {

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

__global__ void gmm_pfrag_kernel(int d, int K, int NB, size_t total, const double* pc, double* pf) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
    gmm_pfrag_item<double>(d, K, NB, pc, pf, idx);
}

__global__ void __launch_bounds__(32 * GMM_M18_WARPS, 1) gmm_m18_kernel(const GmmArgs<double> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  gmm_m18_cta((int)threadIdx.x, a, (int)blockIdx.x, reinterpret_cast<double*>(smem_raw));
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
  // fragment-major copy of the precision factors (stream-ordered scratch, 368 KB at d = 72, K = 16): the B fragments
  // of the products are then contiguous runs
  double* pf = nullptr;
  {
    const size_t len = gmm_pfrag_len(NB, a.K);
    keep_async_pool();
    if (cudaMallocAsync(reinterpret_cast<void**>(&pf), len * sizeof(double), (cudaStream_t)stream) != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "gmm E pass: cudaMallocAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
      return DONK_CUDA_ERROR;
    }
    gmm_pfrag_kernel<<<(unsigned)((len + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a.d, a.K, NB, len, a.pc, pf);
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  a.pf = pf;
  if (occ2) gmm_e_mma_kernel_occ2<NB, RA><<<grid, 32 * pl.Ge, smem, (cudaStream_t)stream>>>(a);
  else gmm_e_mma_kernel<NB, RA><<<grid, 32 * pl.Ge, smem, (cudaStream_t)stream>>>(a);
  a.pf = nullptr;
  if (pf) cudaFreeAsync(pf, (cudaStream_t)stream);
  LAUNCHED("gmm_e_mma_kernel");
  return DONK_OK;
}

int gmm_e_mma_dispatch(GmmArgs<double>& a, const GmmMmaPlan& pl, void* stream) {
  switch (pl.NB) {
    case 2: return gmm_e_mma_launch<2, 2>(a, pl, stream);
    case 5: return gmm_e_mma_launch<5, 2>(a, pl, stream);
    case 6: return gmm_e_mma_launch<6, 2>(a, pl, stream);
    case 9: return gmm_e_mma_launch<9, 2>(a, pl, stream);
    // d = 73..96 and 97..120: their own block counts (the 18-block kernel did up to twice the products on zero padding)
    case 12: return gmm_e_mma_launch<12, 2>(a, pl, stream);
    case 15: return gmm_e_mma_launch<15, 2>(a, pl, stream);
    // d = 73..144: 16 rows per warp (one 128-row tile, one CTA per SM) - every P_k fragment read through L1 feeds two
    // products instead of one: 49.0 -> 29.4 ms per pass at 2 M x 144, K = 16 (DONK_GMM_E_RA18=1: the 8-row variant)
    default: return env_int("DONK_GMM_E_RA18") == 1 ? gmm_e_mma_launch<18, 1>(a, pl, stream) : gmm_e_mma_launch<18, 2>(a, pl, stream);
  }
}

int gmm_m_mma_dispatch(GmmArgs<double>& a, const GmmMmaPlan& pl, void* stream) {
  // d = 137..144 (18 column blocks): nine warps, a pair of block rows each (gmm_m18_cta); same partials, same grid
  if ((a.d + 7) / 8 == 18 && pl.KC == GMM_M18_KC && a.d % 2 == 0 && (reinterpret_cast<size_t>(a.X) & 15) == 0 &&
      !env_int("DONK_GMM_M18_OFF")) {
    const size_t smem18 = gmm_m18_smem(a.d) * sizeof(double);
    if (smem18 <= smem_optin_cap()) {
      int rc18 = allow_smem(gmm_m18_kernel, smem18);
      if (rc18 != DONK_OK) return rc18;
      gmm_m18_kernel<<<((a.K + pl.KC - 1) / pl.KC) * a.nch, 32 * GMM_M18_WARPS, smem18, (cudaStream_t)stream>>>(a);
      LAUNCHED("gmm_m_mma_kernel");
      return DONK_OK;
    }
  }
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


}  // namespace

namespace donk_abi {

// responsibilities of n rows (sklearn predict_proba): the tensor-path E pass for pools of at least 4096 rows (the
// fit_lr prior at config-5 size evaluates millions of rows per call), the scalar kernel below that
// X == nullptr: the rows are the transitions of roll-outs tX (.., T+1, dX), tU (.., T, dU) (GmmArgs: transition view)
int gmm_predict_proba_launch(int n, int d, int K, const double* X, const double* weights, const double* means,
                             const double* prec_chol, double* resp, void* stream, const double* tX, const double* tU,
                             int tT, int tdX, int tdU) {
  if (n < 0 || d < 1 || K < 1) return DONK_BAD_SHAPE;
  if (X == nullptr && (tX == nullptr || tU == nullptr || tT < 1 || d != 2 * tdX + tdU)) return DONK_BAD_SHAPE;
  if (n == 0) return DONK_OK;
  GmmArgs<double> a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.d = d; a.K = K; a.X = X; a.w = weights; a.means = means; a.pc = prec_chol;
  a.resp = resp;
  if (X == nullptr) { a.tX = tX; a.tU = tU; a.tT = tT; a.tdX = tdX; a.tdU = tdU; }
  const GmmMmaPlan pl = gmm_mma_plan(n, d, K);
  if (pl.ok && n >= 4096 && !env_int("DONK_GMM_GENERIC")) return gmm_e_mma_dispatch(a, pl, stream);
  a.nblocks = (n + GMM_TS - 1) / GMM_TS;
  return gmm_e_launch(a, stream);
}

}  // namespace donk_abi

extern "C" {

size_t donk_gmm_stats_size(int d, int K) { return gmm_stats_len(d, K); }
size_t donk_gmm_workspace_bytes(int n_local, int d, int K) { return gmm_ws_reals(n_local, d, K) * sizeof(double); }

int donk_gmm_em_step_mode_f64(int n_local, int d, int K, const double* X, const double* weights, const double* means,
                              const double* prec_chol, const double* resp_in, double* stats, double* shift,
                              void* workspace, size_t workspace_bytes, int mode, void* stream) {
  if (n_local < 0 || d < 1 || K < 1 || mode < 0 || mode > 3) return DONK_BAD_SHAPE;
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
  const bool want_mma = mode == 2 || ((mode == 0 || mode == 3) && (n_local >= 65536 || env_int("DONK_GMM_MMA")));
  const bool mma = pl.ok && !env_int("DONK_GMM_GENERIC") && weights != nullptr && want_mma;
  if (mode == 3) {
    // responsibilities and per-block log-sum-exp totals are already in the workspace (donk_gmm_tc_estep_f32)
    if (resp_in) return DONK_BAD_SHAPE;
  } else if (resp_in) {
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
  return donk_abi::gmm_predict_proba_launch(n, d, K, X, weights, means, prec_chol, resp, stream);
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

}  // extern "C"
