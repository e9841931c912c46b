// donk_b200 — C ABI, part: CTA-per-fit fit_lr for large dimensions (fitlr_coop.cuh; BASELINE config 5) and the
// cluster weights of the GMM prior computed by the E pass of the EM kernels over all samples of the call.
#include "abi_common.cuh"
#include "fitlr_coop.cuh"

using namespace donk;
using namespace donk_abi;

namespace {

template <typename R, int DX, int DU>
__global__ void __launch_bounds__(256, 1) fit_coop_kernel(const FitArgs<R> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WCtx w{(int)threadIdx.x, (int)blockDim.x, (int)threadIdx.x >> 5, (int)threadIdx.x & 31, (int)blockDim.x >> 5};
  fit_coop_cta<R, DX, DU>(w, a, (int)blockIdx.x, (int)gridDim.x, reinterpret_cast<R*>(smem_raw));
}

// rows [x_t | u_t | x_{t+1}] of nB conditions, row index (b N + n) T + t (donk/dynamics/utils.py:4-17)
__global__ void fit_gather_kernel(size_t total, int T, int dX, int dU, const double* X, const double* U, double* out) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
    out[idx] = transitions_item<double>(T, dX, dU, X, U, idx);
}

// wts[b][t][k] = mean_n resp[((b N + n) T + t) K + k]  (prior_gmm.py:33)
__global__ void fit_wts_kernel(int nB, int N, int T, int K, const double* resp, double* wts) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)nB * T * K) return;
  const int k = (int)(idx % K);
  const size_t bt = idx / K;
  const int t = (int)(bt % T);
  const size_t b = bt / T;
  double acc = 0.0;
  for (int n = 0; n < N; ++n) acc += resp[((b * N + n) * T + t) * K + k];
  wts[idx] = acc / N;
}

template <int DX, int DU>
int fit_coop_launch(FitArgs<double>& a, void* stream) {
  const size_t bytes = fit_coop_bytes<double, DX, DU>();
  if (bytes > smem_optin_cap()) return DONK_SMEM_LIMIT;
  cudaStream_t st = (cudaStream_t)stream;
  const int d = 2 * DX + DU;
  double* scratch = nullptr;
  if (a.K > 0) {
    // cluster weights of every fit: gather the samples of a chunk of conditions as rows, run the E pass, average over
    // the N samples of each (condition, t).  Stream-ordered scratch, bounded at ~1 GB.
    const size_t rows_per_cond = (size_t)a.N * a.T;
    const size_t per_cond = rows_per_cond * (size_t)(d + a.K) * sizeof(double);
    const size_t fit_in_1gb = ((size_t)1 << 30) / (per_cond ? per_cond : 1);
    const int chunk = fit_in_1gb < 1 ? 1 : (fit_in_1gb > (size_t)a.B ? a.B : (int)fit_in_1gb);
    if (rows_per_cond * chunk > (size_t)0x7fffffff) return DONK_BAD_SHAPE;
    const size_t wts_len = (size_t)a.B * a.T * a.K;
    const size_t need = (wts_len + (size_t)chunk * rows_per_cond * (d + a.K)) * sizeof(double);
    keep_async_pool();
    if (cudaMallocAsync(reinterpret_cast<void**>(&scratch), need, st) != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "fit_lr: cudaMallocAsync(%zu) failed: %s", need, cudaGetErrorString(cudaGetLastError()));
      return DONK_CUDA_ERROR;
    }
    double* wts = scratch;
    double* rows = scratch + wts_len;
    double* resp = rows + (size_t)chunk * rows_per_cond * d;
    for (int b0 = 0; b0 < a.B; b0 += chunk) {
      const int nB = a.B - b0 < chunk ? a.B - b0 : chunk;
      const size_t nrows = (size_t)nB * rows_per_cond, total = nrows * d;
      size_t blocks = (total + 255) / 256;
      if (blocks > 148 * 32) blocks = 148 * 32;
      fit_gather_kernel<<<(unsigned)blocks, 256, 0, st>>>(total, a.T, DX, DU, a.X + (size_t)b0 * a.N * (a.T + 1) * DX,
                                                           a.U + (size_t)b0 * a.N * a.T * DU, rows);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      int rc = gmm_predict_proba_launch((int)nrows, d, a.K, rows, a.gw, a.gmeans, a.gpc, resp, stream);
      if (rc != DONK_OK) { cudaFreeAsync(scratch, st); return rc; }
      const size_t nw = (size_t)nB * a.T * a.K;
      fit_wts_kernel<<<(unsigned)((nw + 127) / 128), 128, 0, st>>>(nB, a.N, a.T, a.K, resp, wts + (size_t)b0 * a.T * a.K);
      g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    a.wts_in = wts;
  }
  int rc = allow_smem(fit_coop_kernel<double, DX, DU>, bytes);
  if (rc == DONK_OK) {
    const long long nfits = (long long)a.B * a.T;
    const int grid = nfits < 148 ? (int)nfits : 148;  // persistent CTAs, one per SM
    fit_coop_kernel<double, DX, DU><<<grid, 256, bytes, st>>>(a);
  }
  if (scratch) cudaFreeAsync(scratch, st);
  if (rc != DONK_OK) return rc;
  LAUNCHED("fit_coop_kernel");
  return DONK_OK;
}

}  // namespace

namespace donk_abi {

// returns -1 when (dX, dU) has no cooperative instantiation.  Fits the kernel could not finish (the Cholesky attempt of
// the eigenvalue test failed: info == 2) are left to the caller, which reruns the generic kernel on them.
int fit_coop_try_launch(FitArgs<double>& a, void* stream) {
#define DONK_TRY_FITC(DXV, DUV) \
  if (a.dX == DXV && a.dU == DUV) return fit_coop_launch<DXV, DUV>(a, stream);
  DONK_COOP_DIMS(DONK_TRY_FITC)
#undef DONK_TRY_FITC
  return -1;
}

}  // namespace donk_abi
