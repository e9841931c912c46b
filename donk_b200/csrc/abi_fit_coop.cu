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
  double* scratch = nullptr;
  if (a.K > 0) {
    const int rcw = fit_prior_wts_launch(a, stream, &scratch);
    if (rcw != DONK_OK) return rcw;
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

// wts = mean_n predict_proba (prior_gmm.py:26-33) for all B T fits of the call: ONE E pass of the EM kernels over the
// B N T transitions, read straight from the roll-outs (transition view), in chunks of conditions whose responsibilities
// stay below ~1 GB of stream-ordered scratch, then the average over the N samples of each (condition, t).
int fit_prior_wts_launch(FitArgs<double>& a, void* stream, double** scratch_out) {
  cudaStream_t st = (cudaStream_t)stream;
  const int d = 2 * a.dX + a.dU;
  const size_t rows_per_cond = (size_t)a.N * a.T;
  const size_t per_cond = rows_per_cond * (size_t)a.K * sizeof(double);
  const size_t fit_in_1gb = ((size_t)1 << 30) / (per_cond ? per_cond : 1);
  const int chunk = fit_in_1gb < 1 ? 1 : (fit_in_1gb > (size_t)a.B ? a.B : (int)fit_in_1gb);
  if (rows_per_cond * chunk > (size_t)0x7fffffff) return DONK_BAD_SHAPE;
  const size_t wts_len = (size_t)a.B * a.T * a.K;
  const size_t need = (wts_len + (size_t)chunk * rows_per_cond * a.K) * sizeof(double);
  double* scratch = nullptr;
  keep_async_pool();
  if (cudaMallocAsync(reinterpret_cast<void**>(&scratch), need, st) != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "fit_lr: cudaMallocAsync(%zu) failed: %s", need, cudaGetErrorString(cudaGetLastError()));
    return DONK_CUDA_ERROR;
  }
  double* wts = scratch;
  double* resp = scratch + wts_len;
  for (int b0 = 0; b0 < a.B; b0 += chunk) {
    const int nB = a.B - b0 < chunk ? a.B - b0 : chunk;
    const size_t nrows = (size_t)nB * rows_per_cond;
    int rc = gmm_predict_proba_launch((int)nrows, d, a.K, nullptr, a.gw, a.gmeans, a.gpc, resp, stream,
                                      a.X + (size_t)b0 * a.N * (a.T + 1) * a.dX, a.U + (size_t)b0 * a.N * a.T * a.dU, a.T,
                                      a.dX, a.dU);
    if (rc != DONK_OK) { cudaFreeAsync(scratch, st); return rc; }
    const size_t nw = (size_t)nB * a.T * a.K;
    fit_wts_kernel<<<(unsigned)((nw + 127) / 128), 128, 0, st>>>(nB, a.N, a.T, a.K, resp, wts + (size_t)b0 * a.T * a.K);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (cudaError_t e_ = cudaGetLastError(); e_ != cudaSuccess) { cudaFreeAsync(scratch, st); return cuda_fail(e_, "fit_wts_kernel"); }
  }
  a.wts_in = wts;
  *scratch_out = scratch;
  return DONK_OK;
}

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
