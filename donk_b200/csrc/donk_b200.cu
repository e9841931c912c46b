// donk_b200 — CUDA library, part: small helpers (extended costs, SPD factors, KL, Brent, transitions, state fit,
// k-means passes, FP64 probes) and the library-wide state.  Built for sm_100a only (no other code path).
// The C ABI is declared in include/donk_b200.h; the other parts are abi_ilqr.cu, abi_fit.cu, abi_gmm.cu.
#include "abi_common.cuh"
#include "fitlr.cuh"
#include "ilqr.cuh"
#include "kmeans.cuh"
#include "costs.cuh"

namespace donk_abi {
thread_local char g_err[256] = "";
std::atomic<unsigned long long> g_launches{0};
}  // namespace donk_abi

using namespace donk;
using namespace donk_abi;

namespace {

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
// cost constructors / transition log-probabilities (costs.cuh)
// ---------------------------------------------------------------------------------------------
__global__ void cost_l2_kernel(size_t total, int p, const double* t, const double* w, double* C, double* c, double* cc) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
    cost_l2_item<double>(p, t, w, C, c, cc, idx);
}
__global__ void cost_l1_kernel(size_t total, int p, const double* xu, const double* t, const double* w, double alpha,
                               double* C, double* c, double* cc) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
    cost_l1_item<double>(p, xu, t, w, alpha, C, c, cc, idx);
}
__global__ void dyn_log_prob_kernel(size_t total, int N, int T, int dX, int dU, const double* X, const double* U,
                                    const double* F, const double* f, const double* Sinv, const double* hld,
                                    double* out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* diff = reinterpret_cast<double*>(smem_raw) + (size_t)threadIdx.x * dX;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
    out[idx] = dyn_log_prob_item<double>(N, T, dX, dU, X, U, F, f, Sinv, hld, idx, diff);
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
template <int NTW>
__global__ void __launch_bounds__(256, 2) kmeans_pass_mma_kernel(int n, int d, int K, const double* X, const double* centers,
                                                                 int* labels, double* mind2, double* dist, double* part,
                                                                 const double* closest_in, double* pot_part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  kmeans_pass_mma_cta<NTW>((int)threadIdx.x, n, d, K, X, centers, labels, mind2, dist, part, (int)blockIdx.x, (int)gridDim.x,
                           reinterpret_cast<double*>(smem_raw), closest_in, pot_part);
}
// generic k-means++ round: clamp the distance matrix by the current nearest distances; per-block sums of the columns
__global__ void __launch_bounds__(256) kmeanspp_min_kernel(int n, int K, const double* closest, double* dmin, double* pot_part) {
  __shared__ double red[256];
  const int rows_per = (n + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * rows_per, r1 = min(n, r0 + rows_per);
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) kmeanspp_min_item<double>(K, closest, dmin, (size_t)r);
  __syncthreads();
  for (int k = 0; k < K; ++k) {
    double v = 0.0;
    for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) v += dmin[(size_t)r * K + k];
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) pot_part[(size_t)blockIdx.x * K + k] = red[0];
    __syncthreads();
  }
}
// tensor-path pass: two CTAs per SM when two tiles buffers fit, else one; returns -1 when the shape is outside the path
static int kmeans_mma_launch(int n, int d, int K, const double* X, const double* centers, int* labels, double* mind2,
                             double* dist, double* part, const double* closest, double* pot_part, int& grid, void* stream) {
  const size_t smem_mma = kmeans_mma_smem(d) * sizeof(double);
  if (!kmeans_mma_ok(d, K, X) || smem_mma > smem_optin_cap() || n < 4096 || env_int("DONK_KMEANS_SCALAR")) return -1;
  const int cap = 2 * (smem_mma + 1024) <= smem_optin_cap() ? 296 : 148;
  if (grid > cap) grid = cap;
  const int ntw = ((d + 1 + 7) / 8 + 7) / 8;
  int rc;
#define DONK_KM_CASE(V)                                                                                              \
  case V:                                                                                                            \
    rc = allow_smem(kmeans_pass_mma_kernel<V>, smem_mma);                                                            \
    if (rc != DONK_OK) return rc;                                                                                    \
    kmeans_pass_mma_kernel<V><<<grid, 256, smem_mma, (cudaStream_t)stream>>>(n, d, K, X, centers, labels, mind2, dist, part, \
                                                                            closest, pot_part);                       \
    break;
  switch (ntw) {
    DONK_KM_CASE(1) DONK_KM_CASE(2) DONK_KM_CASE(3)
    default: return -1;
  }
#undef DONK_KM_CASE
  return DONK_OK;
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

int donk_quadratic_cost_l2_f64(int n, int p, const double* t, const double* w, double* C, double* c, double* cc,
                               void* stream) {
  if (n < 0 || p < 1) return DONK_BAD_SHAPE;
  const size_t total = (size_t)n * p * p;
  if (total == 0) return DONK_OK;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cost_l2_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(total, p, t, w, C, c, cc);
  LAUNCHED("cost_l2_kernel");
  return DONK_OK;
}

int donk_quadratic_cost_l1_f64(int n, int p, const double* xu, const double* t, const double* w, double alpha,
                               double* C, double* c, double* cc, void* stream) {
  if (n < 0 || p < 1) return DONK_BAD_SHAPE;
  const size_t total = (size_t)n * p * p;
  if (total == 0) return DONK_OK;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cost_l1_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(total, p, xu, t, w, alpha, C, c, cc);
  LAUNCHED("cost_l1_kernel");
  return DONK_OK;
}

int donk_dynamics_log_prob_f64(int B, int N, int T, int dX, int dU, const double* X, const double* U, const double* F,
                               const double* f, const double* covar_inv, const double* half_logdet, double* log_prob,
                               void* stream) {
  if (B < 0 || N < 0 || T < 1 || dX < 1 || dU < 1) return DONK_BAD_SHAPE;
  const size_t total = (size_t)B * N * T;
  if (total == 0) return DONK_OK;
  const int threads = 64;
  const size_t smem = (size_t)threads * dX * sizeof(double);
  if (smem > smem_optin_cap()) return DONK_SMEM_LIMIT;
  int rc = allow_smem(dyn_log_prob_kernel, smem);
  if (rc != DONK_OK) return rc;
  size_t blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 32) blocks = 148 * 32;
  dyn_log_prob_kernel<<<(unsigned)blocks, threads, smem, (cudaStream_t)stream>>>(total, N, T, dX, dU, X, U, F, f,
                                                                               covar_inv, half_logdet, log_prob);
  LAUNCHED("dyn_log_prob_kernel");
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


enum { KM_PASS_CTAS = 148 * 3 };
size_t donk_kmeans_workspace_bytes(int d, int K) { return (size_t)KM_PASS_CTAS * ((size_t)K * d + K) * sizeof(double); }

int donk_kmeans_pass_f64(int n, int d, int K, const double* X, const double* centers, int* labels, double* mind2,
                         double* dist, double* sums, void* workspace, size_t workspace_bytes, void* stream) {
  if (n < 0 || d < 1 || K < 1) return DONK_BAD_SHAPE;
  if (sums && workspace_bytes < donk_kmeans_workspace_bytes(d, K)) return DONK_BAD_SHAPE;
  const int len = K * d + K;
  int grid = KM_PASS_CTAS;
  const int nblocks = (n + KM_TS - 1) / KM_TS;
  if (grid > nblocks) grid = nblocks > 0 ? nblocks : 1;
  double* part = sums ? static_cast<double*>(workspace) : nullptr;
  int rc = kmeans_mma_launch(n, d, K, X, centers, labels, mind2, dist, part, nullptr, nullptr, grid, stream);
  if (rc > 0) return rc;
  if (rc < 0) {
    const size_t smem = kmeans_pass_smem(d, K) * sizeof(double);
    if (smem > smem_optin_cap() || d + 1 > 256) return DONK_SMEM_LIMIT;
    rc = allow_smem(kmeans_pass_kernel, smem);
    if (rc != DONK_OK) return rc;
    kmeans_pass_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(n, d, K, X, centers, labels, mind2, dist, part);
  }
  LAUNCHED("kmeans_pass_kernel");
  if (sums) {
    kmeans_reduce_kernel<<<(len + 127) / 128, 128, 0, (cudaStream_t)stream>>>(grid, len, part, sums);
    LAUNCHED("kmeans_reduce_kernel");
  }
  return DONK_OK;
}


int donk_kmeanspp_round_f64(int n, int d, int Kc, const double* X, const double* cand, const double* closest, double* dmin,
                            double* pot, void* workspace, size_t workspace_bytes, void* stream) {
  if (n < 1 || d < 1 || Kc < 1 || Kc > 16) return DONK_BAD_SHAPE;
  if (workspace_bytes < donk_kmeans_workspace_bytes(d, Kc)) return DONK_BAD_SHAPE;
  double* pot_part = static_cast<double*>(workspace);  // (CTAs, Kc): fits the Lloyd workspace
  int grid = KM_PASS_CTAS;
  const int nblocks = (n + KM_TS - 1) / KM_TS;
  if (grid > nblocks) grid = nblocks;
  int rc = kmeans_mma_launch(n, d, Kc, X, cand, nullptr, nullptr, dmin, nullptr, closest, pot_part, grid, stream);
  if (rc > 0) return rc;
  if (rc == DONK_OK) {
    LAUNCHED("kmeans_pass_kernel");
  } else {
    rc = donk_kmeans_pass_f64(n, d, Kc, X, cand, nullptr, nullptr, dmin, nullptr, nullptr, 0, stream);
    if (rc != DONK_OK) return rc;
    if (grid > 296) grid = 296;
    kmeanspp_min_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, Kc, closest, dmin, pot_part);
    LAUNCHED("kmeanspp_min_kernel");
  }
  kmeans_reduce_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(grid, Kc, pot_part, pot);
  LAUNCHED("kmeans_reduce_kernel");
  return DONK_OK;
}

}  // extern "C"
