// donk_b200 — CUDA library: __global__ wrappers, launch configuration and the C ABI
// declared in include/donk_b200.h.  Built for sm_100a only (no other code path).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "../../include/donk_b200.h"
#include "fitlr.cuh"
#include "gmm.cuh"
#include "ilqr.cuh"

using namespace donk;

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

template <typename R>
int ilqr_step_launch(IlqrArgs<R> a, void* stream) {
  if (a.B < 0 || a.E < 1 || a.T < 1 || a.dX < 1 || a.dU < 1) return DONK_BAD_SHAPE;
  if ((a.flags & ILQR_KL) && (a.flags & (ILQR_BACKWARD | ILQR_FORWARD)) != (ILQR_BACKWARD | ILQR_FORWARD))
    return DONK_BAD_SHAPE;
  if (a.B == 0) return DONK_OK;
  IlqrPlan pl;
  int rc = ilqr_plan(a.B, a.E, a.dX, a.dU, sizeof(R), smem_optin_cap(), env_int("DONK_ILQR_SLOTS"),
                     env_int("DONK_ILQR_THREADS"), pl);
  if (rc != DONK_OK) {
    snprintf(g_err, sizeof(g_err), "ilqr_step: dX=%d dU=%d does not fit %zu bytes of shared memory", a.dX, a.dU,
             smem_optin_cap());
    return rc;
  }
  a.S = pl.S;
  a.ES = pl.ES;
  rc = allow_smem(ilqr_step_kernel<R>, pl.smem);
  if (rc != DONK_OK) return rc;
  ilqr_step_kernel<R><<<pl.grid, pl.threads, pl.smem, (cudaStream_t)stream>>>(a);
  LAUNCHED("ilqr_step_kernel");
  return DONK_OK;
}

template <typename R>
__global__ void extended_costs_kernel(int n, int dX, int dU, const R* K, const R* k, const R* Lam, R* C, R* c) {
  const int p = dX + dU;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * p) return;
  extended_costs_item<R>(idx / p, (int)(idx % p), dX, dU, K, k, Lam, C, c);
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
    return ilqr_step_launch<R>(a, stream);                                                                    \
  }
ILQR_STEP_IMPL(f64, double)
ILQR_STEP_IMPL(f32, float)

int donk_extended_costs_kl_f64(int n, int dX, int dU, const double* K, const double* k, const double* Lam,
                               double* C_kl, double* c_kl, void* stream) {
  if (n < 0 || dX < 1 || dU < 1) return DONK_BAD_SHAPE;
  if (n == 0) return DONK_OK;
  const size_t items = (size_t)n * (dX + dU);
  extended_costs_kernel<double><<<(unsigned)((items + 127) / 128), 128, 0, (cudaStream_t)stream>>>(n, dX, dU, K, k,
                                                                                                    Lam, C_kl, c_kl);
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

}  // extern "C"
