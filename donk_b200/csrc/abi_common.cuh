// donk_b200 — pieces shared by the translation units of the CUDA library (error reporting, launch counter,
// shared-memory opt-in).  The library is built from several .cu files so that nvcc compiles them in parallel.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <cstdint>

#include "../../include/donk_b200.h"
#include "common.cuh"

// (dX, dU) pairs with warp-kernel instantiations (ILQR step, fit_lr): the BASELINE.json configurations that
// fit one warp per problem; other dimensions run the CTA-cooperative or the generic kernels.
#define DONK_WARP_DIMS(X) X(20, 6) X(14, 7) X(6, 2)

// (dX, dU) pairs with CTA-per-problem iLQR instantiations (ilqr_coop.cuh): BASELINE config 5 and a small shape
// that keeps the oracle comparisons of the test-suite short.
#define DONK_COOP_DIMS(X) X(64, 16) X(16, 8)

namespace donk_abi {

extern thread_local char g_err[256];
extern std::atomic<unsigned long long> g_launches;

inline int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return DONK_CUDA_ERROR;
}
#define CU(call)                                                  \
  do {                                                            \
    cudaError_t e_ = (call);                                      \
    if (e_ != cudaSuccess) return donk_abi::cuda_fail(e_, #call); \
  } while (0)
#define LAUNCHED(name)                                                    \
  do {                                                                    \
    donk_abi::g_launches.fetch_add(1, std::memory_order_relaxed);         \
    cudaError_t e_ = cudaGetLastError();                                  \
    if (e_ != cudaSuccess) return donk_abi::cuda_fail(e_, name);          \
  } while (0)

inline int env_int(const char* name, int unset = 0) {
  const char* v = getenv(name);
  return v ? atoi(v) : unset;
}

inline size_t smem_optin_cap() {
  static size_t cap = 0;
  if (!cap) {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cap = v > 0 ? (size_t)v : 48 * 1024;
  }
  return cap;
}

// Stream-ordered scratch (cudaMallocAsync) comes from the device's default memory pool.  Its release threshold is 0 by
// default: every synchronisation hands the freed blocks back to the driver and the next call pays for a fresh
// allocation (~100 ms per GB).  Keep what was freed in the pool instead (set once per device).
inline void keep_async_pool() {
  static bool done[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done[dev] = true;
}

template <typename KernelT>
int allow_smem(KernelT kernel, size_t bytes) {
  if (bytes > 48 * 1024) CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  // ask for the full shared-memory carve-out so that several CTAs fit one SM
  CU(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  return DONK_OK;
}

}  // namespace donk_abi

namespace donk {
template <typename R>
struct IlqrArgs;
}
namespace donk {
template <typename R>
struct FitArgs;
}
namespace donk_abi {
int ilqr_coop_try_launch(const donk::IlqrArgs<double>& a, void* stream);  // abi_ilqr_coop.cu; -1: no instantiation
int fit_coop_try_launch(donk::FitArgs<double>& a, void* stream);          // abi_fit_coop.cu; -1: no instantiation
int gmm_predict_proba_launch(int n, int d, int K, const double* X, const double* weights, const double* means,
                             const double* prec_chol, double* resp, void* stream, const double* tX = nullptr,
                             const double* tU = nullptr, int tT = 0, int tdX = 0, int tdU = 0);  // abi_gmm.cu
// cluster weights of the GMM prior for every fit of a call (abi_fit_coop.cu): sets a.wts_in to stream-ordered scratch the
// caller releases with cudaFreeAsync(*scratch) after the fit launch
int fit_prior_wts_launch(donk::FitArgs<double>& a, void* stream, double** scratch);
}
