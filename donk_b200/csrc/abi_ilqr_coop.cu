// donk_b200 — C ABI, part: CTA-per-problem iLQR step for large state dimensions (ilqr_coop.cuh; BASELINE config 5).
#include "abi_common.cuh"
#include "ilqr_coop.cuh"

using namespace donk;
using namespace donk_abi;

namespace {

template <typename R, int DX, int DU, bool FULL>
__global__ void __launch_bounds__(256, 1) ilqr_coop_kernel(const IlqrArgs<R> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WCtx w{(int)threadIdx.x, (int)blockDim.x, (int)threadIdx.x >> 5, (int)threadIdx.x & 31, (int)blockDim.x >> 5};
  ilqr_coop_cta<R, DX, DU, FULL>(w, a, (int)blockIdx.x, reinterpret_cast<R*>(smem_raw));
}
template <typename R, int DX, int DU, bool FULL>
__global__ void __launch_bounds__(512, 1) ilqr_coop_kernel_w16(const IlqrArgs<R> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WCtx w{(int)threadIdx.x, (int)blockDim.x, (int)threadIdx.x >> 5, (int)threadIdx.x & 31, (int)blockDim.x >> 5};
  ilqr_coop_cta<R, DX, DU, FULL>(w, a, (int)blockIdx.x, reinterpret_cast<R*>(smem_raw));
}

template <int DX, int DU, bool FULL>
int coop_launch_v(const IlqrArgs<double>& a, int G, size_t bytes, void* stream) {
  const unsigned grid = (unsigned)a.B * (unsigned)a.E;
  int rc;
  if (G <= 8) {
    rc = allow_smem(ilqr_coop_kernel<double, DX, DU, FULL>, bytes);
    if (rc != DONK_OK) return rc;
    ilqr_coop_kernel<double, DX, DU, FULL><<<grid, 32 * G, bytes, (cudaStream_t)stream>>>(a);
  } else {
    rc = allow_smem(ilqr_coop_kernel_w16<double, DX, DU, FULL>, bytes);
    if (rc != DONK_OK) return rc;
    ilqr_coop_kernel_w16<double, DX, DU, FULL><<<grid, 32 * G, bytes, (cudaStream_t)stream>>>(a);
  }
  LAUNCHED("ilqr_coop_kernel");
  return DONK_OK;
}

template <int DX, int DU>
int coop_launch(const IlqrArgs<double>& a, void* stream) {
  int G = env_int("DONK_COOP_WARPS");
  if (G <= 0) G = 8;
  if (G > 16) G = 16;
  const size_t bytes = ilqr_coop_bytes<double, DX, DU>();
  if (bytes > smem_optin_cap() || !coop_schedule_fits<DX, DU>(G)) return DONK_SMEM_LIMIT;
  return ilqr_coop_is_full<double, DX, DU>(a) ? coop_launch_v<DX, DU, true>(a, G, bytes, stream)
                                               : coop_launch_v<DX, DU, false>(a, G, bytes, stream);
}

}  // namespace

namespace donk_abi {

// returns -1 when (dX, dU) has no cooperative instantiation (the caller falls back to the generic kernel)
int ilqr_coop_try_launch(const IlqrArgs<double>& a, void* stream) {
#define DONK_TRY_COOP(DXV, DUV) \
  if (a.dX == DXV && a.dU == DUV) return coop_launch<DXV, DUV>(a, stream);
  DONK_COOP_DIMS(DONK_TRY_COOP)
#undef DONK_TRY_COOP
  return -1;
}

}  // namespace donk_abi

#if defined(DONK_PHASE_PROF)
// profiling builds only (profiles/prof_phases_coop.py): read / reset the per-phase cycle counters of the kernel above
extern "C" int donk_coop_phase_prof_read(unsigned long long* out32, int reset) {
  if (out32) cudaMemcpyFromSymbol(out32, donk::donk_coop_phase_cycles, 32 * sizeof(unsigned long long));
  if (reset) {
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(donk::donk_coop_phase_cycles, z, sizeof(z));
  }
  return DONK_OK;
}
#endif
