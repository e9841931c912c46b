// Microbenchmark: cycles of the warp-collective block products of the iLQR step kernel, using the library's own
// warp_mma (common.cuh), for the shapes at (dX,dU) = (20,6), with 1, 2 or 4 warps per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I donk_b200/csrc -o mma_shapes mma_shapes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace donk;

template <int RA, int RB, int K4, int LDA, int LDB, int MODE>
__global__ void k(int iters, double* out, long long* cyc) {
  extern __shared__ __align__(16) double sm[];
  WCtx w{(int)threadIdx.x, (int)blockDim.x, (int)threadIdx.x >> 5, (int)threadIdx.x & 31, (int)blockDim.x >> 5};
  double* A = sm + w.team * 1664;
  double* B = A + 100;   // operands may overlap: read-only
  double* C = A + 800;
  for (int i = w.lane; i < 1664; i += 32) A[i] = 1e-3 * i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      warp_mma<double, RA, RB>(w, A, LDA, B, LDB, K4, [&](int i, int j, double v0, double v1) {
        st2(C + i * 26 + j, v0, v1);
      });
    } else {  // row-block calls (RA = 1 each), as in the triangular products
#pragma unroll
      for (int r = 0; r < RA; ++r)
        warp_mma<double, 1, RB>(w, A + 8 * r, LDA, B, LDB, K4, [&](int i, int j, double v0, double v1) {
          st2(C + (8 * r + i) * 26 + j, v0, v1);
        });
    }
    w.team_sync();
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = C[w.lane];
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int RA, int RB, int K4, int LDA, int LDB, int MODE>
void run(const char* name) {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * sizeof(double)); cudaMalloc(&cyc, 8);
  const int iters = 500;
  for (int warps : {4, 8, 16}) {
    auto kern = k<RA, RB, K4, LDA, LDB, MODE>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1664 * 8);
    kern<<<148, 32 * warps, warps * 1664 * 8>>>(iters, out, cyc);
    kern<<<148, 32 * warps, warps * 1664 * 8>>>(iters, out, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    if (cudaGetLastError() != cudaSuccess) printf("launch error\n");
    const int nd = RA * RB * K4;
    printf("%-28s warps/SMSP %d: %7.0f cycles per product (%3d DMMA -> ideal %5d), %5.1f cycles/DMMA/SMSP\n", name, warps / 4,
           (double)h / iters, nd, nd * 16 * (warps / 4), (double)h / iters / nd / (warps / 4));
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<3, 4, 5, 28, 28, 0>("B1 3x4 k5 (one call)");
  run<3, 4, 5, 28, 28, 1>("B1 3x(1x4) k5 (row calls)");
  run<1, 4, 5, 28, 28, 0>("1x4 k5");
  run<1, 2, 5, 28, 28, 0>("1x2 k5");
  run<1, 1, 5, 28, 28, 0>("1x1 k5");
  run<4, 3, 7, 28, 20, 0>("F3 4x3 k7 (one call)");
  run<1, 3, 7, 20, 20, 0>("1x3 k7");
  run<1, 3, 2, 28, 28, 0>("B5 1x3 k2");
  return 0;
}
