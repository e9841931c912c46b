// Microbenchmark: does a stream of mma.sync.m8n8k4.f64 on one warp slow the instruction issue of ANOTHER warp on the
// same SM sub-partition?  CTA of 8 warps (two per sub-partition): warps 0-3 run DMMA (12 independent accumulators),
// warps 4-7 run a second stream (FP32 FMA / integer / shared-memory loads / DFMA).  Reports cycles of each stream
// alone and together.  If "together" ~ max(alone) the pipes overlap; if ~ sum, DMMA holds the sub-partition's dispatch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_coissue dmma_coissue.cu && ./dmma_coissue
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// mode bit 0: warps 0-3 run DMMA; bits 1..: what warps 4-7 run (0 nothing, 1 FFMA, 2 IMAD, 3 LDS.64, 4 DFMA)
__global__ void k(int iters, int do_dmma, int other, double* out, long long* cyc) {
  __shared__ double sm[2048];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i * 1e-3;
  __syncthreads();
  long long t0 = 0, t1 = 0;
  double acc = 0;
  if (warp < 4) {
    if (do_dmma) {
      double c[12][2];
      const double a = lane * 1e-3, b = lane * 2e-3;
#pragma unroll
      for (int i = 0; i < 12; ++i) { c[i][0] = i; c[i][1] = -i; }
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 12; ++i) dmma(c[i][0], c[i][1], a, b);
      }
      t1 = clock64();
#pragma unroll
      for (int i = 0; i < 12; ++i) acc += c[i][0] + c[i][1];
    }
  } else if (other == 1) {
    float f[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) f[i] = i + lane;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 12; ++i) f[i] = fmaf(f[i], 1.0001f, 0.5f);
    }
    t1 = clock64();
#pragma unroll
    for (int i = 0; i < 12; ++i) acc += f[i];
  } else if (other == 2) {
    int f[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) f[i] = i + lane;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 12; ++i) f[i] = f[i] * 3 + it;
    }
    t1 = clock64();
#pragma unroll
    for (int i = 0; i < 12; ++i) acc += f[i];
  } else if (other == 3) {
    double f[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) f[i] = 0;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 12; ++i) f[i] += sm[(lane + 32 * i + 7 * r + it) & 2047];
    }
    t1 = clock64();
#pragma unroll
    for (int i = 0; i < 12; ++i) acc += f[i];
  } else if (other == 4) {
    double f[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) f[i] = i + lane;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 12; ++i) f[i] = fma(f[i], 1.0000001, 0.5);
    }
    t1 = clock64();
#pragma unroll
    for (int i = 0; i < 12; ++i) acc += f[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (blockIdx.x == 0 && lane == 0) cyc[warp] = t1 - t0;
}

int main() {
  double* out; long long* cyc; long long h[8];
  cudaMalloc(&out, 148 * 256 * sizeof(double)); cudaMalloc(&cyc, 64);
  const int iters = 2000;
  const char* names[] = {"nothing", "FFMA x96", "IMAD x96", "LDS.64 x96 (+DADD)", "DFMA x12"};
  for (int other = 0; other <= 4; ++other)
    for (int dm = 0; dm <= 1; ++dm) {
      if (!dm && !other) continue;
      for (int rep = 0; rep < 2; ++rep) k<<<148, 256>>>(iters, dm, other, out, cyc);
      cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
      printf("DMMA warps 0-3: %-3s  other warps 4-7: %-20s  -> cycles per iteration: DMMA warp %7.1f (12 DMMA = 192 at peak), other warp %7.1f\n",
             dm ? "on" : "off", names[other], dm ? (double)h[0] / iters : 0.0, other ? (double)h[4] / iters : 0.0);
    }
  return 0;
}
