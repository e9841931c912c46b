// Microbenchmark: issue interval and latency of mma.sync.m8n8k4.f64 on sm_100a, per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_issue dmma_issue.cu && ./dmma_issue
// For NACC independent accumulators and W warps per CTA (one CTA per SM), reports cycles per DMMA per warp.
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void k(int iters, double* out, long long* cyc) {
  double c[NACC][2];
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = -i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// DFMA: NACC independent chains
template <int NACC>
__global__ void kf(int iters, double* out, long long* cyc) {
  double c[NACC];
  double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 2e-3;
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int NACC>
void run(int warps) {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * sizeof(double)); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<NACC><<<148, 32 * warps>>>(iters, out, cyc);
  k<NACC><<<148, 32 * warps>>>(iters, out, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DMMA  warps/SM %2d  independent accumulators %2d : %6.1f cycles per DMMA per warp, %6.2f per SMSP\n", warps, NACC,
         (double)h / iters / NACC, (double)h / iters / NACC / ((warps + 3) / 4));
  kf<NACC><<<148, 32 * warps>>>(iters, out, cyc);
  kf<NACC><<<148, 32 * warps>>>(iters, out, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DFMA  warps/SM %2d  independent chains       %2d : %6.1f cycles per DFMA per warp, %6.2f per SMSP\n", warps, NACC,
         (double)h / iters / NACC, (double)h / iters / NACC / ((warps + 3) / 4));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int warps : {4, 8, 16}) {
    run<1>(warps); run<2>(warps); run<4>(warps); run<8>(warps); run<12>(warps);
  }
  return 0;
}
