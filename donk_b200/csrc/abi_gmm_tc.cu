// donk_b200 — C ABI, part: the E pass of GMM EM (GMMPrior.update, donk/dynamics/prior/prior_gmm.py:14-24 ->
// sklearn GaussianMixture._e_step) at the north star's "stated fp32 tolerance" on the 5th-generation tensor cores.
//
// There is no FP64 tcgen05 kind: the fp64 EM kernels (gmm_mma.cuh) run on mma.sync.m8n8k4.f64 at the DFMA rate.  This
// variant computes y = (x - c) P_k - (mu_k - c) P_k for all rows x and clusters k as ONE GEMM on tcgen05.mma kind::tf32
// with fp32 accumulators in tensor memory, each operand split into two TF32 planes (x = x_hi + x_lo) and three products
// (hi*hi + lo*hi + hi*lo: ~2^-21 relative, the "3xTF32" scheme); log-probabilities are reduced from the accumulators
// straight out of TMEM.  Data flow of one CTA (persistent, one per SM, 192 threads):
//     warp 0     one lane: TMA (cp.async.bulk.tensor.2d, 32-byte swizzle) P planes once, then the row tiles of X into a
//                ring of shared-memory stages (full / empty mbarriers)
//     warp 1     one lane: tcgen05.mma M=128, N = kc*dp, K=8 per instruction, 3 * dp/8 instructions per tile into one of
//                two TMEM accumulators; tcgen05.commit releases the stage and publishes the accumulator
//     warps 2-5  tcgen05.ld of their 32 TMEM lanes (= rows), y^2 row sums, log-probability per (row, cluster) to HBM
// A CTA owns kc clusters (their P planes stay resident in shared memory) and walks the row tiles; the 128-row tile of X
// is read by K/kc CTAs at about the same time (one HBM read, the rest L2 hits).  A second small kernel turns the
// (n, K) log-probabilities into responsibilities (fp64, where the fp64 M pass expects them) and the per-block
// log-sum-exp totals of the lower bound.  X is split into planes ONCE per fit (it does not change over the iterations).
#include <cuda.h>

#include "abi_common.cuh"
#include "gmm.cuh"

using namespace donk;
using namespace donk_abi;

namespace {

constexpr int TC_M = 128;        // rows per tile = TMEM lanes
constexpr int TC_THREADS = 320;  // TMA warp, MMA warp, eight epilogue warps (two per TMEM lane quarter)
constexpr int TC_MAX_STAGES = 6;
// A row tile of X enters shared memory in chunks of up to TC_KC k-steps (8 features each): next to the resident P planes
// of two clusters (83 KB at d = 72) a whole 74 KB tile leaves room for ONE stage - load and products then alternate
// (11.0 ms per pass at 10 M x 72); chunks of 3 k-steps give a ring of 5-6 stages
constexpr int TC_KC = 3;
__host__ __device__ constexpr int tc_kc(int KS) { return KS > TC_KC + 1 ? TC_KC : KS; }
constexpr int TC_TMEM_COLS = 512;

struct TcParams {
  int n, d, dp, K, kc, ncg, walkers, ntiles, stages, N;
  const float* bias;  // (K, dp)  (mu_k - c) P_k
  const float* cst;   // (K)      sum log diag P_k + log pi_k - d/2 log(2 pi)
  float* logp;        // (n, K)
  const unsigned char* ximg;  // [ntiles][plane hi|lo][k-step][128 rows][32 B, swizzled]: the A stage images
};

// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major operand tile of 8-float (32-byte) rows written by TMA with the 32-byte swizzle: rows 32 bytes apart,
// groups of 8 rows 256 bytes apart (stride byte offset); the leading byte offset is unused for swizzled K-major tiles.
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version of sm_100
  d |= (uint64_t)6 << 61;  // SWIZZLE_32B
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
      ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
// 8 consecutive TMEM columns of this thread's lane; the registers are valid after tcgen05.wait::ld
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

template <int KS>  // k-steps of 8 features: dp = 8 KS (compile-time so that a cluster's columns are read from TMEM in one batch)
__global__ void __launch_bounds__(TC_THREADS, 1)
gmm_tc_estep_kernel(const __grid_constant__ CUtensorMap tmPhi, const __grid_constant__ CUtensorMap tmPlo, const TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // the swizzle pattern is a function of the shared-memory address bits: operand tiles start on 256-byte boundaries
  unsigned char* smem = smem_raw + ((256u - (smem_u32(smem_raw) & 255u)) & 255u);
  const int N = p.N;
  constexpr int KC = tc_kc(KS), NCH = (KS + KC - 1) / KC;  // k-steps per chunk, chunks per tile
  constexpr uint32_t kstep_b = TC_M * 32;                      // bytes of one k-step of one plane of the tile
  constexpr uint32_t a_stage = 2 * KC * kstep_b, a_tile = 2 * KS * kstep_b;
  const uint32_t b_plane = (uint32_t)KS * N * 32;
  unsigned char* sB = smem;                    // [plane hi|lo][k-step][N rows][32 B]
  unsigned char* sA = sB + 2 * b_plane;        // [stage][plane hi|lo][k-step of the chunk][128 rows][32 B]
  unsigned char* tail = sA + (size_t)p.stages * a_stage;
  unsigned long long* full = reinterpret_cast<unsigned long long*>(tail);  // [stages]
  unsigned long long* empty = full + TC_MAX_STAGES;                        // [stages]
  unsigned long long* bfull = empty + TC_MAX_STAGES;                       // [1]
  unsigned long long* tfull = bfull + 1;                                   // [2]
  unsigned long long* tempty = tfull + 2;                                  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sbias = reinterpret_cast<float*>(tmem_slot + 6);                  // [kc][dp], NEGATED, 16-byte aligned
  float* scst = sbias + p.kc * p.dp;                                       // [kc]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = blockIdx.x % p.ncg, wk = blockIdx.x / p.ncg;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(bfull, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.kc * p.dp; i += TC_THREADS) sbias[i] = -p.bias[(size_t)cg * p.kc * p.dp + i];
  if (threadIdx.x < p.kc) scst[threadIdx.x] = p.cst[cg * p.kc + threadIdx.x];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bfull, 2 * b_plane);
      for (int kk = 0; kk < KS; ++kk) {
        tma_load_2d(sB + (size_t)kk * N * 32, &tmPhi, 8 * kk, cg * N, bfull);
        tma_load_2d(sB + b_plane + (size_t)kk * N * 32, &tmPlo, 8 * kk, cg * N, bfull);
      }
      int s = 0;
      uint32_t ph = 0;
      for (int tile = wk; tile < p.ntiles; tile += p.walkers) {
        // the 128-row tile is stored in global memory as the images of its stages (chunk major; per chunk both planes,
        // k-step major, 32-byte swizzle applied by the split kernel): one contiguous bulk copy per chunk instead of
        // 2 KS boxes of 128 32-byte rows
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          constexpr uint32_t full_b = a_stage;
          const uint32_t bytes = ch + 1 < NCH ? full_b : 2 * (KS - (NCH - 1) * KC) * kstep_b;
          mbar_wait(empty + s, ph ^ 1);
          mbar_arrive_expect_tx(full + s, bytes);
          bulk_copy_g2s(sA + (size_t)s * a_stage, p.ximg + (size_t)tile * a_tile + (size_t)ch * a_stage, bytes, full + s);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // the whole warp walks the loop converged, one elected lane issues (addresses stay in uniform registers)
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    mbar_wait(bfull, 0);
    const uint64_t bd_hi = umma_desc_sw32(smem_u32(sB)), bd_lo = bd_hi + (uint64_t)(b_plane >> 4);
    const uint32_t bstep = ((uint32_t)N * 32u) >> 4, astep = ((uint32_t)TC_M * 32u) >> 4;
    int s = 0, acc = 0;
    uint32_t ph = 0, aph = 0;
    for (int tile = wk; tile < p.ntiles; tile += p.walkers) {
      mbar_wait(tempty + acc, aph ^ 1);
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N);
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        constexpr int k0 = 0;
        const int len = ch + 1 < NCH ? KC : KS - (NCH - 1) * KC;  // compile-time after unrolling
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint64_t ad_hi = umma_desc_sw32(smem_u32(sA + (size_t)s * a_stage)), ad_lo = ad_hi + (uint64_t)((len * kstep_b) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {  // x_hi P_hi + x_lo P_hi + x_hi P_lo
            const uint64_t ad = pass == 1 ? ad_lo : ad_hi, bd = (pass == 2 ? bd_lo : bd_hi) + (uint64_t)((ch * KC) * bstep);
#pragma unroll
            for (int kk = k0; kk < len; ++kk)
              umma_tf32(d_tmem, ad + (uint64_t)(kk * astep), bd + (uint64_t)(kk * bstep), idesc, (ch | pass | kk) ? 1u : 0u);
          }
          tc_commit(empty + s);                        // the stage may be refilled once these products have read it
          if (ch + 1 == NCH) tc_commit(tfull + acc);   // the accumulator is complete
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  } else {
    const int q = warp & 3;  // a warp reads the TMEM lanes 32 (warp % 4) .. + 31
    int acc = 0;
    uint32_t aph = 0;
    for (int tile = wk; tile < p.ntiles; tile += p.walkers) {
      mbar_wait(tfull + acc, aph);
      tc_fence_after();
      const size_t row = (size_t)tile * TC_M + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * N);
      for (int c = (warp - 2) >> 2; c < p.kc; c += 2) {  // the two warps of a lane quarter take alternate clusters
        // all 8 KS columns of the cluster in flight, one wait: the loads are latency-bound when waited for one by one
        uint32_t v[KS][8];
#pragma unroll
        for (int ch = 0; ch < KS; ++ch) tmem_ld8_nowait(taddr + (uint32_t)(c * 8 * KS + 8 * ch), v[ch]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float2 ss = make_float2(0.f, 0.f);  // packed fp32 pairs: y = acc - bias, ss += y * y
        const float4* nb = reinterpret_cast<const float4*>(sbias + c * 8 * KS);
#pragma unroll
        for (int ch = 0; ch < KS; ++ch) {
          const float4 b0 = nb[2 * ch], b1 = nb[2 * ch + 1];
          const float2 y0 = __fadd2_rn(make_float2(__uint_as_float(v[ch][0]), __uint_as_float(v[ch][1])), make_float2(b0.x, b0.y));
          const float2 y1 = __fadd2_rn(make_float2(__uint_as_float(v[ch][2]), __uint_as_float(v[ch][3])), make_float2(b0.z, b0.w));
          const float2 y2 = __fadd2_rn(make_float2(__uint_as_float(v[ch][4]), __uint_as_float(v[ch][5])), make_float2(b1.x, b1.y));
          const float2 y3 = __fadd2_rn(make_float2(__uint_as_float(v[ch][6]), __uint_as_float(v[ch][7])), make_float2(b1.z, b1.w));
          ss = __ffma2_rn(y0, y0, ss); ss = __ffma2_rn(y1, y1, ss); ss = __ffma2_rn(y2, y2, ss); ss = __ffma2_rn(y3, y3, ss);
        }
        // layout [cluster group][row][kc]: the 32 rows of a warp are one contiguous run
        if (row < (size_t)p.n) p.logp[((size_t)cg * p.n + row) * p.kc + c] = scst[c] - 0.5f * (ss.x + ss.y);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + acc);
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

// X (n, d) fp64 -> per 128-row tile the image of the E kernel's A stage: [plane hi|lo][k-step][row][8 features] of (x - c)
// split into TF32 planes (features d..dp-1 and rows beyond n: zeros), 16-byte halves of a row swapped where the 32-byte
// swizzle of the UMMA descriptor swaps them (bit 2 of the row index)
__global__ void __launch_bounds__(256) gmm_tc_split_kernel(size_t n, int d, int dp, const double* X, const double* center, float* img) {
  const size_t tile = blockIdx.x;
  const int plane_f = dp * TC_M;  // floats per plane: dp / 8 k-steps x 128 rows x 8
  const int KS = dp / 8, KC = tc_kc(KS);
  float* out = img + tile * 2 * (size_t)plane_f;
  for (int idx = threadIdx.x; idx < plane_f; idx += 256) {
    const int pos = idx & 7, row = (idx >> 3) & (TC_M - 1), kk = idx >> 10;
    // chunk-major: [chunk][plane][k-step of the chunk][row][8]
    const int ch = kk / KC, kin = kk % KC, len = (ch + 1) * KC <= KS ? KC : KS - ch * KC;
    const size_t o_hi = (size_t)ch * KC * 2 * 1024 + (size_t)kin * 1024 + (size_t)(idx & 1023);
    const size_t o_lo = o_hi + (size_t)len * 1024;
    const int half = (pos >> 2) ^ ((row >> 2) & 1);  // logical half stored at this physical position
    const int j = 8 * kk + 4 * half + (pos & 3);
    const size_t r = tile * TC_M + row;
    float h = 0.f, l = 0.f;
    if (r < n && j < d) {
      const float x = (float)(X[r * d + j] - (center ? center[j] : 0.0));
      h = tf32_round(x);
      l = tf32_round(x - h);
    }
    out[o_hi] = h;
    out[o_lo] = l;
  }
}

// per-iteration operands: PT planes [(k, j)][i] = P_k[i][j] split in TF32, bias (k, j) = (mu_k - c) P_k[:, j], constants
__global__ void gmm_tc_params_kernel(int d, int dp, int K, const double* center, const double* w, const double* means,
                                     const double* pc, float* Phi, float* Plo, float* bias, float* cst) {
  const size_t total = (size_t)K * dp * dp;
  const size_t dd = (size_t)d * d;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % dp);
    const size_t kj = idx / dp;
    const int j = (int)(kj % dp), k = (int)(kj / dp);
    float h = 0.f, l = 0.f;
    if (i < d && j < d && i <= j) {
      const float v = (float)pc[k * dd + (size_t)i * d + j];
      h = tf32_round(v);
      l = tf32_round(v - h);
    }
    Phi[idx] = h;
    Plo[idx] = l;
    if (i == 0) {
      double acc = 0.0;
      if (j < d)
        for (int q = 0; q <= j; ++q) acc = fma(means[(size_t)k * d + q] - (center ? center[q] : 0.0), pc[k * dd + (size_t)q * d + j], acc);
      bias[kj] = (float)acc;
      if (j == 0) {
        double ld = 0.0;
        for (int q = 0; q < d; ++q) ld += log(pc[k * dd + (size_t)q * d + q]);
        cst[k] = (float)(ld + log(w[k]) - 0.5 * d * 1.8378770664093454835606594728112);
      }
    }
  }
}

// log-probabilities [K / kc][n][kc] fp32 -> responsibilities (n, K) fp64 and the log-sum-exp total of every GMM_TS-row
// block (sklearn/mixture/_base.py:552-582).  128 rows per CTA, one thread per row, fixed-order sums; the block's
// responsibilities are one contiguous run of the output and leave through shared memory with coalesced stores.
// respT / nkpart (optional, staged form only): the transposed single-precision responsibilities [tile][K][64] and their
// per-tile column sums in the layout the tensor-core M pass reads - written from the staged rows, so that the M pass needs
// neither the fp64 responsibilities nor its own transpose kernel (donk_gmm_tc_estep_m_f32)
__global__ void __launch_bounds__(128) gmm_tc_softmax_kernel(int n, int K, int kc, const float* logp, double* resp,
                                                             double* lse_part, int staged, float* respT, double* nkpart) {
  extern __shared__ double sresp[];  // [128][K | 1] when staged (odd row pitch: a thread per row, no bank conflicts)
  const int KP = K | 1;
  __shared__ double wsum[4];
  const size_t row0 = (size_t)blockIdx.x * 128, row = row0 + threadIdx.x;
  double lse = 0.0;
  if (row < (size_t)n) {
    // single precision throughout (the log-probabilities ARE fp32 and the pass is held to the fp32 tolerance): one expf
    // per entry, kept in the output slot and normalised in place; the fp64 form (three passes of double exp per row)
    // took 1.4 ms per 10 M rows x 16 clusters, four times its HBM time
    auto lp = [&](int k) { return logp[((size_t)(k / kc) * n + row) * kc + k % kc]; };
    float m = lp(0);
    for (int k = 1; k < K; ++k) { const float v = lp(k); m = v > m ? v : m; }
    double* out = staged ? sresp + (size_t)threadIdx.x * KP : resp + row * K;
    float s = 0.0f;
    for (int k = 0; k < K; ++k) { const float e = __expf(lp(k) - m); s += e; out[k] = (double)e; }
    const double inv = 1.0 / (double)s;
    for (int k = 0; k < K; ++k) out[k] *= inv;
    lse = (double)m + log((double)s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lse += __shfl_xor_sync(0xffffffffu, lse, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = lse;
  __syncthreads();
  static_assert(GMM_TS == 64, "lse_part is kept per 64-row block");
  if (threadIdx.x < 2) {
    const size_t blk = (size_t)blockIdx.x * 2 + threadIdx.x;
    if (blk * GMM_TS < (size_t)n) lse_part[blk] = wsum[2 * threadIdx.x] + wsum[2 * threadIdx.x + 1];
  }
  if (staged) {
    const size_t left = (size_t)n - row0;
    const size_t cnt = (left < 128 ? left : 128) * K;
    if (resp)
      for (size_t i = threadIdx.x; i < cnt; i += 128) resp[row0 * K + i] = sresp[(i / K) * KP + i % K];
    if (respT) {
      const int rows = (int)(left < 128 ? left : 128);
      for (int idx = threadIdx.x; idx < 2 * 64 * K; idx += 128) {
        const int t = idx / (64 * K), rem = idx - t * 64 * K, k = rem / 64, sx = rem - k * 64, rl = t * 64 + sx;
        if ((size_t)t * 64 + row0 < (size_t)((n + 63) / 64) * 64)  // the tile exists
          respT[((size_t)blockIdx.x * 2 + t) * K * 64 + rem] = rl < rows ? (float)sresp[rl * KP + k] : 0.0f;
      }
      if ((int)threadIdx.x < 2 * K) {
        const int t = threadIdx.x / K, k = threadIdx.x - t * K;
        if ((size_t)t * 64 + row0 < (size_t)((n + 63) / 64) * 64) {
          double acc = 0.0;
          for (int sx = 0; sx < 64; ++sx) { const int rl = t * 64 + sx; if (rl < rows) acc += sresp[rl * KP + k]; }
          nkpart[((size_t)blockIdx.x * 2 + t) * K + k] = acc;
        }
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor (rows, dp) with boxes of (8 floats, box_rows) and the 32-byte swizzle the UMMA descriptors expect.
// (Measured alternative: planes stored k-step-major [dp/8][rows][8] and ONE 3-D box per plane and tile - the tile is
// then dp/8 contiguous 4 KB runs - was slower, 15.4 vs 13.7 ms per pass at 10 M x 72, and makes the split copy a scatter.)
int make_map(CUtensorMap* tm, const float* base, size_t rows, int dp, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled is not available"); return DONK_CUDA_ERROR; }
  const cuuint64_t dims[2] = {(cuuint64_t)dp, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)dp * sizeof(float)};
  const cuuint32_t box[2] = {8u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed (%d)", (int)r); return DONK_CUDA_ERROR; }
  return DONK_OK;
}

struct TcPlan {
  int dp, kc, N, stages;
  size_t smem;
  bool ok;
};

TcPlan tc_plan(int d, int K) {
  TcPlan pl{};
  pl.dp = round_up(d, 8);
  const size_t cap = smem_optin_cap();
  const size_t tail = 1024 + 256 + 16 + (size_t)2 * pl.dp * sizeof(float);  // barriers, alignment slack, constants
  for (int kc = 2; kc >= 1 && !pl.ok; --kc) {
    if (K % kc) continue;
    const int N = kc * pl.dp;
    if (N % 16 || N > 256) continue;
    const size_t b = (size_t)2 * (pl.dp / 8) * N * 32, a = (size_t)2 * tc_kc(pl.dp / 8) * TC_M * 32;  // one stage = one chunk
    if (b + 2 * a + tail + (size_t)kc * pl.dp * 4 > cap) continue;
    int st = (int)((cap - b - tail - (size_t)kc * pl.dp * 4) / a);
    if (st > TC_MAX_STAGES) st = TC_MAX_STAGES;
    pl.kc = kc; pl.N = N; pl.stages = st;
    pl.smem = b + (size_t)st * a + tail + (size_t)kc * pl.dp * 4;
    pl.ok = 2 * N <= TC_TMEM_COLS;
  }
  return pl;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

template <int KS>
int tc_launch_ks(const CUtensorMap& c, const CUtensorMap& d, const TcParams& p, size_t smem, cudaStream_t st) {
  int rc = allow_smem(gmm_tc_estep_kernel<KS>, smem);
  if (rc != DONK_OK) return rc;
  gmm_tc_estep_kernel<KS><<<p.ncg * p.walkers, TC_THREADS, smem, st>>>(c, d, p);
  LAUNCHED("gmm_tc_estep_kernel");
  return DONK_OK;
}

int tc_launch(int KS, const CUtensorMap& c, const CUtensorMap& d, const TcParams& p, size_t smem, cudaStream_t st) {
  switch (KS) {
#define DONK_TC_CASE(V) case V: return tc_launch_ks<V>(c, d, p, smem, st);
    DONK_TC_CASE(1) DONK_TC_CASE(2) DONK_TC_CASE(3) DONK_TC_CASE(4) DONK_TC_CASE(5) DONK_TC_CASE(6) DONK_TC_CASE(7)
    DONK_TC_CASE(8) DONK_TC_CASE(9) DONK_TC_CASE(10)
#undef DONK_TC_CASE
  }
  return DONK_SMEM_LIMIT;
}


__device__ __forceinline__ void umma_tf32_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
// 8 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
               "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}

// =====================================================================================================================
// M pass on the tensor cores: the scatter statistics  S_k = sum_n r_nk (x_n - c)(x_n - c)^T  and the first moments
// s_k = sum_n r_nk (x_n - c)  as ONE GEMM  D[(k,i)][j] = sum_n A[(k,i)][n] B[j][n]  with the sample index as the K
// dimension:  B = the transposed pool (features x samples, row d is all ones so that column d of D is s_k), streamed by TMA
// from a transposed split copy made once per fit;  A[(k,i)][n] = r_nk x_ni does not exist in memory - four producer warps
// build it tile by tile from the B tile and the responsibilities, split into TF32 planes, straight into TENSOR MEMORY
// (tcgen05.st: thread = TMEM lane = row of the M-tile, 8 samples = 8 columns per k-step) and the products take their A
// operand from there (tcgen05.mma [d], [a_tmem], b_desc).  The first version staged A in shared memory: 21 ms per pass at
// 10 M x 72, K = 16, bound by the shared-memory port (A written once and read three times per tile on top of the B
// reads: ~930 KB per 64-row tile).  Three products per k-step as in the E pass (a_hi b_hi + a_lo b_hi + a_hi b_lo).  The K*d rows of D are cut into M-tiles of 128 TMEM lanes; a CTA group
// owns mt_per of them (accumulators N columns each) and its CTAs ("walkers") split the row tiles of the pool.  The fp32
// accumulators are drained into fp64 partial sums every FL tiles (the tensor core accumulates with truncation: the bias
// grows with the number of accumulation steps) by the same producer warps, double buffered in TMEM when both sets fit.
//     warp 0     one lane: TMA  B planes (8 samples x N rows boxes) + the tile's responsibilities (1-D bulk copy)
//     warp 1     one lane: tcgen05.mma M=128, N, K=8; commits release the A stage / B stage / publish the accumulators
//     warps 2-9  A producer (one row of the M-tile per thread, the two warps of a TMEM lane quarter split the k-steps),
//                accumulator drain
// =====================================================================================================================
constexpr int TM_TS = 64;            // samples per tile (32: 15.4 instead of 11.4 ms per pass - per-tile hand-shakes)
constexpr int TM_KST = TM_TS / 8;    // k-steps (UMMA instructions per plane pair) per tile
constexpr int TM_THREADS = 320;       // TMA warp, MMA warp, eight producer / drain warps (two per TMEM lane quarter)
constexpr int TM_PW = 8;
constexpr int TM_MAXST = 4;
constexpr int TM_ACOLS = 2 * TM_KST * 8;  // TMEM columns of one A stage: planes hi | lo, 8 samples per k-step

struct TcMParams {
  int n, d, K, N, mt_per, ncg, walkers, ntiles, SA, SB, FL, nacc;
  const unsigned char* bimg;  // [ntiles][plane hi|lo][k-step][N rows][32 B, swizzled]: the B stage images
  const float* respT;  // [ntiles][K][TM_TS]
  double* part;        // [grid][mt_per][128][N]
};

template <int MT>  // M-tiles per CTA (compile-time: the per-tile row offsets of a thread live in registers)
__global__ void __launch_bounds__(TM_THREADS, 1)
gmm_tc_mstep_kernel(const TcMParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((256u - (smem_u32(smem_raw) & 255u)) & 255u);
  const int N = p.N, K = p.K;
  const uint32_t b_plane = (uint32_t)TM_KST * N * 32, b_stage = 2 * b_plane;
  const uint32_t r_stage = (uint32_t)K * TM_TS * 4;
  unsigned char* sB = smem;                                  // [SB][plane][k-step][N rows][32 B]
  unsigned char* sR = sB + (size_t)p.SB * b_stage;           // [SB][K][TM_TS] floats
  unsigned long long* full_b = reinterpret_cast<unsigned long long*>(sR + (size_t)p.SB * r_stage);
  unsigned long long* empty_b = full_b + TM_MAXST;
  unsigned long long* full_a = empty_b + TM_MAXST;
  unsigned long long* empty_a = full_a + TM_MAXST;
  unsigned long long* acc_full = empty_a + TM_MAXST;   // [2]
  unsigned long long* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = blockIdx.x % p.ncg, wk = blockIdx.x / p.ncg;
  const int my_tiles = wk < p.ntiles ? (p.ntiles - wk + p.walkers - 1) / p.walkers : 0;
  // first drain period of this CTA: the CTAs walk the same number of tiles, so drains at the same tile count would hit
  // L2 from all SMs at once (measured 11.7 us per drain); spread over the period they overlap the other CTAs' products
  const int cnt0 = (int)(((long long)blockIdx.x * p.FL) / gridDim.x);

  if (threadIdx.x == 0) {
    for (int s = 0; s < TM_MAXST; ++s) {
      mbar_init(full_b + s, 1); mbar_init(empty_b + s, TM_PW + 1);
      mbar_init(full_a + s, TM_PW); mbar_init(empty_a + s, 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, TM_PW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int acc_cols = MT * N;
  // TMEM columns: [nacc accumulator sets of mt_per * N] [SA stages of the A operand: plane hi | lo, 8 columns per k-step]
  const uint32_t a_cols0 = (uint32_t)(p.nacc * acc_cols);

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int tile = wk + it * p.walkers;
        mbar_wait(empty_b + s, ph ^ 1);
        mbar_arrive_expect_tx(full_b + s, b_stage + r_stage);
        // the tile's B operand is stored in global memory as the image of the shared-memory stage (both planes, k-step
        // major, 32-byte swizzle applied by the split kernel): one contiguous bulk copy instead of 16 boxes of 80 rows
        // that lie 40 MB apart in a plain transposed copy (measured: the TMA boxes bound the kernel at 18.7 ms per pass)
        bulk_copy_g2s(sB + (size_t)s * b_stage, p.bimg + (size_t)tile * b_stage, b_stage, full_b + s);
        bulk_copy_g2s(sR + (size_t)s * r_stage, p.respT + (size_t)tile * K * TM_TS, r_stage, full_b + s);
        if (++s == p.SB) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // The whole warp walks the loop converged and ONE elected lane issues: every address below is then warp-uniform and
    // lives in uniform registers.  (With the loop under `if (lane == 0)` each tcgen05.mma cost ~19 instructions - R2UR
    // moves inside a vote loop - ~60 cycles per 40-cycle product: the issuing thread bounded the kernel.)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    const uint32_t bstep = ((uint32_t)N * 32u) >> 4;  // descriptor address units (16 bytes) per k-step
    int sb = 0, sa = 0, acc = 0, cnt = cnt0;
    uint32_t phb = 0, pha = 0, aph = 0;
    bool fresh = true;
    for (int it = 0; it < my_tiles; ++it) {
      if (fresh) mbar_wait(acc_empty + acc, aph ^ 1);  // the previous drain of this accumulator set is done
      mbar_wait(full_b + sb, phb);
      const uint64_t bd_hi = umma_desc_sw32(smem_u32(sB + (size_t)sb * b_stage));
      const uint64_t bd_lo = bd_hi + (uint64_t)(b_plane >> 4);
      for (int m = 0; m < MT; ++m) {
        mbar_wait(full_a + sa, pha);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_cols + m * N);
        const uint32_t a0 = tmem_base + a_cols0 + (uint32_t)sa * TM_ACOLS;
        if (elect_one()) {
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {  // a_hi b_hi + a_lo b_hi + a_hi b_lo
            const uint32_t ap = a0 + (pass == 1 ? TM_ACOLS / 2 : 0);
            const uint64_t bd = pass == 2 ? bd_lo : bd_hi;
#pragma unroll
            for (int kk = 0; kk < TM_KST; ++kk)
              umma_tf32_ta(d_tmem, ap + 8u * kk, bd + (uint64_t)(kk * bstep), idesc, (pass | kk) ? 1u : (fresh ? 0u : 1u));
          }
          tc_commit(empty_a + sa);
        }
        __syncwarp();
        if (++sa == p.SA) { sa = 0; pha ^= 1; }
      }
      const bool drain = ++cnt == p.FL || it + 1 == my_tiles;
      if (elect_one()) {
        tc_commit(empty_b + sb);
        if (drain) tc_commit(acc_full + acc);
      }
      __syncwarp();
      if (++sb == p.SB) { sb = 0; phb ^= 1; }
      fresh = drain;
      if (drain) {
        cnt = 0;
        if (++acc == p.nacc) { acc = 0; aph ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int t = q * 32 + lane;     // row of the M-tile (= TMEM lane) this thread builds and drains
    const int kh = (warp - 2) >> 2;  // the two warps of a lane quarter split the k-steps (and the drained columns)
    const int KD = K * p.d;
    int sb = 0, sa = 0, acc = 0, cnt = cnt0;
    uint32_t phb = 0, pha = 0, aph = 0;
    bool first = true;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    // row (cluster c, feature i) of every M-tile of this CTA, as byte offsets into the B stage (the feature row, with the
    // 32-byte swizzle of that row folded in) and the responsibility stage; -1: padding row beyond K * d
    int xoff[MT], roff[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const int r = (cg * MT + m) * TC_M + t;
      const int c = r / p.d, i = r % p.d;
      xoff[m] = r < KD ? i * 32 + (((i >> 2) & 1) ? 16 : 0) : -1;
      roff[m] = c * TM_TS * 4;
    }
    const float2 neg1 = make_float2(-1.f, -1.f);
    bool valid_all = true;  // CTA-uniform: no padding rows in any M-tile of this CTA
#pragma unroll
    for (int m = 0; m < MT; ++m) valid_all = valid_all && (cg * MT + m + 1) * TC_M <= KD;
    for (int it = 0; it < my_tiles; ++it) {
      mbar_wait(full_b + sb, phb);
      const unsigned char* bst = sB + (size_t)sb * b_stage;
      const unsigned char* rst = sR + (size_t)sb * r_stage;
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        mbar_wait(empty_a + sa, pha ^ 1);
        const uint32_t a_t = lane_base + a_cols0 + (uint32_t)sa * TM_ACOLS;
        // tcgen05.st is warp-collective (.sync.aligned): every lane issues it; padding rows (beyond K * d) read row 0 and
        // store zeros
        const bool valid = xoff[m] >= 0;
        const unsigned char* bh = bst + (valid ? xoff[m] : 0) + (size_t)kh * (TM_KST / 2) * N * 32;
        const unsigned char* rr = rst + (valid ? roff[m] : 0) + kh * (TM_KST / 2) * 32;
        const float2 scale = valid ? make_float2(1.f, 1.f) : make_float2(0.f, 0.f);
#pragma unroll 2
        for (int kk = kh * (TM_KST / 2); kk < (kh + 1) * (TM_KST / 2); ++kk) {
          // 8 samples of this feature (both planes; logical halves: the swizzle is in xoff) and of this cluster
          const float4 h0 = *reinterpret_cast<const float4*>(bh), h1 = *reinterpret_cast<const float4*>(reinterpret_cast<uintptr_t>(bh) ^ 16u);
          const float4 l0 = *reinterpret_cast<const float4*>(bh + b_plane), l1 = *reinterpret_cast<const float4*>(reinterpret_cast<uintptr_t>(bh + b_plane) ^ 16u);
          const float4 r0 = *reinterpret_cast<const float4*>(rr), r1 = *reinterpret_cast<const float4*>(rr + 16);
          // a = r (x_hi + x_lo) on packed fp32 pairs; a_hi = a with the 13 low mantissa bits cleared (what the tensor
          // core keeps of a TF32 operand), a_lo = a - a_hi exactly (its own low bits are dropped by the tensor core)
          const float2 xp[4] = {__fadd2_rn(make_float2(h0.x, h0.y), make_float2(l0.x, l0.y)), __fadd2_rn(make_float2(h0.z, h0.w), make_float2(l0.z, l0.w)),
                                __fadd2_rn(make_float2(h1.x, h1.y), make_float2(l1.x, l1.y)), __fadd2_rn(make_float2(h1.z, h1.w), make_float2(l1.z, l1.w))};
          float2 rp[4] = {make_float2(r0.x, r0.y), make_float2(r0.z, r0.w), make_float2(r1.x, r1.y), make_float2(r1.z, r1.w)};
          float hi[8], lo[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (!valid_all) rp[e] = __fmul2_rn(rp[e], scale);  // warp-uniform branch
            const float2 a = __fmul2_rn(rp[e], xp[e]);
            const float2 ah = make_float2(__uint_as_float(__float_as_uint(a.x) & 0xFFFFE000u), __uint_as_float(__float_as_uint(a.y) & 0xFFFFE000u));
            const float2 al = __ffma2_rn(ah, neg1, a);
            hi[2 * e] = ah.x; hi[2 * e + 1] = ah.y;
            lo[2 * e] = al.x; lo[2 * e + 1] = al.y;
          }
          tmem_st8(a_t + 8u * kk, hi);
          tmem_st8(a_t + TM_ACOLS / 2 + 8u * kk, lo);
          bh += (size_t)N * 32;
          rr += 32;
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();  // the stores are ordered before the arrival the MMA thread waits for
        __syncwarp();
        if (lane == 0) mbar_arrive(full_a + sa);
        if (++sa == p.SA) { sa = 0; pha ^= 1; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_b + sb);
      if (++sb == p.SB) { sb = 0; phb ^= 1; }
      if (++cnt == p.FL || it + 1 == my_tiles) {  // drain the accumulator set into the fp64 partial sums of this CTA
        cnt = 0;
        mbar_wait(acc_full + acc, aph);
        tc_fence_after();
        for (int m = 0; m < MT; ++m) {
          // partial sums [cta][m][column][lane]: the 32 lanes of a warp are one contiguous 256-byte run per column
          double* out = p.part + ((size_t)blockIdx.x * MT + m) * N * TC_M + t;
          const uint32_t taddr = lane_base + (uint32_t)(acc * acc_cols + m * N);
          for (int c0 = 16 * kh; c0 < N; c0 += 32) {
            uint32_t v[2][8];
            tmem_ld8_nowait(taddr + (uint32_t)c0, v[0]);
            tmem_ld8_nowait(taddr + (uint32_t)c0 + 8, v[1]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const double x = (double)__uint_as_float(v[h][j]);
                double* o = out + (size_t)(c0 + 8 * h + j) * TC_M;
                *o = first ? x : *o + x;
              }
          }
        }
        first = false;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty + acc);
        if (++acc == p.nacc) { acc = 0; aph ^= 1; }
      }
    }
    if (first) {  // a CTA without tiles still owns partial sums: zero them
      for (int m = 0; m < MT; ++m) {
        double* out = p.part + ((size_t)blockIdx.x * MT + m) * N * TC_M + t;
        for (int j = kh; j < N; j += 2) out[(size_t)j * TC_M] = 0.0;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

// X (n, d) fp64 -> per 64-row tile the image of the M kernel's B stage: [plane hi|lo][k-step][feature row j < NR][8 samples]
// of (x - c) split into TF32 planes, row d all ones (rows beyond n: zeros), 16-byte halves of a row swapped where the
// 32-byte swizzle of the UMMA descriptor swaps them (bit 2 of the row index)
__global__ void __launch_bounds__(256) gmm_tc_tblock_kernel(int n, int d, int NR, const double* X, const double* center, float* img) {
  const size_t tile = blockIdx.x;
  const int plane_f = TM_KST * NR * 8;  // floats per plane
  float* out = img + tile * 2 * (size_t)plane_f;
  for (int idx = threadIdx.x; idx < plane_f; idx += 256) {
    const int pos = idx & 7, j = (idx >> 3) % NR, kk = (idx >> 3) / NR;
    const int half = (pos >> 2) ^ ((j >> 2) & 1);  // logical half stored at this physical position
    const size_t r = tile * TM_TS + kk * 8 + half * 4 + (pos & 3);
    float h = 0.f, l = 0.f;
    if (r < (size_t)n) {
      if (j < d) {
        const float x = (float)(X[r * d + j] - (center ? center[j] : 0.0));
        h = tf32_round(x);
        l = tf32_round(x - h);
      } else if (j == d) {
        h = 1.f;
      }
    }
    out[idx] = h;
    out[plane_f + idx] = l;
  }
}

// responsibilities (n, K) fp64 -> fp32 tiles [tile][K][64] (zero beyond n) and the per-tile column sums (fp64, fixed order)
__global__ void __launch_bounds__(256) gmm_tc_respT_kernel(int n, int K, const double* resp, float* respT, double* nkpart) {
  extern __shared__ double srt[];  // [64][K | 1]: odd row pitch, the transposed reads below hit distinct banks
  const int KP = K | 1;
  const size_t row0 = (size_t)blockIdx.x * TM_TS;
  const size_t left = (size_t)n - row0;
  const int rows = (int)(left < (size_t)TM_TS ? left : (size_t)TM_TS);
  for (int idx = threadIdx.x; idx < TM_TS * K; idx += 256) srt[(idx / K) * KP + idx % K] = idx < rows * K ? resp[row0 * K + idx] : 0.0;
  __syncthreads();
  for (int idx = threadIdx.x; idx < TM_TS * K; idx += 256) {
    const int k = idx / TM_TS, s = idx % TM_TS;
    respT[(size_t)blockIdx.x * K * TM_TS + idx] = (float)srt[s * KP + k];
  }
  if ((int)threadIdx.x < K) {
    double acc = 0.0;
    for (int s = 0; s < TM_TS; ++s) acc += srt[s * KP + threadIdx.x];
    nkpart[(size_t)blockIdx.x * K + threadIdx.x] = acc;
  }
}

// per-CTA partial sums -> packed statistics [nk | s | S | lse] (fixed order: deterministic), shift = the centre
__global__ void __launch_bounds__(128) gmm_tc_mreduce_kernel(TcMParams p, size_t total, const double* nkpart, const double* lse_part,
                                                             int nblocks, const double* center, double* stats, double* shift) {
  const int d = p.d, K = p.K, N = p.N;
  const size_t n_elem = (size_t)K * d * (1 + d);
  if ((size_t)blockIdx.x * 128 >= n_elem) {  // the K + 1 tail blocks: nk[k] and the log-sum-exp total (cooperative sums)
    const int which = (int)(blockIdx.x - (n_elem + 127) / 128);
    __shared__ double red[128];
    double acc = 0.0;
    if (which < K) for (int t = threadIdx.x; t < p.ntiles; t += 128) acc += nkpart[(size_t)t * K + which];
    else for (int b = threadIdx.x; b < nblocks; b += 128) acc += lse_part[b];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) stats[which < K ? (size_t)which : total - 1] = red[0];
    return;
  }
  const size_t idx = (size_t)blockIdx.x * 128 + threadIdx.x;
  if (idx >= n_elem) return;
  int k, i, j;
  size_t dst;
  if (idx < (size_t)K * d) { k = (int)(idx / d); i = (int)(idx % d); j = d; dst = (size_t)K + idx; shift[idx] = center ? center[i] : 0.0; }
  else {
    const size_t r = idx - (size_t)K * d;
    k = (int)(r / ((size_t)d * d)); i = (int)((r / d) % d); j = (int)(r % d);
    dst = (size_t)K * (1 + d) + r;
    if (i > j) { const int tswap = i; i = j; j = tswap; }  // both triangles from the upper one: exactly symmetric
  }
  const int row = k * d + i, mt = row / TC_M, cgi = mt / p.mt_per, m = mt % p.mt_per;
  double acc = 0.0;
  for (int w = 0; w < p.walkers; ++w)
    acc += p.part[((((size_t)w * p.ncg + cgi) * p.mt_per + m) * N + j) * TC_M + row % TC_M];
  stats[dst] = acc;
}

struct TcMPlan {
  int N, MT, mt_per, ncg, walkers, SA, SB, nacc;
  size_t smem;
  bool ok;
};

TcMPlan tcm_plan(int d, int K) {
  TcMPlan pl{};
  pl.N = round_up(d + 1, 16);
  if (d < 1 || K < 1 || pl.N > 256 || d > 80) return pl;  // d <= 80: the shapes the tensor-core E pass covers
  pl.MT = (K * d + TC_M - 1) / TC_M;
  const int cap_mt = (TC_TMEM_COLS - 2 * TM_ACOLS) / pl.N;  // accumulators next to two stages of the A operand
  double best = 1e30;
  for (int ncg = 1; ncg <= pl.MT && ncg <= 148; ++ncg) {
    const int per = (pl.MT + ncg - 1) / ncg;
    if (per > cap_mt || per > 8) continue;
    const double cost = (double)per / (148 / ncg);
    if (cost < best - 1e-12) { best = cost; pl.ncg = ncg; pl.mt_per = per; }
  }
  if (!pl.ncg) return pl;
  pl.walkers = 148 / pl.ncg;
  pl.nacc = 2 * pl.mt_per * pl.N + 2 * TM_ACOLS <= TC_TMEM_COLS ? 2 : 1;
  int sa = (TC_TMEM_COLS - pl.nacc * pl.mt_per * pl.N) / TM_ACOLS;
  pl.SA = sa > TM_MAXST ? TM_MAXST : sa;
  const size_t cap = smem_optin_cap();
  const size_t b_stage = (size_t)2 * TM_KST * pl.N * 32, r_stage = (size_t)K * TM_TS * 4;
  const size_t tail = 1024 + 256;
  if (2 * (b_stage + r_stage) + tail > cap) return pl;
  int sb = (int)((cap - tail) / (b_stage + r_stage));
  pl.SB = sb > TM_MAXST ? TM_MAXST : sb;
  pl.smem = pl.SB * (b_stage + r_stage) + tail;
  pl.ok = true;
  return pl;
}

size_t tcm_npad(int n) { return (size_t)round_up(n > 0 ? n : 1, TM_TS); }

}  // namespace

extern "C" {

/* 1 when the tensor-core E pass supports (d, K) on this device (operands must fit shared memory: d <= 80). */
int donk_gmm_tc_supported(int d, int K) { return d >= 1 && K >= 1 && tc_plan(d, K).ok ? 1 : 0; }

/* bytes of the split copy of the pool: two fp32 planes (n, round_up(d, 8)) */
size_t donk_gmm_tc_planes_bytes(int n, int d) { return align256((size_t)2 * round_up(n > 0 ? n : 1, TC_M) * round_up(d, 8) * sizeof(float)); }

/* bytes of the per-iteration scratch: P planes, bias, constants, log-probabilities */
size_t donk_gmm_tc_workspace_bytes(int n, int d, int K) {
  const size_t dp = round_up(d, 8);
  return 2 * align256((size_t)K * dp * dp * 4) + align256((size_t)K * dp * 4) + align256((size_t)K * 4) +
         align256((size_t)n * K * 4) + 256;
}

/* X (n, d) fp64 minus center (d, may be NULL) -> planes (donk_gmm_tc_planes_bytes), once per fit. */
int donk_gmm_tc_prepare_f32(int n, int d, const double* X, const double* center, void* planes, size_t planes_bytes,
                            void* stream) {
  if (n < 0 || d < 1) return DONK_BAD_SHAPE;
  if (planes_bytes < donk_gmm_tc_planes_bytes(n, d) || ((uintptr_t)planes & 255)) return DONK_BAD_SHAPE;
  if (n == 0) return DONK_OK;
  const int dp = round_up(d, 8);
  gmm_tc_split_kernel<<<(unsigned)((n + TC_M - 1) / TC_M), 256, 0, (cudaStream_t)stream>>>((size_t)n, d, dp, X, center, static_cast<float*>(planes));
  LAUNCHED("gmm_tc_split_kernel");
  return DONK_OK;
}

/* E pass at fp32 tolerance: planes from donk_gmm_tc_prepare_f32 (same center), current parameters (fp64) ->
 * responsibilities (n, K) fp64 and per-64-row log-sum-exp totals, both written where the fp64 E pass puts them in the EM
 * workspace (resp | lse_part), so that donk_gmm_em_step_mode_f64(mode = 3) continues with the M pass. */
// layout of the M pass's per-iteration scratch (donk_gmm_tc_mwork_bytes)
static void tcm_layout(void* work, int n, int K, float** respT, double** nkpart, double** part) {
  const int ntiles = (int)(tcm_npad(n) / TM_TS);
  char* wp = reinterpret_cast<char*>(((uintptr_t)work + 255) & ~(uintptr_t)255);
  *respT = reinterpret_cast<float*>(wp); wp += align256((size_t)ntiles * K * TM_TS * 4);
  *nkpart = reinterpret_cast<double*>(wp); wp += align256((size_t)ntiles * K * 8);
  *part = reinterpret_cast<double*>(wp);
}

static int tc_estep_impl(int n, int d, int K, const void* planes, const double* center, const double* weights,
                         const double* means, const double* prec_chol, double* resp, double* lse_part, void* work,
                         size_t work_bytes, void* mwork, void* stream) {
  if (n < 0 || d < 1 || K < 1) return DONK_BAD_SHAPE;
  const TcPlan pl = tc_plan(d, K);
  if (!pl.ok) { snprintf(g_err, sizeof(g_err), "gmm_tc: d=%d K=%d does not fit the tensor-core E pass", d, K); return DONK_SMEM_LIMIT; }
  if (work_bytes < donk_gmm_tc_workspace_bytes(n, d, K) || ((uintptr_t)planes & 255)) return DONK_BAD_SHAPE;
  if (n == 0) return DONK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int dp = pl.dp;
  char* wp = reinterpret_cast<char*>(((uintptr_t)work + 255) & ~(uintptr_t)255);
  float* Phi = reinterpret_cast<float*>(wp); wp += align256((size_t)K * dp * dp * 4);
  float* Plo = reinterpret_cast<float*>(wp); wp += align256((size_t)K * dp * dp * 4);
  float* bias = reinterpret_cast<float*>(wp); wp += align256((size_t)K * dp * 4);
  float* cst = reinterpret_cast<float*>(wp); wp += align256((size_t)K * 4);
  float* logp = reinterpret_cast<float*>(wp);
  {
    const size_t total = (size_t)K * dp * dp;
    gmm_tc_params_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d, dp, K, center, weights, means, prec_chol, Phi, Plo,
                                                                         bias, cst);
    LAUNCHED("gmm_tc_params_kernel");
  }
  CUtensorMap tPh, tPl;
  int rc = make_map(&tPh, Phi, (size_t)K * dp, dp, pl.N);
  if (rc == DONK_OK) rc = make_map(&tPl, Plo, (size_t)K * dp, dp, pl.N);
  if (rc != DONK_OK) return rc;
  TcParams p;
  p.n = n; p.d = d; p.dp = dp; p.K = K; p.kc = pl.kc; p.ncg = K / pl.kc; p.N = pl.N; p.stages = pl.stages;
  p.ntiles = (n + TC_M - 1) / TC_M;
  int walkers = 148 / p.ncg;
  if (walkers < 1) walkers = 1;
  if (walkers > p.ntiles) walkers = p.ntiles;
  p.walkers = walkers;
  p.bias = bias; p.cst = cst; p.logp = logp; p.ximg = static_cast<const unsigned char*>(planes);
  rc = tc_launch(pl.dp / 8, tPh, tPl, p, pl.smem, st);
  if (rc != DONK_OK) return rc;
  const size_t sm_bytes = (size_t)128 * (K | 1) * sizeof(double);
  const int staged = sm_bytes <= 40 * 1024;
  float* respT = nullptr;
  double *nkpart = nullptr, *part = nullptr;
  if (mwork) tcm_layout(mwork, n, K, &respT, &nkpart, &part);
  (void)part;
  gmm_tc_softmax_kernel<<<(unsigned)((n + 127) / 128), 128, staged ? sm_bytes : 0, st>>>(
      n, K, pl.kc, logp, resp, lse_part, staged, staged ? respT : nullptr, staged ? nkpart : nullptr);
  LAUNCHED("gmm_tc_softmax_kernel");
  if (mwork && !staged) {  // many clusters: the rows do not fit the staging area - separate transpose
    const int ntiles = (int)(tcm_npad(n) / TM_TS);
    gmm_tc_respT_kernel<<<ntiles, 256, (size_t)TM_TS * (K | 1) * sizeof(double), st>>>(n, K, resp, respT, nkpart);
    LAUNCHED("gmm_tc_respT_kernel");
  }
  return DONK_OK;
}

int donk_gmm_tc_estep_f32(int n, int d, int K, const void* planes, const double* center, const double* weights,
                          const double* means, const double* prec_chol, double* resp, double* lse_part, void* work,
                          size_t work_bytes, void* stream) {
  return tc_estep_impl(n, d, K, planes, center, weights, means, prec_chol, resp, lse_part, work, work_bytes, nullptr, stream);
}

int donk_gmm_tc_estep_m_f32(int n, int d, int K, const void* planes, const double* center, const double* weights,
                            const double* means, const double* prec_chol, double* resp, double* lse_part, void* work,
                            size_t work_bytes, void* mwork, size_t mwork_bytes, void* stream) {
  if (!mwork || mwork_bytes < donk_gmm_tc_mwork_bytes(n, d, K) || !resp) return DONK_BAD_SHAPE;
  return tc_estep_impl(n, d, K, planes, center, weights, means, prec_chol, resp, lse_part, work, work_bytes, mwork, stream);
}

/* ---- M pass at fp32 tolerance on the tensor cores (see gmm_tc_mstep_kernel) ---- */
int donk_gmm_tc_m_supported(int d, int K) { return tcm_plan(d, K).ok ? 1 : 0; }

/* bytes of the transposed split copy of the pool: two fp32 planes (round_up(d + 1, 16), round_up(n, 64)) */
size_t donk_gmm_tc_mplanes_bytes(int n, int d) { return align256((size_t)2 * round_up(d + 1, 16) * tcm_npad(n) * sizeof(float)); }

/* bytes of the per-iteration scratch: fp32 responsibility tiles, per-tile column sums, per-CTA partial sums */
size_t donk_gmm_tc_mwork_bytes(int n, int d, int K) {
  const TcMPlan pl = tcm_plan(d, K);
  const size_t ntiles = tcm_npad(n) / TM_TS;
  const size_t grid = pl.ok ? (size_t)pl.ncg * pl.walkers : 1, per = pl.ok ? (size_t)pl.mt_per : 1;
  return align256(ntiles * K * TM_TS * 4) + align256(ntiles * K * 8) + align256(grid * per * TC_M * round_up(d + 1, 16) * 8) + 256;
}

int donk_gmm_tc_mprepare_f32(int n, int d, const double* X, const double* center, void* mplanes, size_t mplanes_bytes,
                             void* stream) {
  if (n < 0 || d < 1) return DONK_BAD_SHAPE;
  if (mplanes_bytes < donk_gmm_tc_mplanes_bytes(n, d) || ((uintptr_t)mplanes & 255)) return DONK_BAD_SHAPE;
  if (n == 0) return DONK_OK;
  const int NR = round_up(d + 1, 16);
  gmm_tc_tblock_kernel<<<(unsigned)(tcm_npad(n) / TM_TS), 256, 0, (cudaStream_t)stream>>>(n, d, NR, X, center, static_cast<float*>(mplanes));
  LAUNCHED("gmm_tc_tblock_kernel");
  return DONK_OK;
}

/* M pass: mplanes from donk_gmm_tc_mprepare_f32 (same center), responsibilities (n, K) fp64 and per-64-row log-sum-exp
 * totals as the E pass left them in the EM workspace -> the packed statistics (donk_gmm_stats_size) of this rank's rows,
 * centred on `center` for every cluster (written to shift (K, d)); continue with the all-reduce and donk_gmm_m_finalize_f64. */
int donk_gmm_tc_mstep_f32(int n, int d, int K, const void* mplanes, const double* center, const double* resp,
                          const double* lse_part, double* stats, double* shift, void* work, size_t work_bytes, void* stream) {
  if (n < 0 || d < 1 || K < 1) return DONK_BAD_SHAPE;
  const TcMPlan pl = tcm_plan(d, K);
  if (!pl.ok) { snprintf(g_err, sizeof(g_err), "gmm_tc: d=%d K=%d does not fit the tensor-core M pass", d, K); return DONK_SMEM_LIMIT; }
  if (work_bytes < donk_gmm_tc_mwork_bytes(n, d, K) || ((uintptr_t)mplanes & 255)) return DONK_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = gmm_stats_len(d, K);
  if (n == 0) {
    CU(cudaMemsetAsync(stats, 0, total * sizeof(double), st));
    return DONK_OK;
  }
  const size_t n_pad = tcm_npad(n);
  const int ntiles = (int)(n_pad / TM_TS);
  float* respT;
  double *nkpart, *part;
  tcm_layout(work, n, K, &respT, &nkpart, &part);
  if (resp) {  // NULL: donk_gmm_tc_estep_m_f32 left the transposed responsibilities and their column sums in `work`
    gmm_tc_respT_kernel<<<ntiles, 256, (size_t)TM_TS * (K | 1) * sizeof(double), st>>>(n, K, resp, respT, nkpart);
    LAUNCHED("gmm_tc_respT_kernel");
  }
  int rc = DONK_OK;
  TcMParams p;
  p.n = n; p.d = d; p.K = K; p.N = pl.N; p.mt_per = pl.mt_per; p.ncg = pl.ncg; p.ntiles = ntiles;
  p.walkers = pl.walkers < ntiles ? pl.walkers : ntiles;
  p.SA = pl.SA; p.SB = pl.SB; p.nacc = pl.nacc;
  const int fl = env_int("DONK_GMM_TC_FLUSH");
  p.FL = fl > 0 ? fl : 4096 / TM_TS;  // tiles per drain (error of the fp32 accumulation ~ 2e-7 per tile of 32 rows)
  p.respT = respT; p.part = part; p.bimg = static_cast<const unsigned char*>(mplanes);
  switch (p.mt_per) {
#define DONK_TCM_CASE(V)                                                                          \
    case V:                                                                                       \
      rc = allow_smem(gmm_tc_mstep_kernel<V>, pl.smem);                                           \
      if (rc != DONK_OK) return rc;                                                               \
      gmm_tc_mstep_kernel<V><<<p.ncg * p.walkers, TM_THREADS, pl.smem, st>>>(p);        \
      break;
    DONK_TCM_CASE(1) DONK_TCM_CASE(2) DONK_TCM_CASE(3) DONK_TCM_CASE(4) DONK_TCM_CASE(5) DONK_TCM_CASE(6) DONK_TCM_CASE(7)
    DONK_TCM_CASE(8)
#undef DONK_TCM_CASE
    default: return DONK_SMEM_LIMIT;
  }
  LAUNCHED("gmm_tc_mstep_kernel");
  const size_t n_elem = (size_t)K * d * (1 + d);
  const int nblocks = (n + GMM_TS - 1) / GMM_TS;
  gmm_tc_mreduce_kernel<<<(unsigned)((n_elem + 127) / 128 + K + 1), 128, 0, st>>>(p, total, nkpart, lse_part, nblocks, center, stats, shift);
  LAUNCHED("gmm_tc_mreduce_kernel");
  return DONK_OK;
}

}  // extern "C"
