// donk_b200 — shared device helpers.
//
// Kernels are written as bulk-synchronous CTA phases over shared memory:
//     TEAM_FOR(i, n_items) { ... }   ctx.sync();
// No thread-private state is carried across a sync except CTA-uniform values.
// That discipline lets the very same kernel bodies compile two ways:
//   * nvcc, sm_100a: TEAM_FOR strides the CTA's threads, sync = __syncthreads();
//   * g++ with -DDONK_EMU (tests/emu only): one host thread plays the whole CTA,
//     TEAM_FOR is a sequential loop (optionally reversed, to expose intra-phase
//     dependencies) and sync is a no-op.
// The emulation build is test infrastructure for kernel logic on a box without a
// GPU; it is never loaded by the product path (donk_b200/_lib.py loads the CUDA
// library only and raises if it is missing).
#pragma once
#include <cstring>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DONK_HD __host__ __device__ __forceinline__
#define DONK_D __device__ __forceinline__
#else
#define DONK_HD inline
#define DONK_D inline
#endif

namespace donk {

#if defined(DONK_EMU)
extern int g_emu_reverse;  // tests flip this to run every phase back to front
struct Ctx {
  int tid, nt;
  inline void sync() const {}
};
static inline int emu_idx(int it, int n) { return g_emu_reverse ? n - 1 - it : it; }
#define TEAM_FOR(i, n)                                                               \
  for (int i##_it = 0, i##_n = (n), i = donk::emu_idx(0, i##_n); i##_it < i##_n;     \
       ++i##_it, i = donk::emu_idx(i##_it, i##_n))
#else
// The "team" that executes one generic kernel body is the whole CTA.
struct Ctx {
  int tid, nt;
  DONK_D void sync() const { __syncthreads(); }
};
#define TEAM_FOR(i, n) for (int i = ctx.tid; i < (n); i += ctx.nt)
#endif

// Warp-team kernels (ilqr_warp.cuh): a CTA of G warps, one problem per warp; CTA-wide steps only hand over
// the staged per-condition inputs.  Under emulation one host thread plays all warps of the CTA in turn.
#if defined(DONK_EMU)
struct WCtx {
  int tid, nthreads, team, lane, G;
  inline void cta_sync() const {}
  inline void team_sync() const {}
};
#define CTA_FOR(i, n)                                                               \
  for (int i##_it = 0, i##_n = (n), i = donk::emu_idx(0, i##_n); i##_it < i##_n;     \
       ++i##_it, i = donk::emu_idx(i##_it, i##_n))
#define LANE_FOR(i, n) CTA_FOR(i, n)
#define TEAMS_DO(tm) for (int tm = 0; tm < w.G; ++tm)
#define WARP_ROWS_FOR(r, n) for (int r = 0; r < (n); ++r)
#else
struct WCtx {
  int tid, nthreads, team, lane, G;
  DONK_D void cta_sync() const { __syncthreads(); }
  DONK_D void team_sync() const { __syncwarp(); }
};
#define CTA_FOR(i, n) for (int i = w.tid; i < (n); i += w.nthreads)
#define LANE_FOR(i, n) for (int i = w.lane; i < (n); i += 32)
#define TEAMS_DO(tm) for (int tm = w.team, tm##_once = 0; tm##_once == 0; tm##_once = 1)
// rows of a staged matrix dealt to the warps of the CTA (the lanes then walk the row: no div/mod per element)
#define WARP_ROWS_FOR(r, n) for (int r = w.team; r < (n); r += w.G)
#endif

DONK_HD int round_up(int x, int m) { return (x + m - 1) / m * m; }

// 1/sqrt(x): device intrinsic (MUFU.RSQ64H + Newton steps); plain expression under emulation
DONK_HD double rsqrt_r(double x) {
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}
DONK_HD float rsqrt_r(float x) {
#if defined(__CUDA_ARCH__)
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}

// Read-only global load (inputs never written by the running kernel).
template <typename R>
DONK_HD R ldg(const R* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// 4 consecutive values from shared memory; p must be 16-byte aligned for double
// (two LDS.128) — all tile origins and leading dimensions are multiples of 4.
DONK_HD void ld4(const double* p, double o[4]) {
#if defined(__CUDA_ARCH__)
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *reinterpret_cast<const double2*>(p + 2);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
#else
  o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = p[3];
#endif
}
DONK_HD void ld4(const float* p, float o[4]) {
#if defined(__CUDA_ARCH__)
  const float4 a = *reinterpret_cast<const float4*>(p);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
#else
  o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = p[3];
#endif
}

// 4x4 register tile of a product whose operands are both stored k-major:
//   acc[r][c] = sum_{k<kdim} A[k*lda + i0 + r] * Bm[k*ldb + j0 + c]
// 16 FMAs per 4 LDS.128 (double) per k.  i0, j0, lda, ldb multiples of 4.
template <typename R>
DONK_HD void tile4x4(const R* __restrict__ A, int lda, int i0, const R* __restrict__ Bm, int ldb, int j0,
                     int kdim, R acc[4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = R(0);
  const R* a = A + i0;
  const R* b = Bm + j0;
#pragma unroll 2
  for (int k = 0; k < kdim; ++k) {
    R av[4], bv[4];
    ld4(a, av);
    ld4(b, bv);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = fma(av[r], bv[c], acc[r][c]);
    a += lda;
    b += ldb;
  }
}

// TM x TN register tile of a product whose operands are both stored k-major (generalises tile4x4):
//   acc[r][c] = sum_{k<kdim} A[k*lda + r] * Bm[k*ldb + c]      (A, Bm already offset to the tile origin)
// VA / VB: the operand origin and leading dimension are 16-byte aligned and TM / TN even -> LDS.128.
template <typename R, int TM, int TN, bool VA, bool VB>
DONK_HD void tile_km(const R* __restrict__ A, int lda, const R* __restrict__ Bm, int ldb, int kdim, R (&acc)[TM][TN]) {
#pragma unroll
  for (int r = 0; r < TM; ++r)
#pragma unroll
    for (int c = 0; c < TN; ++c) acc[r][c] = R(0);
#pragma unroll 4
  for (int k = 0; k < kdim; ++k) {
    R av[TM], bv[TN];
    if (VA && TM == 4) ld4(A, av);
    else {
#pragma unroll
      for (int r = 0; r < TM; ++r) av[r] = A[r];
    }
    if (VB && TN == 4) ld4(Bm, bv);
    else {
#pragma unroll
      for (int c = 0; c < TN; ++c) bv[c] = Bm[c];
    }
#pragma unroll
    for (int r = 0; r < TM; ++r)
#pragma unroll
      for (int c = 0; c < TN; ++c) acc[r][c] = fma(av[r], bv[c], acc[r][c]);
    A += lda;
    Bm += ldb;
  }
}

// Warp-collective block product on the FP64 tensor path (mma.sync.m8n8k4.f64, "DMMA"):
//   C[i][j] = sum_{k < 4*K4} A[k*lda + i] * Bm[k*ldb + j],   i < 8*RA, j < 8*RB   (both operands k-major)
// A and Bm point at the block origin.  Every lane ends up with two adjacent columns of one row per 8x8
// fragment and hands them to epi(i, j, v0, v1) (j even; values of columns j and j+1).  One LDS.64 per
// fragment and k4-step feeds 8 DFMA-equivalents per lane, which is what makes the step kernel FP64-bound
// instead of shared-memory-bound (profiles/r01/README.md).  Leading dimensions with lda % 16 in {4, 12}
// make the fragment loads bank-conflict free.  Rows/columns beyond the valid range may be over-read
// (their outputs are discarded by the epilogue); the k range must be zero-padded in one operand and
// finite in the other.  float has no such instruction: each lane computes its two outputs with FFMA.
// Values a lane fetches for the epilogue of a product (addends read from global memory): NP raw values per fragment.
// frag_pre issues the loads (raw loads only, no arithmetic on the loaded values: they stay in flight while other
// work runs — possibly a whole earlier product); warp_mma_ep runs the product and hands them to the epilogue.
template <typename R, int RA, int RB, int NP>
struct FragVals {
#if defined(__CUDA_ARCH__)
  R v[RA][RB][NP];
#else
  R v[8 * RA][4 * RB][NP];
#endif
};

template <typename R, int RA, int RB, int NP, class Pre>
DONK_D void frag_pre(const WCtx& w, FragVals<R, RA, RB, NP>& fv, Pre&& pre) {
#if defined(__CUDA_ARCH__)
  const int g = w.lane >> 2, tq = w.lane & 3;
#pragma unroll
  for (int ra = 0; ra < RA; ++ra)
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
#pragma unroll
      for (int q = 0; q < NP; ++q) fv.v[ra][rb][q] = R(0);
      pre(8 * ra + g, 8 * rb + 2 * tq, fv.v[ra][rb]);
    }
#else
  (void)w;
  for (int i = 0; i < 8 * RA; ++i)
    for (int j = 0; j < 8 * RB; j += 2) {
      for (int q = 0; q < NP; ++q) fv.v[i][j / 2][q] = R(0);
      pre(i, j, fv.v[i][j / 2]);
    }
#endif
}

// AT / BT: the operand is stored the other way round (A[i*lda + k] / Bm[j*ldb + k]).  Both orientations are
// conflict-free for the fragment loads at leading dimensions with ld % 16 in {4, 12}, so a matrix that one product
// produces row by row can feed the next product in either role without a transposed copy.
template <typename R, int RA, int RB, int NP, bool AT = false, bool BT = false, class Epi>
DONK_D void warp_mma_ep(const WCtx& w, const R* __restrict__ A, int lda, const R* __restrict__ Bm, int ldb, int K4,
                        const FragVals<R, RA, RB, NP>& fv, Epi&& epi) {
#if defined(__CUDA_ARCH__)
  const int g = w.lane >> 2, tq = w.lane & 3;
  R c[RA][RB][2];
#pragma unroll
  for (int ra = 0; ra < RA; ++ra)
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) { c[ra][rb][0] = R(0); c[ra][rb][1] = R(0); }
  if (sizeof(R) == 8) {
    const R* ap = AT ? A + g * lda + tq : A + tq * lda + g;
    const R* bp = BT ? Bm + g * ldb + tq : Bm + tq * ldb + g;
    const int sa = AT ? 8 * lda : 8, sb = BT ? 8 * ldb : 8;  // step to the next 8-row / 8-column fragment
#pragma unroll 8
    for (int k4 = 0; k4 < K4; ++k4) {
      R af[RA], bf[RB];
#pragma unroll
      for (int ra = 0; ra < RA; ++ra) af[ra] = ap[sa * ra];
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) bf[rb] = bp[sb * rb];
#pragma unroll
      for (int ra = 0; ra < RA; ++ra)
#pragma unroll
        for (int rb = 0; rb < RB; ++rb) {
          double d0 = (double)c[ra][rb][0], d1 = (double)c[ra][rb][1];
          asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(d0), "+d"(d1)
                       : "d"((double)af[ra]), "d"((double)bf[rb]));
          c[ra][rb][0] = (R)d0;
          c[ra][rb][1] = (R)d1;
        }
      ap += AT ? 4 : 4 * lda;
      bp += BT ? 4 : 4 * ldb;
    }
  } else {
    for (int k = 0; k < 4 * K4; ++k) {
#pragma unroll
      for (int ra = 0; ra < RA; ++ra) {
        const R av = AT ? A[(8 * ra + g) * lda + k] : A[k * lda + 8 * ra + g];
#pragma unroll
        for (int rb = 0; rb < RB; ++rb) {
          const int j0 = 8 * rb + 2 * tq;
          c[ra][rb][0] = fma(av, BT ? Bm[j0 * ldb + k] : Bm[k * ldb + j0], c[ra][rb][0]);
          c[ra][rb][1] = fma(av, BT ? Bm[(j0 + 1) * ldb + k] : Bm[k * ldb + j0 + 1], c[ra][rb][1]);
        }
      }
    }
  }
#pragma unroll
  for (int ra = 0; ra < RA; ++ra)
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
      epi(8 * ra + g, 8 * rb + 2 * tq, c[ra][rb][0], c[ra][rb][1], fv.v[ra][rb]);
#else
  (void)w;
  for (int i = 0; i < 8 * RA; ++i)
    for (int j = 0; j < 8 * RB; j += 2) {
      R v0 = R(0), v1 = R(0);
      for (int k = 0; k < 4 * K4; ++k) {
        const R av = AT ? A[i * lda + k] : A[k * lda + i];
        v0 = fma(av, BT ? Bm[j * ldb + k] : Bm[k * ldb + j], v0);
        v1 = fma(av, BT ? Bm[(j + 1) * ldb + k] : Bm[k * ldb + j + 1], v1);
      }
      epi(i, j, v0, v1, fv.v[i][j / 2]);
    }
#endif
}

template <typename R, int RA, int RB, int NP, bool AT = false, bool BT = false, class Pre, class Epi>
DONK_D void warp_mma_pre(const WCtx& w, const R* __restrict__ A, int lda, const R* __restrict__ Bm, int ldb, int K4,
                         Pre&& pre, Epi&& epi) {
  FragVals<R, RA, RB, NP> fv;
  frag_pre<R, RA, RB, NP>(w, fv, pre);
  warp_mma_ep<R, RA, RB, NP, AT, BT>(w, A, lda, Bm, ldb, K4, fv, epi);
}

// Block-upper-triangular variant: only the fragments with rb >= ra are computed (symmetric results: Q, V, Sigma).
// One call instead of one call per block row: the A and B fragments of a k-step are loaded once for all of its
// fragments (RA + RB loads instead of sum_r (1 + RB - r)), all accumulators are independent, one prologue / epilogue.
template <typename R, int RA, int RB, int NP, bool AT = false, bool BT = false, class Pre, class Epi>
DONK_D void warp_mma_tri(const WCtx& w, const R* __restrict__ A, int lda, const R* __restrict__ Bm, int ldb, int K4,
                         Pre&& pre, Epi&& epi) {
#if defined(__CUDA_ARCH__)
  const int g = w.lane >> 2, tq = w.lane & 3;
  R c[RA][RB][2], pv[RA][RB][NP];
#pragma unroll
  for (int ra = 0; ra < RA; ++ra)
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
      if (rb >= ra) {
        c[ra][rb][0] = R(0); c[ra][rb][1] = R(0);
#pragma unroll
        for (int q = 0; q < NP; ++q) pv[ra][rb][q] = R(0);
        pre(8 * ra + g, 8 * rb + 2 * tq, pv[ra][rb]);
      }
  if (sizeof(R) == 8) {
    const R* ap = AT ? A + g * lda + tq : A + tq * lda + g;
    const R* bp = BT ? Bm + g * ldb + tq : Bm + tq * ldb + g;
    const int sa = AT ? 8 * lda : 8, sb = BT ? 8 * ldb : 8;
#pragma unroll 8
    for (int k4 = 0; k4 < K4; ++k4) {
      R af[RA], bf[RB];
#pragma unroll
      for (int ra = 0; ra < RA; ++ra) af[ra] = ap[sa * ra];
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) bf[rb] = bp[sb * rb];
#pragma unroll
      for (int ra = 0; ra < RA; ++ra)
#pragma unroll
        for (int rb = 0; rb < RB; ++rb)
          if (rb >= ra) {
            double d0 = (double)c[ra][rb][0], d1 = (double)c[ra][rb][1];
            asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(d0), "+d"(d1)
                : "d"((double)af[ra]), "d"((double)bf[rb]));
            c[ra][rb][0] = (R)d0;
            c[ra][rb][1] = (R)d1;
          }
      ap += AT ? 4 : 4 * lda;
      bp += BT ? 4 : 4 * ldb;
    }
  } else {
    for (int k = 0; k < 4 * K4; ++k) {
#pragma unroll
      for (int ra = 0; ra < RA; ++ra) {
        const R av = AT ? A[(8 * ra + g) * lda + k] : A[k * lda + 8 * ra + g];
#pragma unroll
        for (int rb = 0; rb < RB; ++rb)
          if (rb >= ra) {
            const int j0 = 8 * rb + 2 * tq;
            c[ra][rb][0] = fma(av, BT ? Bm[j0 * ldb + k] : Bm[k * ldb + j0], c[ra][rb][0]);
            c[ra][rb][1] = fma(av, BT ? Bm[(j0 + 1) * ldb + k] : Bm[k * ldb + j0 + 1], c[ra][rb][1]);
          }
      }
    }
  }
#pragma unroll
  for (int ra = 0; ra < RA; ++ra)
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
      if (rb >= ra) epi(8 * ra + g, 8 * rb + 2 * tq, c[ra][rb][0], c[ra][rb][1], pv[ra][rb]);
#else
  (void)w;
  for (int i = 0; i < 8 * RA; ++i)
    for (int j = 0; j < 8 * RB; j += 2) {
      if (j / 8 < i / 8) continue;
      R pv[NP];
      for (int q = 0; q < NP; ++q) pv[q] = R(0);
      pre(i, j, pv);
      R v0 = R(0), v1 = R(0);
      for (int k = 0; k < 4 * K4; ++k) {
        const R av = AT ? A[i * lda + k] : A[k * lda + i];
        v0 = fma(av, BT ? Bm[j * ldb + k] : Bm[k * ldb + j], v0);
        v1 = fma(av, BT ? Bm[(j + 1) * ldb + k] : Bm[k * ldb + j + 1], v1);
      }
      epi(i, j, v0, v1, pv);
    }
#endif
}

template <typename R, int RA, int RB, bool AT = false, bool BT = false, class Epi>
DONK_D void warp_mma(const WCtx& w, const R* __restrict__ A, int lda, const R* __restrict__ Bm, int ldb, int K4,
                     Epi&& epi) {
  warp_mma_pre<R, RA, RB, 1, AT, BT>(w, A, lda, Bm, ldb, K4, [](int, int, R*) {},
                             [&](int i, int j, R v0, R v1, const R*) { epi(i, j, v0, v1); });
}

// Sum of x over the lanes of the warp, added to *slot (shared memory) by lane 0.  Under emulation the lanes of
// a LANE_FOR run one after the other, so each simply adds its value.
template <typename R>
DONK_D void team_accumulate(const WCtx& w, R* slot, R x) {
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  if (w.lane == 0) *slot += x;
#else
  (void)w;
  *slot += x;
#endif
}

// hint: bring the 128-byte line at p into L2 (no-op under emulation)
DONK_D void prefetch_l2(const void* p) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

// two adjacent shared-memory values at a 16-byte aligned address: one STS.128 / LDS.128
template <typename R>
DONK_HD void st2(R* p, R v0, R v1) {
#if defined(__CUDA_ARCH__)
  if (sizeof(R) == 8) {
    *reinterpret_cast<double2*>(p) = make_double2((double)v0, (double)v1);
    return;
  }
#endif
  p[0] = v0; p[1] = v1;
}
template <typename R>
DONK_HD void ld2(const R* p, R& v0, R& v1) {
#if defined(__CUDA_ARCH__)
  if (sizeof(R) == 8) {
    const double2 t = *reinterpret_cast<const double2*>(p);
    v0 = (R)t.x; v1 = (R)t.y;
    return;
  }
#endif
  v0 = p[0]; v1 = p[1];
}
// The same pair, for rows that are `ld` apart with ld % 16 in {4, 12} (the fragment-friendly leading dimensions):
// a 16-byte access per lane would put the two rows of a quarter-warp on overlapping banks (2-way conflict on every
// accumulator store).  Two 8-byte accesses whose order alternates with the row parity touch every bank exactly once.
// Measured (B200, config 3): the conflict-free form is SLOWER (58.8 vs 53.9 ms per launch) - the kernel is bound by
// issued instructions, not by shared-memory wavefronts - so it is off unless DONK_PAIR_SPLIT is defined.
template <typename R>
DONK_HD void st2s(R* p, R v0, R v1, int row) {
#if defined(__CUDA_ARCH__) && !defined(DONK_PAIR_SPLIT)
  (void)row;
  st2(p, v0, v1);
#elif defined(__CUDA_ARCH__)
  const int odd = row & 1;
  p[odd] = odd ? v1 : v0;
  p[odd ^ 1] = odd ? v0 : v1;
#else
  (void)row;
  p[0] = v0; p[1] = v1;
#endif
}
template <typename R>
DONK_HD void ld2s(const R* p, R& v0, R& v1, int row) {
#if defined(__CUDA_ARCH__) && !defined(DONK_PAIR_SPLIT)
  (void)row;
  ld2(p, v0, v1);
#elif defined(__CUDA_ARCH__)
  const int odd = row & 1;
  const R a = p[odd], b = p[odd ^ 1];
  v0 = odd ? b : a;
  v1 = odd ? a : b;
#else
  (void)row;
  v0 = p[0]; v1 = p[1];
#endif
}

// *slot += sum_{r < n} f(r), n <= 32: one item per lane on the device (every lane joins the shuffle), a plain loop
// under emulation.  The caller synchronises the team before reading *slot.
template <typename R, class F>
DONK_D void team_reduce_add(const WCtx& w, R* slot, int n, F&& f) {
#if !defined(DONK_EMU)
  R x = (int)w.lane < n ? f((int)w.lane) : R(0);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  if (w.lane == 0) *slot += x;
#else
  (void)w;
  for (int r = 0; r < n; ++r) *slot += f(r);
#endif
}

// two adjacent read-only values; one 16-byte load when the address is known to be 16-byte aligned
template <typename R>
DONK_HD void ldg2(const R* p, bool aligned16, R& v0, R& v1) {
#if defined(__CUDA_ARCH__)
  if (sizeof(R) == 8 && aligned16) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(p));
    v0 = (R)t.x; v1 = (R)t.y;
    return;
  }
#endif
  (void)aligned16;
  v0 = ldg(p); v1 = ldg(p + 1);
}

// smallest leading dimension >= n that is a multiple of 4 and makes DMMA fragment loads conflict free
DONK_HD constexpr int ld_pad(int n) {
  int m = (n + 3) / 4 * 4;
  return m % 8 == 0 ? m + 4 : m;
}

// 8-byte (or 4-byte for float) asynchronous global -> shared copy (LDGSTS); plain copy under emulation.
template <typename R>
DONK_D void cp_async(R* dst, const R* src) {
#if defined(__CUDA_ARCH__)
  const unsigned s = (unsigned)__cvta_generic_to_shared(dst);
  if (sizeof(R) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(src) : "memory");
#else
  *dst = *src;
#endif
}
// 16-byte variant (two adjacent reals of a double row; both addresses 16-byte aligned)
template <typename R>
DONK_D void cp_async_pair(R* dst, const R* src) {
#if defined(__CUDA_ARCH__)
  const unsigned s = (unsigned)__cvta_generic_to_shared(dst);
  if (sizeof(R) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(src) : "memory");
#else
  dst[0] = src[0]; dst[1] = src[1];
#endif
}
DONK_D void cp_async_wait_all() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#endif
}
// close the copies issued so far into one group / wait until at most N of this thread's groups are still pending
// (groups complete in issue order): lets a later group (next step's shared inputs) stay in flight while an earlier one
// (this step's policy) is consumed
DONK_D void cp_async_commit() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
DONK_D void cp_async_wait_group() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

// ---- bulk asynchronous copies (the TMA engine's 1-D form, UBLKCP) with an mbarrier as completion object -------------
// One instruction moves a whole contiguous row (global -> shared, both 16-byte aligned, size a multiple of 16) and
// reports its bytes to the mbarrier; the consumers wait on the barrier's phase.  Under emulation: a plain copy.
DONK_D void mbar_init(unsigned long long* bar, int arrivals) {
#if defined(__CUDA_ARCH__)
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(arrivals) : "memory");
#else
  (void)bar; (void)arrivals;
#endif
}
// the single expected arrival of a phase, announcing how many bytes the bulk copies of the phase will deliver
DONK_D void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
#if defined(__CUDA_ARCH__)
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
#else
  (void)bar; (void)bytes;
#endif
}
DONK_D void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst)),
               "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
#else
  (void)bar;
  memcpy(dst, src, bytes);
#endif
}
// wait for the phase with the given parity; a protocol error (byte count mismatch) traps instead of hanging the GPU
DONK_D void mbar_wait(unsigned long long* bar, unsigned parity) {
#if defined(__CUDA_ARCH__)
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  for (unsigned spin = 0;; ++spin) {
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (spin > (1u << 24)) __trap();
  }
#else
  (void)bar; (void)parity;
#endif
}
// orders earlier generic-proxy writes to shared memory (zero fill) before later async-proxy writes (bulk copies)
DONK_D void fence_proxy_async() {
#if defined(__CUDA_ARCH__)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

// Upper-triangular tile enumeration: idx in [0, nt(nt+1)/2) -> (ti <= tj).
DONK_HD void tri_tile(int idx, int nt, int& ti, int& tj) {
  int t = 0;
  while (idx >= nt - t) {
    idx -= nt - t;
    ++t;
  }
  ti = t;
  tj = t + idx;
}

}  // namespace donk

// status codes returned by the C-ABI (same values as include/donk_b200.h)
#ifndef DONK_OK
#define DONK_OK 0
#define DONK_BAD_SHAPE 1
#define DONK_NOT_PD 2
#define DONK_CUDA_ERROR 3
#define DONK_SMEM_LIMIT 4
#endif
