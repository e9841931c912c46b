// donk_b200 — building blocks of the CTA-cooperative kernels (ilqr_coop.cuh, fitlr_coop.cuh): register-tile products on
// the FP64 tensor path (mma.sync.m8n8k4.f64) and the longest-tile-first schedule that deals the tiles of a product to
// the warps of a CTA.
#pragma once
#include "common.cuh"

namespace donk {

// ---------------------------------------------------------------------------------------------------------------------
// Register-tile product on the FP64 tensor path, RA x RB fragments of 8 x 8:
//   C[i][j] = sum_{k < 4 K4} A(k, i) B(k, j),  i < 8 RA, j < 8 RB;  with TRI only the fragments ra <= rb (a tile whose
//   origin lies on the diagonal of a symmetric result).
// Operand storage: k-major (A[k*lda + i]) or, with AT / BT, the natural orientation (A[i*lda + k]).  pre(i, j, pv)
// loads up to NP epilogue addends per fragment lane BEFORE the product (raw loads that stay in flight behind it),
// epi(i, j, v0, v1, pv) receives columns j, j+1 of row i.  Everything about the shape is a compile-time constant: a
// fragment set chosen at run time would put every mma.sync behind a predicate the compiler cannot prove warp-uniform
// (one WARPSYNC + NOP per DMMA in the SASS of the first version of this kernel).
// ---------------------------------------------------------------------------------------------------------------------
template <typename R, int RA, int RB, int NP, bool AT, bool BT, bool TRI, class Pre, class Epi>
DONK_D void tile_mma(const WCtx& w, const R* __restrict__ A, int lda, const R* __restrict__ Bm, int ldb, int K4,
                     Pre&& pre, Epi&& epi) {
#if defined(__CUDA_ARCH__)
  static_assert(sizeof(R) == 8, "tile_mma: double only");
  const int g = w.lane >> 2, tq = w.lane & 3;
  R c[RA][RB][2], pv[RA][RB][NP > 0 ? NP : 1];
#pragma unroll
  for (int ra = 0; ra < RA; ++ra)
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      if (TRI && ra > rb) continue;
      c[ra][rb][0] = R(0); c[ra][rb][1] = R(0);
#pragma unroll
      for (int q = 0; q < (NP > 0 ? NP : 1); ++q) pv[ra][rb][q] = R(0);
      if (NP > 0) pre(8 * ra + g, 8 * rb + 2 * tq, pv[ra][rb]);
    }
  const R* ap = AT ? A + g * lda + tq : A + tq * lda + g;
  const R* bp = BT ? Bm + g * ldb + tq : Bm + tq * ldb + g;
  const int sa = AT ? 8 * lda : 8, sb = BT ? 8 * ldb : 8;
  const int da = AT ? 4 : 4 * lda, db = BT ? 4 : 4 * ldb;
  // software pipeline: the fragments of k-step k4 + 1 are loaded before the products of k-step k4 are issued (two register
  // sets, the loop advances two k-steps per trip so that the set index is a compile-time constant; K4 is even)
  R af[2][RA], bf[2][RB];
#pragma unroll
  for (int ra = 0; ra < RA; ++ra) af[0][ra] = ap[sa * ra];
#pragma unroll
  for (int rb = 0; rb < RB; ++rb) bf[0][rb] = bp[sb * rb];
#pragma unroll 2
  for (int k4 = 0; k4 < K4; k4 += 2) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      ap += da;
      bp += db;
      if (h == 0 || k4 + 2 < K4) {  // the other register set <- next k-step
#pragma unroll
        for (int ra = 0; ra < RA; ++ra) af[h ^ 1][ra] = ap[sa * ra];
#pragma unroll
        for (int rb = 0; rb < RB; ++rb) bf[h ^ 1][rb] = bp[sb * rb];
      }
#pragma unroll
      for (int ra = 0; ra < RA; ++ra)
#pragma unroll
        for (int rb = 0; rb < RB; ++rb) {
          if (TRI && ra > rb) continue;
          asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
              : "+d"(c[ra][rb][0]), "+d"(c[ra][rb][1])
              : "d"(af[h][ra]), "d"(bf[h][rb]));
        }
    }
  }
#pragma unroll
  for (int ra = 0; ra < RA; ++ra)
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      if (TRI && ra > rb) continue;
      epi(8 * ra + g, 8 * rb + 2 * tq, c[ra][rb][0], c[ra][rb][1], pv[ra][rb]);
    }
#else
  (void)w;
  for (int ra = 0; ra < RA; ++ra)
    for (int rb = 0; rb < RB; ++rb) {
      if (TRI && ra > rb) continue;
      for (int ii = 0; ii < 8; ++ii)
        for (int jj = 0; jj < 8; jj += 2) {
          const int i = 8 * ra + ii, j = 8 * rb + jj;
          R pv[NP > 0 ? NP : 1];
          for (int q = 0; q < (NP > 0 ? NP : 1); ++q) pv[q] = R(0);
          if (NP > 0) pre(i, j, pv);
          R v0 = R(0), v1 = R(0);
          for (int k = 0; k < 4 * K4; ++k) {
            const R av = AT ? A[i * lda + k] : A[k * lda + i];
            v0 = fma(av, BT ? Bm[j * ldb + k] : Bm[k * ldb + j], v0);
            v1 = fma(av, BT ? Bm[(j + 1) * ldb + k] : Bm[k * ldb + j + 1], v1);
          }
          epi(i, j, v0, v1, pv);
        }
    }
#endif
}

// The same with the number of block columns (1..RB) and the diagonal flag chosen at run time, CTA-uniformly per tile:
// dispatches to the compile-time shapes (the last tile of a block row is narrower; the first one of a symmetric result
// sits on the diagonal).
template <typename R, int RA, int RB, int NP, bool AT, bool BT, bool CAN_TRI, class Pre, class Epi>
DONK_D void tile_mma_n(const WCtx& w, int nb, bool tri, const R* __restrict__ A, int lda, const R* __restrict__ Bm, int ldb,
                       int K4, Pre&& pre, Epi&& epi) {
  if (nb >= RB) {
    if (CAN_TRI && tri) tile_mma<R, RA, RB, NP, AT, BT, CAN_TRI>(w, A, lda, Bm, ldb, K4, pre, epi);
    else tile_mma<R, RA, RB, NP, AT, BT, false>(w, A, lda, Bm, ldb, K4, pre, epi);
  } else if constexpr (RB > 1) {
    tile_mma_n<R, RA, RB - 1, NP, AT, BT, CAN_TRI>(w, nb, tri, A, lda, Bm, ldb, K4, pre, epi);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tile schedule: the tiles of a product (RA x RB fragment blocks on a grid, ragged for triangular products) are
// assigned longest-first to the least loaded warp.  Computed once per CTA by one thread (the dimensions are
// compile-time constants, so is the result); stored as (ra0, rb0) bytes per (product, warp, slot).
// ---------------------------------------------------------------------------------------------------------------------
struct CoopSched {
  unsigned char n[8][16];          // tiles of (product, warp)
  unsigned char t[8][16][12][2];   // (ra0, rb0)
};

// number of active fragments of the tile at (ra0, rb0): block rows < na, block columns < nb, upper triangle if tri
DONK_HD int coop_tile_cost(int ra0, int rb0, int RA, int RB, int na, int nb, bool tri) {
  int cnt = 0;
  for (int ra = ra0; ra < ra0 + RA && ra < na; ++ra)
    for (int rb = rb0; rb < rb0 + RB && rb < nb; ++rb)
      if (!tri || ra <= rb) ++cnt;
  return cnt;
}

// rows [0, na) x columns [0, nb) of blocks.  With tri the first tile of a block row starts at the diagonal block.
// skip_sub0: the warps of SM sub-partition 0 (warp 0, 4, 8, 12) take no tile - in B2 warp 0 runs the serial
// factorisation of Quu, whose dependent DFMAs would queue behind the DMMAs of a neighbour on the same FP64 pipe
// (measured: the chain took 2-3x its latency-bound time with warp 4 multiplying next to it).
DONK_HD void coop_schedule(CoopSched& s, int ph, int G, int RA, int RB, int na, int nb, bool tri, bool skip_sub0) {
  // Warps g and g + 4 issue from the same SM sub-partition (one FP64 tensor pipe each): the load that is balanced is
  // the sub-partition's, then the warp's.
  int load[16], sub[4];
  for (int g = 0; g < 16; ++g) { load[g] = 0; s.n[ph][g] = 0; }
  for (int q = 0; q < 4; ++q) sub[q] = 0;
  auto usable = [&](int g) { return !skip_sub0 || (G > 4 ? (g & 3) != 0 : (g != 0 || G == 1)); };
  // collect tiles
  unsigned char tl[192][3];
  int nt = 0;
  for (int ra0 = 0; ra0 < na; ra0 += RA) {
    const int c0 = tri ? ra0 : 0;  // first block column with an active fragment in these rows
    for (int rb0 = c0; rb0 < nb; rb0 += RB) {
      const int cost = coop_tile_cost(ra0, rb0, RA, RB, na, nb, tri);
      if (cost == 0 || nt >= 192) continue;
      tl[nt][0] = (unsigned char)ra0; tl[nt][1] = (unsigned char)rb0; tl[nt][2] = (unsigned char)cost;
      ++nt;
    }
  }
  // longest first onto the least loaded sub-partition / warp (stable: deterministic)
  for (int round = 0; round < nt; ++round) {
    int best = -1;
    for (int i = 0; i < nt; ++i)
      if (tl[i][2] != 0 && (best < 0 || tl[i][2] > tl[best][2])) best = i;
    int gw = -1;
    for (int g = 0; g < G; ++g) {
      if (!usable(g)) continue;
      if (gw < 0 || sub[g & 3] < sub[gw & 3] || (sub[g & 3] == sub[gw & 3] && load[g] < load[gw])) gw = g;
    }
    const int slot = s.n[ph][gw];
    if (slot < 12) {
      s.t[ph][gw][slot][0] = tl[best][0];
      s.t[ph][gw][slot][1] = tl[best][1];
      s.n[ph][gw] = (unsigned char)(slot + 1);
    }
    load[gw] += tl[best][2];
    sub[gw & 3] += tl[best][2];
    tl[best][2] = 0;
  }
}

}  // namespace donk
