// donk_b200 — GMM EM, E and M passes on the FP64 tensor path (mma.sync.m8n8k4.f64).
//
// Same iteration as gmm.cuh (sklearn.mixture.GaussianMixture as called from GMMPrior.update,
// donk/dynamics/prior/prior_gmm.py:14-24); these are the fast kernels, the scalar-FMA ones in gmm.cuh stay
// as the generic fallback (d > 144) and as the emulation reference.
//
//   gmm_e_mma_cta  E pass.  CTA = G warps, tile of 8*RA*G rows staged feature-major (k-major A operand);
//                  per cluster y = x P_k + sP_k as DMMA fragments against the upper-triangular P_k read
//                  through L1 (all warps of the CTA walk the same P_k), only blocks on or above the
//                  diagonal are multiplied; row sums of y^2 by a 4-lane shuffle; then softmax, resp, lse.
//   gmm_m_mma_cta  M pass.  CTA = (group of KC clusters, row range).  Rows are staged once, shifted by the
//                  mixture mean x_bar (cluster independent, keeps cancellation at the cluster-separation
//                  level); the scatter sum_n r_nk (x-x_bar)(x-x_bar)^T is accumulated as DMMA fragments that stay in
//                  registers for the whole range: the sample index is the k dimension, the A fragment is
//                  scaled by r_nk in registers, the B fragment is shared by the KC clusters.
//                  sum r (x - x_bar) and sum r come from a small FMA side loop.
#pragma once
#include "common.cuh"
#include "gmm.cuh"

namespace donk {

#ifndef DONK_E_JG_WIDE
#define DONK_E_JG_WIDE 6
#endif

// ------------------------------------------------------------------------------------------------------
// E pass
// ------------------------------------------------------------------------------------------------------
template <int NB, int RA>
DONK_HD size_t gmm_e_mma_smem(int K, int G) {
  const int TS = 8 * RA * G, ldt = ld_pad(TS);
  return (size_t)8 * NB * ldt + (size_t)K * 8 * NB + round_up(K, 4) + (size_t)TS * (K | 1) + 8;
}

template <typename R, int NB, int RA>
DONK_D void gmm_e_mma_cta(const WCtx& w, const GmmArgs<R>& a, int cta, int ncta, R* sm) {
  const int d = a.d, K = a.K, G = w.G;
  const int TS = 8 * RA * G, ldt = ld_pad(TS), D8 = 8 * NB;
  R* Xt = sm;                       // [D8][ldt]   feature-major rows of the tile
  R* sP = Xt + (size_t)D8 * ldt;    // [K][D8]     -(mu_k P_k)
  R* ldet = sP + (size_t)K * D8;    // [K]         sum log diag P_k + log pi_k
  R* logp = ldet + round_up(K, 4);  // [TS][KP]: odd row pitch (a thread pair per row in the softmax: no bank conflicts)
  const int KP = K | 1;
  const size_t dd = (size_t)d * d;
  CTA_FOR(idx, D8 * ldt) Xt[idx] = R(0);
  CTA_FOR(idx, K * D8) {
    const int k = idx / D8, j = idx % D8;
    R acc = R(0);
    if (j < d) {
      const R* P = a.pc + (size_t)k * dd;
      for (int i = 0; i <= j; ++i) acc = fma(ldg(a.means + (size_t)k * d + i), ldg(P + (size_t)i * d + j), acc);
    }
    sP[idx] = -acc;
  }
  CTA_FOR(k, K) {
    const R* P = a.pc + (size_t)k * dd;
    R acc = R(0);
    for (int i = 0; i < d; ++i) acc += log(ldg(P + (size_t)i * d + i));
    ldet[k] = acc + log(ldg(a.w + k));
  }
  w.cta_sync();
  const R cst = R(d) * R(1.8378770664093453);  // d log(2 pi)
  for (int blk = cta; blk < a.nblocks; blk += ncta) {
    const size_t n0 = (size_t)blk * TS;
    // element idx = r d + j of the tile, consecutive threads on consecutive elements (coalesced).  The device form steps
    // (r, j) - and for the transition view (sample, t) - incrementally: idx / d, idx % d (and row / T) per element cost
    // ~60 instructions each, 14 % of the issue slots at d = 72, K = 16 and more with fewer clusters
#if defined(DONK_EMU)
    CTA_FOR(idx, TS * d) {
      const int r = idx / d, j = idx % d;
      Xt[j * ldt + r] = n0 + r < (size_t)a.n ? gmm_row_item(a, n0 + r, j) : R(0);
    }
#else
    {
      const int NT = w.nthreads, qd = NT / d, rd = NT % d;
      int r = w.tid / d, j = w.tid % d;
      if (a.tX == nullptr) {
        const R* base = a.X + n0 * d;
        const int rows = a.n - n0 < (size_t)TS ? (int)(a.n - n0) : TS;
#pragma unroll 4
        for (int idx = w.tid; idx < TS * d; idx += NT) {
          Xt[j * ldt + r] = r < rows ? ldg(base + idx) : R(0);
          r += qd; j += rd;
          if (j >= d) { j -= d; ++r; }
        }
      } else {
        const unsigned T = (unsigned)a.tT;
        unsigned sm = (unsigned)(n0 + r) / T, t = (unsigned)(n0 + r) - sm * T;  // sample (c N + n), time step of row n0 + r
        const int dX = a.tdX, p = a.tdX + a.tdU;
#pragma unroll 2
        for (int idx = w.tid; idx < TS * d; idx += NT) {
          R v = R(0);
          if (n0 + r < (size_t)a.n) {
            const size_t xrow = (size_t)sm * (T + 1) + t;  // x_t and x_{t+1} are adjacent in the roll-outs
            v = j < dX ? ldg(a.tX + xrow * dX + j)
                       : (j < p ? ldg(a.tU + ((size_t)sm * T + t) * a.tdU + (j - dX)) : ldg(a.tX + xrow * dX + (j - a.tdU)));
          }
          Xt[j * ldt + r] = v;
          int dr = qd;
          j += rd;
          if (j >= d) { j -= d; ++dr; }
          r += dr;
          t += dr;
          while (t >= T) { t -= T; ++sm; }
        }
      }
    }
#endif
    w.cta_sync();
    for (int k = 0; k < K; ++k) {
      const R* P = a.pc + (size_t)k * dd;
      TEAMS_DO(tm) {
        const int i0 = 8 * RA * tm;
#if !defined(DONK_EMU)
        const int g = w.lane >> 2, tq = w.lane & 3;
        const R* pfk = a.pf + (size_t)k * (NB * (NB + 1) / 2) * 64;  // the device form always has the copy
        // Column blocks in groups of JG: the accumulators of a group (RA x JG fragments) fit the registers of a
        // two-CTA-per-SM launch without spilling; the A fragments of the rows a group needs are re-read per group.
        constexpr int JG = NB >= 12 ? DONK_E_JG_WIDE : 3;  // wide kernels (d > 96) run one CTA per SM with 255 registers: wider groups, fewer A re-reads
        R ssr[RA];
#pragma unroll
        for (int ra = 0; ra < RA; ++ra) ssr[ra] = R(0);
#pragma unroll
        for (int j0 = 0; j0 < NB; j0 += JG) {
          R c[RA][JG][2];
#pragma unroll
          for (int ra = 0; ra < RA; ++ra)
#pragma unroll
            for (int jj = 0; jj < JG; ++jj) { c[ra][jj][0] = R(0); c[ra][jj][1] = R(0); }
#pragma unroll
          for (int fb = 0; fb < NB; ++fb) {
            if (fb >= j0 + JG) continue;  // P_k is upper triangular: no block of this group lies on or above row fb
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int f = 8 * fb + 4 * h + tq;  // feature index = k dimension of the product
              R af[RA];
#pragma unroll
              for (int ra = 0; ra < RA; ++ra) af[ra] = Xt[f * ldt + i0 + 8 * ra + g];
#pragma unroll
              for (int jj = 0; jj < JG; ++jj) {
                const int jb = j0 + jj;
                if (jb >= NB || jb < fb) continue;  // blocks below the diagonal are zero
                // B fragment from the fragment-major copy the launcher made: one contiguous 256-byte run per warp (two
                // L1 wavefronts instead of four 64-byte rows of P_k, no predicate, no address arithmetic)
                const R bf = ldg(pfk + ((gmm_pfrag_tile(NB, fb, jb) * 2 + h) * 32 + w.lane));
#pragma unroll
                for (int ra = 0; ra < RA; ++ra) {
                  double d0 = (double)c[ra][jj][0], d1 = (double)c[ra][jj][1];
                  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                      : "+d"(d0), "+d"(d1)
                      : "d"((double)af[ra]), "d"((double)bf));
                  c[ra][jj][0] = (R)d0;
                  c[ra][jj][1] = (R)d1;
                }
              }
            }
          }
#pragma unroll
          for (int ra = 0; ra < RA; ++ra)
#pragma unroll
            for (int jj = 0; jj < JG; ++jj) {
              if (j0 + jj >= NB) continue;
              const int j = 8 * (j0 + jj) + 2 * tq;
              const R y0 = c[ra][jj][0] + sP[k * D8 + j], y1 = c[ra][jj][1] + sP[k * D8 + j + 1];
              ssr[ra] = fma(y0, y0, ssr[ra]);  // padded columns: c = 0 and sP = 0
              ssr[ra] = fma(y1, y1, ssr[ra]);
            }
        }
#pragma unroll
        for (int ra = 0; ra < RA; ++ra) {
          R ss = ssr[ra];
          ss += __shfl_xor_sync(0xffffffffu, ss, 1);
          ss += __shfl_xor_sync(0xffffffffu, ss, 2);
          if (tq == 0) logp[(i0 + 8 * ra + g) * KP + k] = -(cst + ss) / 2 + ldet[k];
        }
#else
        for (int r = 0; r < 8 * RA; ++r) {
          R ss = R(0);
          for (int j = 0; j < d; ++j) {
            R y = sP[k * D8 + j];
            for (int f = 0; f <= j; ++f) y = fma(Xt[f * ldt + i0 + r], P[(size_t)f * d + j], y);
            ss = fma(y, y, ss);
          }
          logp[(i0 + r) * KP + k] = -(cst + ss) / 2 + ldet[k];
        }
#endif
      }
    }
    w.cta_sync();
#if !defined(DONK_EMU)
    {
      // responsibilities (softmax over the clusters, sklearn/mixture/_base.py:552-582): two threads per row, each
      // taking every other cluster (TS <= blockDim / 2); exp is evaluated once per entry (resp = e_k / sum e_k);
      // the tile's sum of log-sum-exp is reduced by warp shuffles instead of one thread walking the rows
      const int r = w.tid >> 1, h = w.tid & 1;
      const bool act = r < TS;  // RA = 1 tiles (d > 72) have fewer rows than thread pairs
      const bool live = act && n0 + r < (size_t)a.n;
      R* lp = logp + (act ? r : 0) * KP;
      R m = -INFINITY;
      if (act)
        for (int k = h; k < K; k += 2) m = lp[k] > m ? lp[k] : m;
      const R mo = __shfl_xor_sync(0xffffffffu, m, 1);
      m = mo > m ? mo : m;
      R sum = R(0);
      if (act)
        for (int k = h; k < K; k += 2) {
          const R e = exp(lp[k] - m);
          lp[k] = e;
          sum += e;
        }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      const R inv = act ? R(1) / sum : R(0);
      if (live)
        for (int k = h; k < K; k += 2) a.resp[(n0 + r) * K + k] = lp[k] * inv;
      R lse = (live && h == 0) ? m + log(sum) : R(0);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lse += __shfl_xor_sync(0xffffffffu, lse, o);
      if (w.lane == 0) logp[TS * KP + w.team] = lse;  // the 8 spare slots behind the tile
    }
    w.cta_sync();
    if (a.lse_part) {
      CTA_FOR(one, 1) {
        R acc = R(0);
        for (int wp = 0; wp < G; ++wp) acc += logp[TS * KP + wp];
        a.lse_part[blk] = acc;
      }
    }
#else
    CTA_FOR(r, TS) {
      if (n0 + r >= (size_t)a.n) { logp[r * KP] = R(0); continue; }
      R* lp = logp + r * KP;
      R m = lp[0];
      for (int k = 1; k < K; ++k) m = lp[k] > m ? lp[k] : m;
      R s = R(0);
      for (int k = 0; k < K; ++k) s += exp(lp[k] - m);
      const R lse = m + log(s);
      for (int k = 0; k < K; ++k) a.resp[(n0 + r) * K + k] = exp(lp[k] - lse);
      lp[0] = lse;
    }
    w.cta_sync();
    if (a.lse_part) {
      CTA_FOR(one, 1) {
        R acc = R(0);
        for (int r = 0; r < TS; ++r) acc += logp[r * KP];
        a.lse_part[blk] = acc;
      }
    }
#endif
    w.cta_sync();
  }
}

// ------------------------------------------------------------------------------------------------------
// M pass
// ------------------------------------------------------------------------------------------------------
enum { GMM_MTS = 64 };  // rows per staged tile of the M pass

DONK_HD size_t gmm_m_mma_smem(int d, int KC) {
  const int D8 = round_up(d, 8), ld = ld_pad(D8);
  return (size_t)2 * GMM_MTS * ld + (size_t)2 * GMM_MTS * KC + D8 + 8 + 2;  // + two mbarriers
}

// block q of the row-major enumeration of the upper triangle of an nb x nb block grid
DONK_HD void tri_block(int q, int nb, int& ib, int& jb) {
  int r = 0;
  while (q >= nb - r) { q -= nb - r; ++r; }
  ib = r; jb = r + q;
}

template <typename R, int NBW, int KC>
DONK_D void gmm_m_mma_cta(const WCtx& w, const GmmArgs<R>& a, int cta, R* sm) {
  const int d = a.d, K = a.K, G = w.G, TS = GMM_MTS;
  const int nb = (d + 7) / 8, D8 = 8 * nb, ld = ld_pad(D8), nblk = nb * (nb + 1) / 2;
  const int ngroups = (K + KC - 1) / KC;
  const int kg = cta / a.nch, ch = cta % a.nch;
  (void)ngroups;
  R* Xs = sm;                                   // [2][TS][ld]  rows minus x_bar (double buffered)
  R* rr = Xs + (size_t)2 * TS * ld;             // [2][TS][KC]  responsibilities of the KC clusters
  R* xbar = rr + (size_t)2 * TS * KC;           // [D8]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(xbar + D8 + 8);  // completion of the row copies
  // rows are 16-byte aligned multiples of 16 bytes: each row of a tile is ONE bulk copy (TMA engine, 1-D form)
  const bool bulk = sizeof(R) == 8 && d % 2 == 0 && (reinterpret_cast<size_t>(a.X) & 15) == 0;
  const size_t rows_per = ((size_t)a.n + a.nch - 1) / a.nch;
  const size_t row0 = (size_t)ch * rows_per;
  const size_t row1 = row0 + rows_per < (size_t)a.n ? row0 + rows_per : (size_t)a.n;
  const size_t pl = gmm_part_len(d);

  CTA_FOR(j, D8) {  // x_bar = sum_k pi_k mu_k: the common centring point of this iteration
    R acc = R(0);
    if (j < d)
      for (int k = 0; k < K; ++k) acc = fma(ldg(a.w + k), ldg(a.means + (size_t)k * d + j), acc);
    xbar[j] = acc;
  }
  CTA_FOR(idx, 2 * TS * ld) Xs[idx] = R(0);
  CTA_FOR(i, 2) mbar_init(bars + i, 1);
  w.cta_sync();
  fence_proxy_async();  // the zero fill above (generic proxy) precedes the bulk copies (async proxy) into the same rows
  if (cta == 0 && a.shift_out) {
    CTA_FOR(idx, K * d) a.shift_out[idx] = xbar[idx % d];
  }

  // block range of every warp (contiguous in row-major order: consecutive blocks share their A fragment).  The
  // remainder goes to the LAST warps: the first ones also carry the first-moment sums (threads 0..d), and the
  // barrier of every tile waits for the slowest warp (45 blocks over 8 warps: 5,5,5,6,6,6,6,6 instead of 6 x 7 + 3)
  const int base = nblk / G, rem = nblk % G;

#if !defined(DONK_EMU)
  const int g = w.lane >> 2, tq = w.lane & 3;
  const int extra = w.team - (G - rem);
  const int per = base + (extra >= 0 ? 1 : 0);
  const int q0 = w.team * base + (extra > 0 ? extra : 0);
  int ib[NBW], jb[NBW];
  bool ok[NBW];
  R acc[NBW][KC][2];
#pragma unroll
  for (int i = 0; i < NBW; ++i) {
    ok[i] = i < per && q0 + i < nblk;
    ib[i] = 0; jb[i] = 0;
    if (ok[i]) tri_block(q0 + i, nb, ib[i], jb[i]);
#pragma unroll
    for (int c = 0; c < KC; ++c) { acc[i][c][0] = R(0); acc[i][c][1] = R(0); }
  }
  R xbA[NBW], xbB[NBW];  // x_bar of the lane's fragment columns
  int offA[NBW], offB[NBW];
#pragma unroll
  for (int i = 0; i < NBW; ++i) {
    offA[i] = 8 * ib[i] + g; offB[i] = 8 * jb[i] + g;
    xbA[i] = xbar[offA[i]]; xbB[i] = xbar[offB[i]];
  }
  // First moments and cluster sizes: column sj <= d (sj == d: the column of ones) over the rows ssub, ssub + nsub, ... of
  // every tile, ALL threads of the CTA taking part (the first d + 1 threads alone walking all 64 rows put ~14 % more
  // FP64 work on their three warps, and every tile barrier waited for them: 25 % of the stall samples of the kernel)
  const int nsub = w.nthreads / (d + 1), ssub = w.tid / (d + 1), sj = w.tid % (d + 1);
  const bool side_on = ssub < nsub;
  const R side_xb = (side_on && sj < d) ? xbar[sj] : R(0);
  R side[KC];  // partial sums of r_nc (x - x_bar)[sj] (sj < d) / of r_nc (sj == d)
#pragma unroll
  for (int c = 0; c < KC; ++c) side[c] = R(0);
#else
  std::vector<R> accS((size_t)KC * d * d, R(0)), accs((size_t)KC * d, R(0)), accn(KC, R(0));
#endif

  auto stage = [&](size_t n0, int buf) {
    R* xs = Xs + (size_t)buf * TS * ld;
    R* rk = rr + (size_t)buf * TS * KC;
    if (bulk) {
      const size_t left = row1 - n0;
      const int valid = left < (size_t)TS ? (int)left : TS;
      CTA_FOR(one, 1) mbar_arrive_expect_tx(bars + buf, (unsigned)((size_t)valid * d * sizeof(R)));
      CTA_FOR(r, TS) {
        if (r < valid) {
          bulk_copy_g2s(xs + r * ld, a.X + (n0 + r) * d, (unsigned)(d * sizeof(R)), bars + buf);
        } else {
          for (int j = 0; j < d; ++j) xs[r * ld + j] = R(0);
        }
      }
    } else {
      CTA_FOR(idx, TS * d) {
        const int r = idx / d, j = idx % d;
        if (n0 + r < row1) cp_async(xs + r * ld + j, a.X + (n0 + r) * d + j);
        else xs[r * ld + j] = R(0);
      }
    }
    CTA_FOR(idx, TS * KC) {
      const int r = idx / KC, c = idx % KC;
      const int k = kg * KC + c;
      if (n0 + r < row1 && k < K) cp_async(rk + idx, a.resp + (n0 + r) * K + k);
      else rk[idx] = R(0);
    }
  };

  int buf = 0;
  unsigned tile = 0;
  if (row0 < row1) stage(row0, 0);
  for (size_t n0 = row0; n0 < row1; n0 += TS, ++tile) {
    cp_async_wait_all();
    if (bulk) mbar_wait(bars + buf, (tile >> 1) & 1);  // the rows of this tile have landed
    w.cta_sync();
    if (n0 + TS < row1) stage(n0 + TS, buf ^ 1);  // next tile lands while this one is multiplied
    R* xs = Xs + (size_t)buf * TS * ld;
    const R* rk = rr + (size_t)buf * TS * KC;
#if !defined(DONK_EMU)
    // the tile is centred on x_bar in registers (rows beyond the range carry r = 0 and contribute nothing)
#pragma unroll 1
    for (int k4 = 0; k4 < TS / 4; ++k4) {
      const R* xrow = xs + (4 * k4 + tq) * ld;
      // all fragment loads of the k-step first (unconditional, fixed offsets), then the products: the loads of
      // one step overlap the products of the previous one instead of stalling each product on its own operand
      R rv[KC], av[NBW], bv[NBW];
#pragma unroll
      for (int c = 0; c < KC; ++c) rv[c] = rk[(4 * k4 + tq) * KC + c];
#pragma unroll
      for (int i = 0; i < NBW; ++i) { av[i] = xrow[offA[i]] - xbA[i]; bv[i] = xrow[offB[i]] - xbB[i]; }
#pragma unroll
      for (int i = 0; i < NBW; ++i) {
        if (!ok[i]) continue;
#pragma unroll
        for (int c = 0; c < KC; ++c) {
          double d0 = (double)acc[i][c][0], d1 = (double)acc[i][c][1];
          asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(d0), "+d"(d1)
                       : "d"((double)(av[i] * rv[c])), "d"((double)bv[i]));
          acc[i][c][0] = (R)d0;
          acc[i][c][1] = (R)d1;
        }
      }
    }
    if (side_on) {
      for (int r = ssub; r < TS; r += nsub) {
        const R xv = sj < d ? xs[r * ld + sj] - side_xb : R(1);
#pragma unroll
        for (int c = 0; c < KC; ++c) side[c] = fma(rk[r * KC + c], xv, side[c]);
      }
    }
#else
    CTA_FOR(idx, TS * d) {  // centre the tile on x_bar (in place; rows beyond the range stay zero)
      const int r = idx / d, j = idx % d;
      if (n0 + r < row1) xs[r * ld + j] -= xbar[j];
    }
    for (int c = 0; c < KC; ++c)
      for (int r = 0; r < TS; ++r) {
        const R rv = rk[r * KC + c];
        accn[c] += rv;
        for (int i = 0; i < d; ++i) {
          const R wi = rv * xs[r * ld + i];
          accs[(size_t)c * d + i] += wi;
          for (int j = i; j < d; ++j) accS[((size_t)c * d + i) * d + j] = fma(wi, xs[r * ld + j], accS[((size_t)c * d + i) * d + j]);
        }
      }
#endif
    buf ^= 1;  // no barrier here: the one at the top of the next iteration orders this tile's reads before the
               // copies that overwrite its buffer (they are issued after that barrier)
  }
  cp_async_wait_all();

  // partial of (cluster, range): [nk | s (d) | S (d,d) both halves]
#if !defined(DONK_EMU)
  // the row subsets of the first moments meet in the (now idle) tile buffers: [nsub][KC][d + 1], summed in fixed order
  w.cta_sync();
  if (side_on) {
#pragma unroll
    for (int c = 0; c < KC; ++c) Xs[((size_t)ssub * KC + c) * (d + 1) + sj] = side[c];
  }
  w.cta_sync();
#pragma unroll
  for (int c = 0; c < KC; ++c) {
    const int k = kg * KC + c;
    if (k >= K) continue;
    R* out = a.part + ((size_t)k * a.nch + ch) * pl;
    if (w.tid <= d) {
      R tot = R(0);
      for (int sb = 0; sb < nsub; ++sb) tot += Xs[((size_t)sb * KC + c) * (d + 1) + w.tid];
      out[w.tid < d ? 1 + w.tid : 0] = tot;
    }
#pragma unroll
    for (int i = 0; i < NBW; ++i) {
      if (!ok[i]) continue;
      const int gi = 8 * ib[i] + g, gj = 8 * jb[i] + 2 * tq;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (gi <= gj + h && gj + h < d) {
          out[1 + d + (size_t)gi * d + gj + h] = acc[i][c][h];
          out[1 + d + (size_t)(gj + h) * d + gi] = acc[i][c][h];
        }
      }
    }
  }
#else
  for (int c = 0; c < KC; ++c) {
    const int k = kg * KC + c;
    if (k >= K) continue;
    R* out = a.part + ((size_t)k * a.nch + ch) * pl;
    out[0] = accn[c];
    for (int i = 0; i < d; ++i) out[1 + i] = accs[(size_t)c * d + i];
    for (int i = 0; i < d; ++i)
      for (int j = i; j < d; ++j) {
        out[1 + d + (size_t)i * d + j] = accS[((size_t)c * d + i) * d + j];
        out[1 + d + (size_t)j * d + i] = accS[((size_t)c * d + i) * d + j];
      }
  }
#endif
}

// ------------------------------------------------------------------------------------------------------
// M pass for 18 column blocks (d = 137..144: the transitions of BASELINE config 5), device only.
// The generic shape above gives a warp 11 consecutive blocks of the upper triangle: per k-step 24 fragment loads, 22
// centring subtractions and 22 scalings for 22 products - 2 scalar FP64 instructions per DMMA on the pipe the DMMAs
// use (8.9 TFLOP/s at 2 M x 144, K = 16).  Here a warp owns a PAIR of block rows (W, 17 - W): 18 - W + W + 1 = 19
// blocks for every W, with two A fragments (scaled once per cluster) and at most 18 B fragments per k-step; the tile is
// centred on x_bar ONCE in shared memory after it lands (one pass + one barrier per 64-row tile instead of 20
// subtractions per k-step and lane).  Nine pairs on nine warps put three warps on one sub-partition (44 ms per pass at
// 2 M x 144: that sub-partition's DMMA pipe is the critical path); EIGHT warps take pairs 0..7 and share the ninth pair
// (rows 8 and 9, whose B fragments every warp loads anyway) as 2-3 extra blocks each: 21-22 blocks per warp, two warps
// per sub-partition, 255 registers.
// ------------------------------------------------------------------------------------------------------
#if !defined(DONK_EMU)
enum { GMM_M18_WARPS = 8, GMM_M18_KC = 2, GMM_M18_ACC = 22 };
DONK_HD size_t gmm_m18_smem(int d) {
  const int D8 = round_up(d, 8), ld = ld_pad(D8);
  return (size_t)2 * GMM_MTS * ld + (size_t)2 * GMM_MTS * GMM_M18_KC + D8 + 8 + 2;
}

// the ninth pair (block rows 8 and 9) dealt to the eight warps: row 8 blocks [X8, X8 + N8), row 9 blocks [X9, X9 + N9)
template <int W> struct GmmM18Extra {
  static constexpr int X8 = W == 0 ? 8 : W == 1 ? 11 : W == 2 ? 14 : W == 3 ? 17 : 0;
  static constexpr int N8 = W <= 2 ? 3 : W == 3 ? 1 : 0;
  static constexpr int X9 = W == 3 ? 9 : W == 4 ? 10 : W == 5 ? 12 : W == 6 ? 14 : W == 7 ? 16 : 0;
  static constexpr int N9 = W == 3 ? 1 : W >= 4 ? 2 : 0;
};

template <int W>  // one 64-row tile for the warp that owns block rows W and 17 - W (+ its share of rows 8 and 9)
DONK_D void gmm_m18_tile(double (&acc)[GMM_M18_ACC][GMM_M18_KC][2], const double* xs, const double* rk, int ld, int g, int tq) {
  using E = GmmM18Extra<W>;
  constexpr int NA = 18 - W, RB = 17 - W;  // blocks of row W: jb = W .. 17; blocks of row 17 - W: jb = RB .. 17
#pragma unroll 1
  for (int k4 = 0; k4 < GMM_MTS / 4; ++k4) {
    const double* xrow = xs + (4 * k4 + tq) * ld + g;
    const double r0 = rk[(4 * k4 + tq) * GMM_M18_KC], r1 = rk[(4 * k4 + tq) * GMM_M18_KC + 1];
    const double xa = xrow[8 * W], xb = xrow[8 * RB];
    const double a00 = xa * r0, a01 = xa * r1, a10 = xb * r0, a11 = xb * r1;
    double e80 = 0.0, e81 = 0.0, e90 = 0.0, e91 = 0.0;
    if (E::N8 > 0) { const double x8 = xrow[64]; e80 = x8 * r0; e81 = x8 * r1; }
    if (E::N9 > 0) { const double x9 = xrow[72]; e90 = x9 * r0; e91 = x9 * r1; }
#pragma unroll
    for (int q = 0; q < NA; ++q) {
      const double bv = xrow[8 * (W + q)];
      if (E::N8 > 0 && W + q >= E::X8 && W + q < E::X8 + E::N8) {  // block (8, W + q)
        const int qe = 19 + (W + q - E::X8);
        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[qe][0][0]), "+d"(acc[qe][0][1]) : "d"(e80), "d"(bv));
        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[qe][1][0]), "+d"(acc[qe][1][1]) : "d"(e81), "d"(bv));
      }
      if (E::N9 > 0 && W + q >= E::X9 && W + q < E::X9 + E::N9) {  // block (9, W + q)
        const int qe = 19 + E::N8 + (W + q - E::X9);
        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[qe][0][0]), "+d"(acc[qe][0][1]) : "d"(e90), "d"(bv));
        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[qe][1][0]), "+d"(acc[qe][1][1]) : "d"(e91), "d"(bv));
      }
      asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[q][0][0]), "+d"(acc[q][0][1]) : "d"(a00), "d"(bv));
      asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[q][1][0]), "+d"(acc[q][1][1]) : "d"(a01), "d"(bv));
      if (W + q >= RB) {  // the same B fragment serves block (17 - W, W + q)
        const int qb = NA + (W + q - RB);
        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[qb][0][0]), "+d"(acc[qb][0][1]) : "d"(a10), "d"(bv));
        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[qb][1][0]), "+d"(acc[qb][1][1]) : "d"(a11), "d"(bv));
      }
    }
  }
}

template <int W>
DONK_D void gmm_m18_store(const double (&acc)[GMM_M18_ACC][GMM_M18_KC][2], double* out, int c, int d, int g, int tq) {
  using E = GmmM18Extra<W>;
  constexpr int NA = 18 - W, RB = 17 - W;
#pragma unroll
  for (int q = 0; q < 19 + E::N8 + E::N9; ++q) {
    const int ib = q < NA ? W : q < 19 ? RB : q < 19 + E::N8 ? 8 : 9;
    const int jb = q < NA ? W + q : q < 19 ? RB + (q - NA) : q < 19 + E::N8 ? E::X8 + (q - 19) : E::X9 + (q - 19 - E::N8);
    const int gi = 8 * ib + g, gj = 8 * jb + 2 * tq;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (gi <= gj + h && gj + h < d && gi < d) {
        out[1 + d + (size_t)gi * d + gj + h] = acc[q][c][h];
        out[1 + d + (size_t)(gj + h) * d + gi] = acc[q][c][h];
      }
    }
  }
}

DONK_D void gmm_m18_cta(int tid, const GmmArgs<double>& a, int cta, double* sm) {
#if defined(__CUDA_ARCH__)
  constexpr int KC = GMM_M18_KC, TS = GMM_MTS, NT = 32 * GMM_M18_WARPS;
  const int d = a.d, K = a.K;
  const int D8 = round_up(d, 8), ld = ld_pad(D8);
  const int kg = cta / a.nch, ch = cta % a.nch;
  double* Xs = sm;                                // [2][TS][ld]
  double* rr = Xs + (size_t)2 * TS * ld;          // [2][TS][KC]
  double* xbar = rr + (size_t)2 * TS * KC;        // [D8]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(xbar + D8 + 8);
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  const size_t rows_per = ((size_t)a.n + a.nch - 1) / a.nch;
  const size_t row0 = (size_t)ch * rows_per;
  const size_t row1 = row0 + rows_per < (size_t)a.n ? row0 + rows_per : (size_t)a.n;
  const size_t pl = gmm_part_len(d);
  for (int j = tid; j < D8; j += NT) {
    double acc = 0.0;
    if (j < d)
      for (int k = 0; k < K; ++k) acc = fma(__ldg(a.w + k), __ldg(a.means + (size_t)k * d + j), acc);
    xbar[j] = acc;
  }
  for (int idx = tid; idx < 2 * TS * ld; idx += NT) Xs[idx] = 0.0;
  if (tid < 2) mbar_init(bars + tid, 1);
  __syncthreads();
  fence_proxy_async();
  if (cta == 0 && a.shift_out)
    for (int idx = tid; idx < K * d; idx += NT) a.shift_out[idx] = xbar[idx % d];

  double acc[GMM_M18_ACC][KC][2];
#pragma unroll
  for (int q = 0; q < GMM_M18_ACC; ++q)
#pragma unroll
    for (int c = 0; c < KC; ++c) { acc[q][c][0] = 0.0; acc[q][c][1] = 0.0; }
  // first moments / cluster sizes: thread sj <= d walks the rows of every tile (5 of the 8 warps, ~3 % of their FP64 work)
  const int sj = tid;
  const bool side_on = sj <= d;
  double side[KC] = {0.0, 0.0};

  auto stage = [&](size_t n0, int buf) {
    double* xs = Xs + (size_t)buf * TS * ld;
    double* rk = rr + (size_t)buf * TS * KC;
    const size_t left = row1 - n0;
    const int valid = left < (size_t)TS ? (int)left : TS;
    if (tid == 0) mbar_arrive_expect_tx(bars + buf, (unsigned)((size_t)valid * d * sizeof(double)));
    for (int r = tid; r < TS; r += NT) {
      if (r < valid) bulk_copy_g2s(xs + r * ld, a.X + (n0 + r) * d, (unsigned)(d * sizeof(double)), bars + buf);
      else for (int j = 0; j < d; ++j) xs[r * ld + j] = 0.0;
    }
    for (int idx = tid; idx < TS * KC; idx += NT) {
      const int r = idx / KC, c = idx % KC, k = kg * KC + c;
      if (n0 + r < row1 && k < K) cp_async(rk + idx, a.resp + (n0 + r) * K + k);
      else rk[idx] = 0.0;
    }
  };

  int buf = 0;
  unsigned tile = 0;
  if (row0 < row1) stage(row0, 0);
  for (size_t n0 = row0; n0 < row1; n0 += TS, ++tile) {
    cp_async_wait_all();
    mbar_wait(bars + buf, (tile >> 1) & 1);
    fence_proxy_async();  // this thread's centring writes into the other buffer precede the bulk copies about to refill it
    __syncthreads();      // the tile and its responsibilities are visible; every warp has left the other buffer
    if (n0 + TS < row1) stage(n0 + TS, buf ^ 1);
    double* xs = Xs + (size_t)buf * TS * ld;
    const double* rk = rr + (size_t)buf * TS * KC;
    {  // centre the tile on x_bar, once, in place (rows beyond the range carry r = 0)
      const size_t left = row1 - n0;
      const int valid = left < (size_t)TS ? (int)left : TS;
      for (int idx = tid; idx < valid * D8; idx += NT) {
        const int r = idx / D8, j = idx - r * D8;
        if (j < d) xs[r * ld + j] -= xbar[j];
      }
    }
    __syncthreads();
    switch (warp) {
      case 0: gmm_m18_tile<0>(acc, xs, rk, ld, g, tq); break;
      case 1: gmm_m18_tile<1>(acc, xs, rk, ld, g, tq); break;
      case 2: gmm_m18_tile<2>(acc, xs, rk, ld, g, tq); break;
      case 3: gmm_m18_tile<3>(acc, xs, rk, ld, g, tq); break;
      case 4: gmm_m18_tile<4>(acc, xs, rk, ld, g, tq); break;
      case 5: gmm_m18_tile<5>(acc, xs, rk, ld, g, tq); break;
      case 6: gmm_m18_tile<6>(acc, xs, rk, ld, g, tq); break;
      default: gmm_m18_tile<7>(acc, xs, rk, ld, g, tq); break;
    }
    if (side_on) {
      for (int r = 0; r < TS; ++r) {
        const double xv = sj < d ? xs[r * ld + sj] : 1.0;
#pragma unroll
        for (int c = 0; c < KC; ++c) side[c] = fma(rk[r * KC + c], xv, side[c]);
      }
    }
    buf ^= 1;
  }
  cp_async_wait_all();
#pragma unroll
  for (int c = 0; c < KC; ++c) {
    const int k = kg * KC + c;
    if (k >= K) continue;
    double* out = a.part + ((size_t)k * a.nch + ch) * pl;
    if (side_on) out[sj < d ? 1 + sj : 0] = side[c];
    switch (warp) {
      case 0: gmm_m18_store<0>(acc, out, c, d, g, tq); break;
      case 1: gmm_m18_store<1>(acc, out, c, d, g, tq); break;
      case 2: gmm_m18_store<2>(acc, out, c, d, g, tq); break;
      case 3: gmm_m18_store<3>(acc, out, c, d, g, tq); break;
      case 4: gmm_m18_store<4>(acc, out, c, d, g, tq); break;
      case 5: gmm_m18_store<5>(acc, out, c, d, g, tq); break;
      case 6: gmm_m18_store<6>(acc, out, c, d, g, tq); break;
      default: gmm_m18_store<7>(acc, out, c, d, g, tq); break;
    }
  }
#endif
}
#endif  // !DONK_EMU

// ---- launch planning shared by the CUDA library and the emulation ---------------------------------------------
struct GmmMmaPlan {
  int NB, RA, Ge;      // E pass: 8-column blocks, 8-row fragments per warp, warps per CTA
  int NBW, KC, Gm;     // M pass: blocks per warp, clusters per CTA, warps per CTA
  int nch;             // row ranges per cluster group
  int ok;
};
inline GmmMmaPlan gmm_mma_plan(int n, int d, int K) {
  GmmMmaPlan p{};
  const int nb = (d + 7) / 8;
  p.ok = nb <= 18;
  p.NB = nb <= 2 ? 2 : nb <= 5 ? 5 : nb <= 6 ? 6 : nb <= 9 ? 9 : nb <= 12 ? 12 : nb <= 15 ? 15 : 18;
  p.RA = 2;
  p.Ge = 8;
  p.KC = 2;
  p.Gm = nb <= 9 ? 8 : 16;
  p.NBW = nb <= 9 ? 6 : 11;
  // 28..45 blocks (d = 49..72): 16 warps x 3 blocks and FOUR clusters per CTA, one CTA per SM.  The B fragment, the
  // centring and the tile itself are shared by four clusters instead of two (10 shared-memory loads and 6 subtractions
  // per 12 products instead of 14 and 12; every row is read by K/4 CTAs instead of K/2) with the same 48 accumulator
  // registers per lane: 87.2 -> 76.6 ms per EM iteration at 10 M x 72, K = 16.
  const bool wide = nb >= 7 && nb <= 9;
  if (wide) { p.KC = 4; p.Gm = 16; p.NBW = 3; }
  const size_t smem = gmm_m_mma_smem(d, p.KC) * sizeof(double) + 1024;
  int per_sm = (int)((228 * 1024) / smem);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  if (wide) per_sm = 1;
  const int groups = (K + p.KC - 1) / p.KC;
  int nch = (148 * per_sm) / groups;
  const int max_ch = (n + GMM_MTS - 1) / GMM_MTS;
  if (nch > max_ch) nch = max_ch;
  if (nch < 1) nch = 1;
  p.nch = nch;
  return p;
}

// workspace layout in reals: resp (n,K) | lse_part (nblocks) | part (K,nch,1+d+d*d).  Sized from the very plans the
// launchers use: the scalar M pass writes K * gmm_nch partials, the tensor-path M pass K * gmm_mma_plan().nch.
inline int gmm_ws_nch(int n, int d, int K) {
  const int a = gmm_nch(n, d, K), b = gmm_mma_plan(n, d, K).nch;
  return a > b ? a : b;
}
inline size_t gmm_ws_reals(int n, int d, int K) {
  const size_t nblocks = ((size_t)n + GMM_TS - 1) / GMM_TS;
  return (size_t)n * K + nblocks + (size_t)K * gmm_ws_nch(n, d, K) * gmm_part_len(d);
}

}  // namespace donk
