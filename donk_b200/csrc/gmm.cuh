// donk_b200 — full-covariance Gaussian-mixture EM over the (sharded) transition pool.
//
// Replaces sklearn.mixture.GaussianMixture.fit as called by GMMPrior.update
// (donk/dynamics/prior/prior_gmm.py:14-24); the iteration follows scikit-learn 1.9.0
// (sklearn/mixture/_base.py:263-278, _gaussian_mixture.py:193-196,312-313,364-370,490-553):
//   E: y = x P_k - mu_k P_k, logp_k = -(d log 2pi + |y|^2)/2 + sum log diag P_k + log pi_k,
//      lse = logsumexp_k, resp = exp(logp - lse), lower bound = mean lse
//   M: nk = sum r + 10 eps, mu = sum r x / nk, Sigma = sum r (x-mu)(x-mu)^T / nk + reg I,
//      pi = nk / n (renormalised), P_k = chol(Sigma_k)^-T
//
// Three kernels per iteration and rank:
//   gmm_e_cta      responsibilities of a block of TS rows (4-row x 4-column register tiles against
//                  the upper-triangular P_k, read through L1/L2) + per-block sum of lse;
//   gmm_m_cta      for one cluster and one contiguous range of rows: nk, sum r (x - mu_k) and the
//                  d x d scatter sum r (x-mu_k)(x-mu_k)^T in REGISTER tiles that persist over the
//                  whole range (statistics centred on the current mean: no cancellation);
//   gmm_reduce     fixed-order sum of the per-range partials into the packed statistics that the
//                  host all-reduces across ranks (NCCL), then gmm_finalize_cta turns them into the
//                  new parameters (K tiny Cholesky factorisations, replicated on every rank).
#pragma once
#include "common.cuh"
#if defined(DONK_EMU)
#include <vector>
#endif

namespace donk {

enum { GMM_TS = 64, GMM_MAXT = 2 };  // rows per staged block; register tiles per thread in gmm_m_cta

// packed statistics: [ nk_raw (K) | s (K,d) | S (K,d,d) | sum lse (1) ]
DONK_HD size_t gmm_stats_len(int d, int K) { return (size_t)K * (1 + d + (size_t)d * d) + 1; }
DONK_HD size_t gmm_part_len(int d) { return 1 + d + (size_t)d * d; }  // one (cluster, range) partial

template <typename R>
struct GmmArgs {
  int n, d, K;
  const R* X;                        // (n,d)
  const R *w, *means, *pc;           // (K) (K,d) (K,d,d) upper-triangular precision Cholesky
  R* resp;                           // (n,K)
  R* lse_part;                       // (nblocks) or null
  int nblocks;
  // M pass
  int nch;                           // row ranges per cluster
  R* part;                           // (K,nch,1+d+d*d)
  R* shift_out;                      // (K,d) centring point of the statistics (written by the M pass), or null
  // E pass only, optional: the rows are transitions read straight from roll-outs instead of a materialised pool
  // (fit_lr's prior weights, prior_gmm.py:26-33): row (c N + n) T + t = [x_t | u_t | x_{t+1}] of tX (.., T+1, dX) and
  // tU (.., T, dU) (donk/dynamics/utils.py:4-17); X is then null and d = 2 dX + dU
  const R *tX, *tU;
  int tT, tdX, tdU;
  // tensor-path E pass, optional: the precision factors in fragment order (gmm_pfrag_item): the B fragment of block
  // (fb, jb), k-step h is one contiguous 256-byte run
  const R* pf;
};

// Fragment-major copy of the upper-triangular P_k for the tensor-path E pass: for cluster k, block row fb (features
// 8 fb .. 8 fb + 7 = the k dimension of the product), block column jb >= fb and k-step h, lane (g = lane / 4, tq = lane % 4)
// holds P_k[8 fb + 4 h + tq][8 jb + g] (zero outside the d x d matrix) at ((k NT + tile(fb, jb)) 2 + h) 32 + lane.
DONK_HD int gmm_pfrag_tile(int NB, int fb, int jb) { return fb * NB - fb * (fb - 1) / 2 + (jb - fb); }
DONK_HD size_t gmm_pfrag_len(int NB, int K) { return (size_t)K * (NB * (NB + 1) / 2) * 64; }
template <typename R>
DONK_HD void gmm_pfrag_item(int d, int K, int NB, const R* pc, R* pf, size_t idx) {
  const int lane = (int)(idx % 32), h = (int)((idx / 32) % 2);
  const size_t kt = idx / 64;
  const int NT = NB * (NB + 1) / 2, t = (int)(kt % NT), k = (int)(kt / NT);
  int fb = 0, rem = t;
  while (rem >= NB - fb) { rem -= NB - fb; ++fb; }
  const int jb = fb + rem;
  const int f = 8 * fb + 4 * h + (lane & 3), j = 8 * jb + (lane >> 2);
  pf[idx] = (f < d && j < d && k < K) ? pc[((size_t)k * d + f) * d + j] : R(0);
}

// element j of row `row` of the E pass input (pool row or transition view)
template <typename R>
DONK_D R gmm_row_item(const GmmArgs<R>& a, size_t row, int j) {
  if (a.tX == nullptr) return ldg(a.X + row * a.d + j);
  const size_t s = (unsigned)row / (unsigned)a.tT;  // sample (c N + n); the launcher keeps the row count below 2^31
  const int t = (int)((unsigned)row - (unsigned)s * (unsigned)a.tT);
  if (j < a.tdX) return ldg(a.tX + (s * (a.tT + 1) + t) * a.tdX + j);
  if (j < a.tdX + a.tdU) return ldg(a.tU + (s * a.tT + t) * a.tdU + (j - a.tdX));
  return ldg(a.tX + (s * (a.tT + 1) + t + 1) * a.tdX + (j - a.tdX - a.tdU));
}

// ---------------------------------------------------------------------------------------------
// E pass: one CTA per block of GMM_TS rows (grid-stride).
// shared: xs (TS x dp) | sP (K x dp) | logp (TS x K) | ldet (K)
// ---------------------------------------------------------------------------------------------
DONK_HD size_t gmm_e_smem(int d, int K) {
  const int dp = round_up(d, 4);
  return (size_t)GMM_TS * dp + (size_t)K * dp + (size_t)GMM_TS * K + round_up(K, 4);
}

template <typename R>
DONK_D void gmm_e_cta(const Ctx& ctx, const GmmArgs<R>& a, int cta, int ncta, R* sm) {
  const int d = a.d, K = a.K, dp = round_up(d, 4), TS = GMM_TS;
  R* xs = sm;
  R* sP = xs + (size_t)TS * dp;
  R* logp = sP + (size_t)K * dp;
  R* ldet = logp + (size_t)TS * K;
  const size_t dd = (size_t)d * d;
  // sP[k][j] = -(mu_k P_k)[j];  ldet[k] = sum log diag P_k + log pi_k
  TEAM_FOR(idx, K * d) {
    const int k = idx / d, j = idx % d;
    const R* P = a.pc + (size_t)k * dd;
    R acc = R(0);
    for (int i = 0; i <= j; ++i) acc = fma(ldg(a.means + (size_t)k * d + i), ldg(P + (size_t)i * d + j), acc);
    sP[k * dp + j] = -acc;
  }
  TEAM_FOR(k, K) {
    const R* P = a.pc + (size_t)k * dd;
    R acc = R(0);
    for (int i = 0; i < d; ++i) acc += log(ldg(P + (size_t)i * d + i));
    ldet[k] = acc + log(ldg(a.w + k));
  }
  ctx.sync();
  const int ntj = (d + 3) / 4, nrt = TS / 4;
  const R cst = R(d) * R(1.8378770664093453);  // d log(2 pi)
  for (int blk = cta; blk < a.nblocks; blk += ncta) {
    const size_t n0 = (size_t)blk * TS;
    TEAM_FOR(idx, TS * dp) {
      const int r = idx / dp, j = idx % dp;
      xs[idx] = (n0 + r < (size_t)a.n && j < d) ? gmm_row_item(a, n0 + r, j) : R(0);
    }
    ctx.sync();
    TEAM_FOR(idx, K * nrt) {
      const int k = idx / nrt, rt = idx % nrt;
      const R* P = a.pc + (size_t)k * dd;
      R ss[4] = {R(0), R(0), R(0), R(0)};
      for (int jt = 0; jt < ntj; ++jt) {
        const int j0 = 4 * jt;
        R acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = R(0);
        const int kmax = j0 + 4 < d ? j0 + 4 : d;
        for (int i = 0; i < kmax; ++i) {
          R av[4], bv[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) av[r] = xs[(4 * rt + r) * dp + i];
#pragma unroll
          for (int c = 0; c < 4; ++c) bv[c] = (j0 + c < d && i <= j0 + c) ? ldg(P + (size_t)i * d + j0 + c) : R(0);
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = fma(av[r], bv[c], acc[r][c]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (j0 + c < d) {
            const R s = sP[k * dp + j0 + c];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const R y = acc[r][c] + s;
              ss[r] = fma(y, y, ss[r]);
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) logp[(4 * rt + r) * K + k] = -(cst + ss[r]) / 2 + ldet[k];
    }
    ctx.sync();
    TEAM_FOR(r, TS) {
      if (n0 + r >= (size_t)a.n) { logp[r * K] = R(0); continue; }
      R* lp = logp + r * K;
      R m = lp[0];
      for (int k = 1; k < K; ++k) m = lp[k] > m ? lp[k] : m;
      R s = R(0);
      for (int k = 0; k < K; ++k) s += exp(lp[k] - m);
      const R lse = m + log(s);
      for (int k = 0; k < K; ++k) a.resp[(n0 + r) * K + k] = exp(lp[k] - lse);
      lp[0] = lse;
    }
    ctx.sync();
    if (a.lse_part && ctx.tid == 0) {
      R acc = R(0);
      for (int r = 0; r < TS; ++r) acc += logp[r * K];
      a.lse_part[blk] = acc;
    }
    ctx.sync();
  }
}

// ---------------------------------------------------------------------------------------------
// M pass: CTA (k, ch) accumulates the statistics of cluster k over rows [ch*rows_per, ...).
// shared: xc (TS x dp) centred rows | wx (TS x dp) resp-weighted centred rows | rk (TS) | mu (dp)
// ---------------------------------------------------------------------------------------------
DONK_HD size_t gmm_m_smem(int d) {
  const int dp = round_up(d, 4);
  return (size_t)2 * GMM_TS * dp + GMM_TS + dp;
}

template <typename R>
DONK_D void gmm_m_cta(const Ctx& ctx, const GmmArgs<R>& a, int cta, R* sm) {
  const int d = a.d, K = a.K, dp = round_up(d, 4), TS = GMM_TS;
  const int k = cta / a.nch, ch = cta % a.nch;
  R* xc = sm;
  R* wx = xc + (size_t)TS * dp;
  R* rk = wx + (size_t)TS * dp;
  R* mu = rk + TS;
  const int nt = dp / 4, ntri = nt * (nt + 1) / 2;
  const size_t rows_per = ((size_t)a.n + a.nch - 1) / a.nch;
  const size_t row0 = (size_t)ch * rows_per;
  const size_t row1 = row0 + rows_per < (size_t)a.n ? row0 + rows_per : (size_t)a.n;
  R* out = a.part + ((size_t)k * a.nch + ch) * gmm_part_len(d);

  TEAM_FOR(j, dp) mu[j] = j < d ? ldg(a.means + (size_t)k * d + j) : R(0);
  // Register tiles that live across the whole row range.  Under emulation one host thread plays
  // every CTA thread, so its "registers" are a heap array with one slot per tile.
#if defined(DONK_EMU)
  const int my_tiles = ntri;
  std::vector<R> accs_v((size_t)ntri * 16, R(0));
  R* accs = accs_v.data();
  std::vector<R> sums_v(dp + 1, R(0));
  R* sums = sums_v.data();
#else
  R accs[GMM_MAXT * 16];
#pragma unroll
  for (int i = 0; i < GMM_MAXT * 16; ++i) accs[i] = R(0);
  const int my_tiles = GMM_MAXT;
  R sum_j = R(0);   // this thread's column of sum r (x - mu), thread j < d; thread d keeps nk
#endif
  ctx.sync();

  for (size_t n0 = row0; n0 < row1; n0 += TS) {
    TEAM_FOR(r, TS) rk[r] = n0 + r < row1 ? a.resp[(n0 + r) * K + k] : R(0);
    ctx.sync();
    TEAM_FOR(idx, TS * dp) {
      const int r = idx / dp, j = idx % dp;
      const R v = (n0 + r < row1 && j < d) ? ldg(a.X + (n0 + r) * d + j) - mu[j] : R(0);
      xc[idx] = v;
      wx[idx] = rk[r] * v;
    }
    ctx.sync();
#if defined(DONK_EMU)
    for (int tl = 0; tl < ntri; ++tl) {
      R* acc = accs + (size_t)tl * 16;
#else
#pragma unroll
    for (int q = 0; q < GMM_MAXT; ++q) {
      const int tl = ctx.tid + q * ctx.nt;
      if (tl >= ntri) break;
      R* acc = accs + q * 16;
#endif
      int ti, tj;
      tri_tile(tl, nt, ti, tj);
      const R* pa_ = wx + 4 * ti;
      const R* pb_ = xc + 4 * tj;
      for (int r = 0; r < TS; ++r) {
        R av[4], bv[4];
        ld4(pa_, av);
        ld4(pb_, bv);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i * 4 + j] = fma(av[i], bv[j], acc[i * 4 + j]);
        pa_ += dp;
        pb_ += dp;
      }
    }
#if defined(DONK_EMU)
    for (int j = 0; j < d; ++j)
      for (int r = 0; r < TS; ++r) sums[j] += wx[r * dp + j];
    for (int r = 0; r < TS; ++r) sums[dp] += rk[r];
#else
    if (ctx.tid < d) {
      for (int r = 0; r < TS; ++r) sum_j += wx[r * dp + ctx.tid];
    } else if (ctx.tid == d) {
      for (int r = 0; r < TS; ++r) sum_j += rk[r];
    }
#endif
    ctx.sync();
  }
  (void)my_tiles;
  // write this range's partial: [nk | s (d) | S (d,d) both halves]
#if defined(DONK_EMU)
  out[0] = sums[dp];
  for (int j = 0; j < d; ++j) out[1 + j] = sums[j];
  for (int tl = 0; tl < ntri; ++tl) {
    const R* acc = accs + (size_t)tl * 16;
#else
  if (ctx.tid < d) out[1 + ctx.tid] = sum_j;
  else if (ctx.tid == d) out[0] = sum_j;
#pragma unroll
  for (int q = 0; q < GMM_MAXT; ++q) {
    const int tl = ctx.tid + q * ctx.nt;
    if (tl >= ntri) break;
    const R* acc = accs + q * 16;
#endif
    int ti, tj;
    tri_tile(tl, nt, ti, tj);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gi = 4 * ti + i, gj = 4 * tj + j;
        if (gi <= gj && gj < d) {
          out[1 + d + (size_t)gi * d + gj] = acc[i * 4 + j];
          out[1 + d + (size_t)gj * d + gi] = acc[i * 4 + j];
        }
      }
  }
}

// Fixed-order reduction of the per-range partials and the per-block lse sums into `stats`.
template <typename R>
DONK_HD void gmm_reduce_item(const GmmArgs<R>& a, size_t idx, R* stats) {
  const int d = a.d, K = a.K;
  const size_t pl = gmm_part_len(d), total = gmm_stats_len(d, K);
  if (idx + 1 == total) {
    R acc = R(0);
    for (int bI = 0; bI < a.nblocks; ++bI) acc += a.lse_part[bI];
    stats[idx] = acc;
    return;
  }
  // map packed index -> (k, offset inside a partial)
  size_t k, off;
  if (idx < (size_t)K) { k = idx; off = 0; }
  else if (idx < (size_t)K * (1 + d)) { const size_t r = idx - K; k = r / d; off = 1 + r % d; }
  else { const size_t r = idx - (size_t)K * (1 + d); k = r / ((size_t)d * d); off = 1 + d + r % ((size_t)d * d); }
  R acc = R(0);
  for (int ch = 0; ch < a.nch; ++ch) acc += a.part[(k * a.nch + ch) * pl + off];
  stats[idx] = acc;
}

// ---------------------------------------------------------------------------------------------
// CTA-parallel Cholesky helpers on a shared-memory matrix (ld = lda).
// ---------------------------------------------------------------------------------------------
// A (d x d, symmetric, read only) -> strict lower part of Lm and diagonal dg. Returns 0 / 1 (not PD)
// through *flag (shared).  Column phases: d syncs.
template <typename R>
DONK_D void chol_cta(const Ctx& ctx, int d, const R* A, int lda, R* Lm, int ldl, R* dg, R* flag) {
  if (ctx.tid == 0) *flag = R(0);
  ctx.sync();
  for (int j = 0; j < d; ++j) {
    TEAM_FOR(r, d - j) {
      const int i = j + r;
      R sj = A[j * lda + j], si = A[i * lda + j];
      for (int k = 0; k < j; ++k) {
        const R ljk = Lm[j * ldl + k];
        sj = fma(-ljk, ljk, sj);
        si = fma(-Lm[i * ldl + k], ljk, si);
      }
      const R dj = sqrt(sj);
      if (i == j) {
        dg[j] = dj;
        if (!(sj > R(0))) *flag = R(1);
      } else {
        Lm[i * ldl + j] = si / dj;
      }
    }
    ctx.sync();
  }
}

// M-step finalisation for cluster k = cta.  shared: C (d x dp) | Lm (d x dp) | dg (dp) | dl (dp) | scal (4).
// Above d = 116 two matrices do not fit 227 KB: the factor then overwrites the strict lower triangle of C in place and the
// substitution for L^-1 keeps its vector in the output array (gmm_fin_inplace; d up to 160).
DONK_HD bool gmm_fin_inplace(int d) {
  const int dp = round_up(d, 4);
  return ((size_t)2 * d * dp + 2 * dp + 4) * 8 > (size_t)220 * 1024;
}
DONK_HD size_t gmm_fin_smem(int d) {
  const int dp = round_up(d, 4);
  return (size_t)(gmm_fin_inplace(d) ? 1 : 2) * d * dp + 2 * dp + 4;
}

template <typename R>
struct GmmFinArgs {
  int d, K;
  R n_total, reg_covar;
  const R* stats;
  const R* shift;  // (K,d) centring point the statistics were accumulated around (null: the current means)
  R *w, *means, *cov, *pc, *lower_bound;
  int* info;
};

template <typename R>
DONK_D void gmm_finalize_cta(const Ctx& ctx, const GmmFinArgs<R>& a, int k, R* sm) {
  const int d = a.d, K = a.K, dp = round_up(d, 4);
  const bool inplace = gmm_fin_inplace(d);
  R* C = sm;
  R* Lm = inplace ? C : C + (size_t)d * dp;
  R* dg = Lm + (size_t)d * dp;
  R* dl = dg + dp;
  R* scal = dl + dp;
  const R eps10 = R(10) * (sizeof(R) == 8 ? R(2.220446049250313e-16) : R(1.1920929e-7));
  const R nk_raw = a.stats[k];
  const R nk = nk_raw + eps10;  // _gaussian_mixture.py:312
  const R* s = a.stats + K + (size_t)k * d;
  const R* S = a.stats + (size_t)K * (1 + d) + (size_t)k * d * d;
  // delta = new mean - centre;  new mean = (sum r x) / nk with sum r x = s + nk_raw centre
  TEAM_FOR(j, d) {
    const R mo = a.shift ? a.shift[(size_t)k * d + j] : a.means[(size_t)k * d + j];
    const R mn = (s[j] + nk_raw * mo) / nk;
    dl[j] = mn - mo;
    dg[j] = mn;
  }
  ctx.sync();
  TEAM_FOR(j, d) a.means[(size_t)k * d + j] = dg[j];
  // Sigma = sum r (x - mu_new)(x - mu_new)^T / nk + reg I, expanded around the centring point
  TEAM_FOR(idx, d * d) {
    const int i = idx / d, j = idx % d;
    if (i > j) continue;
    R v = (S[idx] - s[i] * dl[j] - dl[i] * s[j] + nk_raw * dl[i] * dl[j]) / nk;
    if (i == j) v += a.reg_covar;
    C[i * dp + j] = v;
    C[j * dp + i] = v;
    a.cov[(size_t)k * d * d + idx] = v;
    a.cov[(size_t)k * d * d + (size_t)j * d + i] = v;
  }
  if (ctx.tid == 0) {
    // weights = nk, then normalised to sum 1 (sklearn 1.9 _m_step, _gaussian_mixture.py:883-895)
    R tot = R(0);
    for (int q = 0; q < K; ++q) tot += a.stats[q] + eps10;
    a.w[k] = nk / tot;
    if (k == 0 && a.lower_bound) *a.lower_bound = a.stats[gmm_stats_len(d, K) - 1] / a.n_total;
  }
  ctx.sync();
  chol_cta(ctx, d, C, dp, Lm, dp, dg, scal);
  if (ctx.tid == 0 && a.info) a.info[k] = scal[0] != R(0);
  // P = (L^-1)^T: column c of L^-1 by forward substitution, written as row c of P
  TEAM_FOR(c, d) {
    // scratch for x: column c of C (stride dp), or - when the factor lives in C - row c of the output itself
    R* col = inplace ? a.pc + (size_t)k * d * d + (size_t)c * d : C + c;
    const int cs = inplace ? 1 : dp;
    for (int i = 0; i < d; ++i) {
      if (i < c) { col[i * cs] = R(0); continue; }
      R acc = i == c ? R(1) : R(0);
      for (int q = c; q < i; ++q) acc = fma(-Lm[i * dp + q], col[q * cs], acc);
      col[i * cs] = acc / dg[i];
    }
    if (!inplace)
      for (int i = 0; i < d; ++i) a.pc[(size_t)k * d * d + (size_t)c * d + i] = col[i * dp];
  }
}

// Upper-triangular P with P P^T = precision (sklearn _compute_precision_cholesky_from_precisions:
// flip, Cholesky, flip).  One CTA per cluster.  shared: C (d x dp) | Lm (d x dp) | dg (dp) | scal (4)
template <typename R>
DONK_D void gmm_prec_chol_cta(const Ctx& ctx, int d, const R* prec, R* pc, int* info, int k, R* sm) {
  const int dp = round_up(d, 4);
  R* C = sm;
  R* Lm = gmm_fin_inplace(d) ? C : C + (size_t)d * dp;
  R* dg = Lm + (size_t)d * dp;
  R* scal = dg + 2 * dp;
  const R* A = prec + (size_t)k * d * d;
  TEAM_FOR(idx, d * d) {
    const int i = idx / d, j = idx % d;
    C[i * dp + j] = A[(size_t)(d - 1 - i) * d + (d - 1 - j)];
  }
  ctx.sync();
  chol_cta(ctx, d, C, dp, Lm, dp, dg, scal);
  if (ctx.tid == 0 && info) info[k] = scal[0] != R(0);
  TEAM_FOR(idx, d * d) {
    const int i = idx / d, j = idx % d;  // P[i][j] = L[d-1-i][d-1-j]
    const int li = d - 1 - i, lj = d - 1 - j;
    pc[(size_t)k * d * d + idx] = li == lj ? dg[li] : (li > lj ? Lm[li * dp + lj] : R(0));
  }
}

// Row ranges per cluster for the M pass: fill the 148 SMs of a B200 once (K * nch ~ resident CTAs).
inline int gmm_nch(int n, int d, int K) {
  const size_t smem = gmm_m_smem(d) * sizeof(double) + 1024;
  int per_sm = (int)((228 * 1024) / smem);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int nch = (148 * per_sm) / (K > 0 ? K : 1);
  const int max_ch = (n + GMM_TS - 1) / GMM_TS;
  if (nch > max_ch) nch = max_ch;
  if (nch < 1) nch = 1;
  return nch;
}
inline int gmm_m_threads(int d) {
  const int dp = round_up(d, 4), nt = dp / 4, ntri = nt * (nt + 1) / 2;
  for (int per = 1; per <= GMM_MAXT; ++per) {
    int th = round_up((ntri + per - 1) / per, 32);
    if (th < round_up(d + 1, 32)) th = round_up(d + 1, 32);
    if (th <= 384) return th < 128 ? 128 : th;
  }
  return 0;  // d too large for the register-tile M pass
}
// gmm_ws_reals (workspace size) lives in gmm_mma.cuh: it needs the plan of the tensor-path M pass as well.

}  // namespace donk
