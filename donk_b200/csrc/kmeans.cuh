// donk_b200 — k-means (Lloyd) pass used to initialise GMMPrior.update on the device.
//
// The reference's first GMMPrior.update (donk/dynamics/prior/prior_gmm.py:21) starts sklearn's EM from k-means
// labels (sklearn/mixture/_base.py:119-128).  For pools that live on the GPU (BASELINE config 4: 10 M rows) a host
// k-means would dwarf the EM itself; this kernel runs one Lloyd iteration per pass over the pool.  Parity with
// sklearn's KMeans is statistical only (different seeding stream), see DESIGN.md section 6.
//   kmeans_pass_cta     tile of rows staged in shared memory, centres resident (transposed, 4 centres per LDS.128):
//                       squared distances on a 1 row x 4 centres register tile, nearest centre, and - in the same
//                       pass, from the staged tile - the per-cluster sums of the rows (thread j owns column j of
//                       the CTA's accumulators: no atomics, fixed order)
//   kmeans_reduce_item  fixed-order reduction of the per-CTA partial sums
// HBM-bound: X is read once per Lloyd iteration (8 d bytes per row).
#pragma once
#include "common.cuh"

namespace donk {

enum { KM_TS = 64 };

DONK_HD int kmeans_kp(int K) { return (K + 3) & ~3; }
DONK_HD int kmeans_ld(int d) { return d | 1; }  // odd leading dimension: rows of a warp hit distinct banks
DONK_HD size_t kmeans_pass_smem(int d, int K) {
  const int Kp = kmeans_kp(K);
  return (size_t)d * Kp + Kp + (size_t)KM_TS * Kp + (size_t)K * d + K + KM_TS + KM_TS + (size_t)KM_TS * kmeans_ld(d);
}

// labels (n) int32, mind2 (n), dist (n,K), part (per-CTA [K*d sums | K counts]): each optional
template <typename R>
DONK_D void kmeans_pass_cta(const Ctx& ctx, int n, int d, int K, const R* X, const R* centers, int* labels, R* mind2,
                            R* dist_out, R* part, int cta, int ncta, R* sm) {
  const int TS = KM_TS, ld = kmeans_ld(d), Kp = kmeans_kp(K), KQ = Kp / 4;
  R* cs = sm;                         // [d][Kp]  transposed centres (16-byte aligned rows)
  R* cn = cs + (size_t)d * Kp;        // [Kp]     |c_k|^2  (padding: huge)
  R* dist = cn + Kp;                  // [TS][Kp] |c_k|^2 - 2 x.c_k
  R* acc = dist + (size_t)TS * Kp;    // [K][d] sums | [K] counts
  R* xn = acc + (size_t)K * d + K;    // [TS] |x_r|^2
  int* lab = reinterpret_cast<int*>(xn + TS);  // [TS]
  R* xs = xn + 2 * TS;                // [TS][ld]
  TEAM_FOR(idx, d * Kp) {
    const int j = idx / Kp, k = idx % Kp;
    cs[idx] = k < K ? ldg(centers + (size_t)k * d + j) : R(0);
  }
  TEAM_FOR(idx, K * d + K) acc[idx] = R(0);
  ctx.sync();
  TEAM_FOR(k, Kp) {
    R a = R(0);
    for (int j = 0; j < d; ++j) a = fma(cs[j * Kp + k], cs[j * Kp + k], a);
    cn[k] = k < K ? a : R(1e300);
  }
  const int nblocks = (n + TS - 1) / TS;
  for (int blk = cta; blk < nblocks; blk += ncta) {
    const size_t n0 = (size_t)blk * TS;
    ctx.sync();
    TEAM_FOR(idx, TS * d) {
      const int r = idx / d, j = idx % d;
      xs[r * ld + j] = n0 + r < (size_t)n ? ldg(X + (n0 + r) * d + j) : R(0);
    }
    ctx.sync();
    TEAM_FOR(idx, TS * KQ) {  // 1 row x 4 centres per thread; the threads of a row share its x (broadcast)
      const int r = idx / KQ, kq = idx % KQ;
      const R* xr = xs + r * ld;
      const R* ck = cs + 4 * kq;
      R a0 = R(0), a1 = R(0), a2 = R(0), a3 = R(0);
      for (int j = 0; j < d; ++j) {
        const R x = xr[j];
        R c[4];
        ld4(ck + j * Kp, c);
        a0 = fma(x, c[0], a0); a1 = fma(x, c[1], a1); a2 = fma(x, c[2], a2); a3 = fma(x, c[3], a3);
      }
      R* o = dist + r * Kp + 4 * kq;
      o[0] = cn[4 * kq] - 2 * a0; o[1] = cn[4 * kq + 1] - 2 * a1;
      o[2] = cn[4 * kq + 2] - 2 * a2; o[3] = cn[4 * kq + 3] - 2 * a3;
    }
    if (mind2 || dist_out) {
      TEAM_FOR(r, TS) {
        R a = R(0);
        for (int j = 0; j < d; ++j) a = fma(xs[r * ld + j], xs[r * ld + j], a);
        xn[r] = a;
      }
    }
    ctx.sync();
    TEAM_FOR(r, TS) {
      if (n0 + r >= (size_t)n) { lab[r] = -1; continue; }
      int best = 0;
      R bd = dist[r * Kp];
      for (int k = 1; k < K; ++k)
        if (dist[r * Kp + k] < bd) { bd = dist[r * Kp + k]; best = k; }
      lab[r] = best;
      if (labels) labels[n0 + r] = best;
      if (mind2) { const R v = xn[r] + bd; mind2[n0 + r] = v > R(0) ? v : R(0); }
    }
    if (dist_out) {
      TEAM_FOR(idx, TS * K) {
        const int r = idx / K, k = idx % K;
        if (n0 + r < (size_t)n) { const R v = xn[r] + dist[r * Kp + k]; dist_out[(n0 + r) * K + k] = v > R(0) ? v : R(0); }
      }
    }
    if (part) {
      ctx.sync();
      TEAM_FOR(j, d + 1) {
        for (int r = 0; r < TS; ++r) {
          const int k = lab[r];
          if (k < 0) break;
          if (j < d) acc[k * d + j] += xs[r * ld + j];
          else acc[K * d + k] += R(1);
        }
      }
    }
  }
  if (part) {
    ctx.sync();
    TEAM_FOR(idx, K * d + K) part[(size_t)cta * (K * d + K) + idx] = acc[idx];
  }
}

template <typename R>
DONK_HD void kmeans_reduce_item(int ncta, int len, const R* part, R* out, int idx) {
  R acc = R(0);
  for (int c = 0; c < ncta; ++c) acc += part[(size_t)c * len + idx];
  out[idx] = acc;
}

}  // namespace donk
