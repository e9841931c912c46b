// donk_b200 — k-means (Lloyd) passes used to initialise GMMPrior.update on the device.
//
// The reference's first GMMPrior.update (donk/dynamics/prior/prior_gmm.py:21) starts sklearn's EM from k-means
// labels (sklearn/mixture/_base.py:119-128).  For pools that live on the GPU (BASELINE config 4: 10 M rows) a host
// k-means would dwarf the EM itself; these two kernels run Lloyd iterations on the device.  Parity with
// sklearn's KMeans is statistical only (different seeding stream), see DESIGN.md section 6.
//   kmeans_assign_cta   tile of rows staged in shared memory, centres resident: nearest centre and squared distance
//   kmeans_update_cta   per-CTA partial sums of the rows of each cluster (thread j owns column j: no atomics)
//   kmeans_reduce_item  fixed-order reduction of the partials
#pragma once
#include "common.cuh"

namespace donk {

enum { KM_TS = 64 };

DONK_HD size_t kmeans_assign_smem(int d, int K) { return (size_t)KM_TS * (d + 1) + (size_t)K * d + K + KM_TS + (size_t)KM_TS * K; }

template <typename R>
DONK_D void kmeans_assign_cta(const Ctx& ctx, int n, int d, int K, const R* X, const R* centers, int* labels, R* mind2,
                              int cta, int ncta, R* sm) {
  const int TS = KM_TS, ld = d + 1;
  R* xs = sm;                      // [TS][ld]
  R* cs = xs + (size_t)TS * ld;    // [K][d]
  R* cn = cs + (size_t)K * d;      // [K]  |c_k|^2
  R* xn = cn + K;                  // [TS] |x_r|^2
  R* dist = xn + TS;               // [TS][K]
  TEAM_FOR(idx, K * d) cs[idx] = ldg(centers + idx);
  ctx.sync();
  TEAM_FOR(k, K) {
    R acc = R(0);
    for (int j = 0; j < d; ++j) acc = fma(cs[k * d + j], cs[k * d + j], acc);
    cn[k] = acc;
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
    TEAM_FOR(idx, TS * (K + 1)) {
      const int r = idx / (K + 1), k = idx % (K + 1);
      R acc = R(0);
      if (k < K) {
        for (int j = 0; j < d; ++j) acc = fma(xs[r * ld + j], cs[k * d + j], acc);
        dist[r * K + k] = cn[k] - 2 * acc;
      } else {
        for (int j = 0; j < d; ++j) acc = fma(xs[r * ld + j], xs[r * ld + j], acc);
        xn[r] = acc;
      }
    }
    ctx.sync();
    TEAM_FOR(r, TS) {
      if (n0 + r >= (size_t)n) continue;
      int best = 0;
      R bd = dist[r * K];
      for (int k = 1; k < K; ++k)
        if (dist[r * K + k] < bd) { bd = dist[r * K + k]; best = k; }
      labels[n0 + r] = best;
      if (mind2) {
        const R v = xn[r] + bd;
        mind2[n0 + r] = v > R(0) ? v : R(0);
      }
    }
  }
}

DONK_HD size_t kmeans_update_smem(int d, int K) { return (size_t)K * d + K + KM_TS; }

// partial (K*d sums | K counts) of the rows [row0, row1) of this CTA
template <typename R>
DONK_D void kmeans_update_cta(const Ctx& ctx, int n, int d, int K, const R* X, const int* labels, R* part, int cta,
                              int ncta, R* sm) {
  R* acc = sm;                     // [K][d]
  R* cnt = acc + (size_t)K * d;    // [K]
  int* lab = reinterpret_cast<int*>(cnt + K);  // [TS]
  TEAM_FOR(idx, K * d + K) sm[idx] = R(0);
  const size_t rows_per = ((size_t)n + ncta - 1) / ncta;
  const size_t row0 = (size_t)cta * rows_per;
  const size_t row1 = row0 + rows_per < (size_t)n ? row0 + rows_per : (size_t)n;
  for (size_t n0 = row0; n0 < row1; n0 += KM_TS) {
    ctx.sync();
    TEAM_FOR(r, KM_TS) lab[r] = n0 + r < row1 ? labels[n0 + r] : -1;
    ctx.sync();
    TEAM_FOR(j, d + 1) {
      for (int r = 0; r < KM_TS; ++r) {
        const int k = lab[r];
        if (k < 0) break;
        if (j < d) acc[k * d + j] += ldg(X + (n0 + r) * d + j);
        else cnt[k] += R(1);
      }
    }
  }
  ctx.sync();
  TEAM_FOR(idx, K * d + K) part[(size_t)cta * (K * d + K) + idx] = sm[idx];
}

template <typename R>
DONK_HD void kmeans_reduce_item(int ncta, int len, const R* part, R* out, int idx) {
  R acc = R(0);
  for (int c = 0; c < ncta; ++c) acc += part[(size_t)c * len + idx];
  out[idx] = acc;
}

}  // namespace donk
