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

// ---------------------------------------------------------------------------------------------------------------------
// Tensor-path form of the same pass (device only; K <= 16, d even, d + 1 <= 16 NTW x 8).  The scalar form above spends
// ~35 k cycles per 64-row tile on index arithmetic, a 73-thread serial accumulation and shared-memory read-modify-writes
// (6.3 ms per pass at 10 M x 72, K = 16 = 0.92 TB/s of an HBM-bound kernel).  Here
//   - every 8 d-byte row of a tile is one bulk asynchronous copy (cp.async.bulk on an mbarrier, double buffered);
//   - x . c_k for the 8 rows of a warp and all (padded to 16) centres is 2 x d/4 DMMAs against the transposed centres
//     resident in shared memory; |x|^2 rides along; the nearest centre is a 4-lane shuffle reduction (first minimum,
//     as numpy.argmin);
//   - the per-cluster sums are a "one-hot" product  S[k][j] = sum_r [lab_r == k] x[r][j]  on DMMA with the row index as
//     the k dimension (column d of the tile is a column of ones: the counts), accumulators in registers for the whole
//     pass: no shared-memory accumulation, no atomics, fixed order.
// ---------------------------------------------------------------------------------------------------------------------
DONK_HD int kmeans_mma_ld(int d) {  // columns: d features, the ones column, zero padding; ld % 16 in {4, 12}
  const int m = (d + 1 + 7) / 8 * 8;
  return m % 16 == 0 ? m + 4 : (m % 16 == 8 ? m + 4 : m);
}
DONK_HD size_t kmeans_mma_smem(int d) {
  const int ld = kmeans_mma_ld(d), d4 = (d + 3) / 4 * 4;
  return (size_t)2 * KM_TS * ld + (size_t)d4 * 20 + 16 + 2 * KM_TS / 2 + 4;  // tiles | centres^T (ld 20) | |c|^2 | labels | mbarriers
}
DONK_HD bool kmeans_mma_ok(int d, int K, const void* X) {
  return K <= 16 && d % 2 == 0 && d >= 4 && (d + 1 + 7) / 8 <= 24 && (reinterpret_cast<size_t>(X) & 15) == 0;
}

#if defined(__CUDA_ARCH__) || !defined(DONK_EMU)
template <int NTW>  // 8-column blocks of the sums per warp (8 warps: NTW = 1 up to d = 63, 2 up to d = 127, 3 up to d = 191)
DONK_D void kmeans_pass_mma_cta(int tid, int n, int d, int K, const double* X, const double* centers, int* labels,
                                double* mind2, double* dist_out, double* part, int cta, int ncta, double* sm,
                                const double* closest_in = nullptr, double* pot_part = nullptr) {
#if defined(__CUDA_ARCH__)
  const int TS = KM_TS, ld = kmeans_mma_ld(d), d4 = (d + 3) / 4 * 4, KS = d4 / 4, NBD = (d + 1 + 7) / 8;
  double* xs = sm;                                 // [2][TS][ld]
  double* cs = xs + (size_t)2 * TS * ld;           // [d4][20]  centres transposed, clusters padded to 16
  double* cn = cs + (size_t)d4 * 20;               // [16]      |c_k|^2 (padding: huge)
  int* lab = reinterpret_cast<int*>(cn + 16);      // [2][TS]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(lab + 2 * TS);
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  for (int idx = tid; idx < 2 * TS * ld; idx += 256) xs[idx] = (idx % ld) == d ? 1.0 : 0.0;  // ones column, zero padding
  for (int idx = tid; idx < d4 * 20; idx += 256) {
    const int j = idx / 20, k = idx % 20;
    cs[idx] = (k < K && j < d) ? __ldg(centers + (size_t)k * d + j) : 0.0;
  }
  if (tid < 2) mbar_init(bars + tid, 1);
  __syncthreads();
  if (tid < 16) {
    double a = 0.0;
    for (int j = 0; j < d; ++j) a = fma(cs[j * 20 + tid], cs[j * 20 + tid], a);
    cn[tid] = tid < K ? a : 1e300;
  }
  fence_proxy_async();
  __syncthreads();
  const int nblocks = (n + TS - 1) / TS;
  auto stage = [&](int blk, int buf) {  // threads 0..63: one row each
    const size_t n0 = (size_t)blk * TS;
    const int valid = n - n0 < (size_t)TS ? (int)(n - n0) : TS;
    if (tid == 0) mbar_arrive_expect_tx(bars + buf, (unsigned)((size_t)valid * d * sizeof(double)));
    __syncwarp();
    if (tid < TS) {
      double* row = xs + ((size_t)buf * TS + tid) * ld;
      if (tid < valid) bulk_copy_g2s(row, X + (n0 + tid) * d, (unsigned)(d * sizeof(double)), bars + buf);
      else for (int j = 0; j < d; ++j) row[j] = 0.0;
    }
  };
  double S[NTW][2][2];  // sums of the warp's column blocks: [block][cluster half][fragment]
#pragma unroll
  for (int t = 0; t < NTW; ++t)
#pragma unroll
    for (int m = 0; m < 2; ++m) { S[t][m][0] = 0.0; S[t][m][1] = 0.0; }
  // greedy k-means++ round (closest_in given): dist_out receives min(|x - c_k|^2, closest) and pot_part the per-CTA sums
  // of those minima per candidate (the potential each candidate would leave)
  double pot[4] = {0.0, 0.0, 0.0, 0.0};
  unsigned it = 0;
  if (cta < nblocks && tid < 64) stage(cta, 0);
  for (int blk = cta; blk < nblocks; blk += ncta, ++it) {
    const int buf = it & 1;
    const size_t n0 = (size_t)blk * TS;
    mbar_wait(bars + buf, (it >> 1) & 1);
    const double* xt = xs + (size_t)buf * TS * ld;
    int* lb = lab + buf * TS;
    {  // ---- distances of the warp's 8 rows to all centres ------------------------------------------------------
      double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0}, xx = 0.0;
      const double* xr = xt + (8 * warp + g) * ld + tq;
      const double* cb = cs + tq * 20 + g;
#pragma unroll 2
      for (int ks = 0; ks < KS; ++ks) {
        const double a = xr[4 * ks], b0 = cb[ks * 80], b1 = cb[ks * 80 + 8];
        if (4 * ks + tq < d) xx = fma(a, a, xx);  // beyond d: the ones column / padding
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b0));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b1));
      }
      xx += __shfl_xor_sync(0xffffffffu, xx, 1);
      xx += __shfl_xor_sync(0xffffffffu, xx, 2);
      // the lane holds x_r . c_k for r = 8 warp + g and k = 2 tq, 2 tq + 1, 8 + 2 tq, 9 + 2 tq
      double dv[4] = {cn[2 * tq] - 2 * c0[0], cn[2 * tq + 1] - 2 * c0[1], cn[8 + 2 * tq] - 2 * c1[0], cn[9 + 2 * tq] - 2 * c1[1]};
      const int kv[4] = {2 * tq, 2 * tq + 1, 8 + 2 * tq, 9 + 2 * tq};
      double bd = dv[0];
      int bk = kv[0];
#pragma unroll
      for (int q = 1; q < 4; ++q)
        if (dv[q] < bd) { bd = dv[q]; bk = kv[q]; }  // kv ascending: the first minimum wins
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, bd, o);
        const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
        if (od < bd || (od == bd && ok < bk)) { bd = od; bk = ok; }
      }
      const size_t row = n0 + 8 * warp + g;
      const bool live = row < (size_t)n;
      if (tq == 0) {
        lb[8 * warp + g] = live ? bk : -1;
        if (live && labels) labels[row] = bk;
        if (live && mind2) { const double v = xx + bd; mind2[row] = v > 0.0 ? v : 0.0; }
      }
      if (live && dist_out) {
        const double cl = closest_in ? __ldg(closest_in + row) : 1e300;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (kv[q] < K) {
            double v = xx + dv[q];
            v = v > 0.0 ? v : 0.0;
            v = v < cl ? v : cl;
            dist_out[row * K + kv[q]] = v;
            pot[q] += v;
          }
      }
    }
    __syncthreads();  // labels of the tile complete; every warp is done with the other buffer
    if (blk + ncta < nblocks && tid < 64) stage(blk + ncta, buf ^ 1);
    if (part) {  // ---- one-hot product: S[k][j] += [lab_r == k] x[r][j] over the 64 rows of the tile ----------------
#pragma unroll 4
      for (int ks = 0; ks < TS / 4; ++ks) {
        const int r = 4 * ks + tq, l = lb[r];
        const double a0 = l == g ? 1.0 : 0.0, a1 = l == 8 + g ? 1.0 : 0.0;
#pragma unroll
        for (int t = 0; t < NTW; ++t) {
          const int nt = warp + 8 * t;
          if (nt < NBD) {
            const double b = xt[r * ld + 8 * nt + g];
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(S[t][0][0]), "+d"(S[t][0][1]) : "d"(a0), "d"(b));
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(S[t][1][0]), "+d"(S[t][1][1]) : "d"(a1), "d"(b));
          }
        }
      }
    }
  }
  if (pot_part) {  // candidates k = 2 tq, 2 tq + 1, 8 + 2 tq, 9 + 2 tq: sum over the rows (g, then the warps)
    __syncthreads();
    double* red = xs;  // [8 warps][16]
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double v = pot[q];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (g == 0) red[warp * 16 + (q < 2 ? 2 * tq + q : 8 + 2 * tq + (q - 2))] = v;
    }
    __syncthreads();
    if (tid < K) {
      double v = 0.0;
      for (int wq = 0; wq < 8; ++wq) v += red[wq * 16 + tid];
      pot_part[(size_t)cta * K + tid] = v;
    }
  }
  if (part) {  // per-CTA partial: [K*d sums | K counts]; the lane holds S[k = 8 m + g][j = 8 nt + 2 tq + h]
    double* out = part + (size_t)cta * ((size_t)K * d + K);
#pragma unroll
    for (int t = 0; t < NTW; ++t) {
      const int nt = warp + 8 * t;
      if (nt >= NBD) continue;
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int k = 8 * m + g, j = 8 * nt + 2 * tq + h;
          if (k < K && j < d) out[k * d + j] = S[t][m][h];
          else if (k < K && j == d) out[K * d + k] = S[t][m][h];
        }
    }
  }
#endif
}
#endif

// greedy k-means++ round, generic form (small pools, shapes outside the tensor path, emulation): item = row n of the
// distance matrix `dmin` (n,K) filled by a pass with dist output: clamp by closest[n]; the potentials are summed by the
// caller in row order
template <typename R>
DONK_HD void kmeanspp_min_item(int K, const R* closest, R* dmin, size_t row) {
  const R cl = closest[row];
  for (int k = 0; k < K; ++k) dmin[row * K + k] = dmin[row * K + k] < cl ? dmin[row * K + k] : cl;
}

template <typename R>
DONK_HD void kmeans_reduce_item(int ncta, int len, const R* part, R* out, int idx) {
  R acc = R(0);
  for (int c = 0; c < ncta; ++c) acc += part[(size_t)c * len + idx];
  out[idx] = acc;
}

}  // namespace donk
