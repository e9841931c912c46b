// donk_b200 — fit_lr, warp-per-fit variant with compile-time dimensions (fast path, with or without the GMM prior).
//
// Same math and reference lines as fitlr.cuh (fit_lr, donk/dynamics/linear_dynamics.py:137-189).  The generic
// kernel spends a CTA and ~200 __syncthreads on every (condition, t); here one warp owns a fit and needs only
// ~16 KB of shared memory, so ~14 fits are in flight per SM (ncu: the first warp version, 42 KB per fit, ran 4
// warps per SM and was bound by dependent-FMA latency):
//   1. moments in one pass on the FP64 tensor path: every lane loads its DMMA fragment elements of
//      [x_t | u_t | x_{t+1}] straight from global memory (one sample row = k index), shifted by the first sample
//      (keeps the one-pass covariance free of cancellation); S += Xs^T Xs with A and B fragments being the same
//      registers; column sums ride along;
//   2. the eigenvalue test of linear_dynamics.py:178-181 is decided by a Cholesky attempt on Sigma_pp - reg I
//      (succeeds iff lambda_min > reg, i.e. iff the reference would not regularise); only if it fails the smallest
//      eigenvalue is computed (Householder tridiagonalisation + Sturm multisection) and the diagonal lifted.
//      All factorisations work in the lower triangle of Sigma_pp; the upper triangle keeps the original;
//   1b. optional GMM prior (GMMPrior.eval, prior_gmm.py:26-41): log-probabilities of the N samples under the K
//      clusters as Y^T = P_k^T X^T on the tensor path - the fragments of P_k for one block of output features stay
//      in registers while the sample blocks stream past (samples read from global memory in the order
//      [x_t | x_{t+1} | u_t], which is contiguous; rows of the upper-triangular P_k that cannot contribute are
//      skipped at compile time); softmax, mixture moments, NIW posterior MAP covariance (prior.py:29-45);
//   3. Cholesky, forward solve Y = L^-1 Sigma_px (one right-hand side per lane), Schur complement
//      Sigma_x'x' - Y^T Y as DMMA blocks (symmetric and PSD by construction), backward solve F^T = L^-T Y.
#pragma once
#include "common.cuh"
#include "fitlr.cuh"
#if defined(DONK_EMU)
#include <vector>
#endif

namespace donk {

template <int DX, int DU>
struct FitWarpDims {
  static constexpr int P = DX + DU, D = P + DX;
  static constexpr int D8 = (D + 7) / 8 * 8, NBD = D8 / 8;
  static constexpr int P4 = (P + 3) / 4 * 4, PA = ld_pad(P), DXP = ld_pad(DX);
  static constexpr int NBX = (DX + 7) / 8;
  static constexpr int oSpp = 0, oSpx = oSpp + P4 * PA + 8, oSxx = oSpx + P4 * DXP + 8;
  static constexpr int oMu = oSxx + DX * DXP + 8;  // mean (D8)
  static constexpr int oTd = oMu + D8, oTe = oTd + PA + 4, oV = oTe + PA + 4, oWw = oV + PA + 4;
  static constexpr int oScal = oWw + PA + 4, TEAM = oScal + 8;
  // prior: runtime-sized tail behind TEAM: resp (N8 x K) | wts (K8) | mu_prior (D8); no resp when the cluster weights
  // come precomputed (FitArgs::wts_in: one E pass of the EM kernels over all samples of the call, abi_fit_coop.cu)
  static constexpr int NSB_MAX = 8;  // sample blocks of 8 the prior path holds partial sums for (N <= 64)
  static_assert(DX % 2 == 0, "prior path: [x_t | x_{t+1}] is read as one run of 2 dX values in blocks of 4");
  static constexpr int KX4 = 2 * DX / 4, KU4 = (DU + 3) / 4, KT4 = KX4 + KU4;
  DONK_HD static int prior_tail(int N, int K, bool pre = false) {
    return K > 0 ? (pre ? 0 : ((N + 7) / 8 * 8) * (K | 1)) + (K + 7) / 8 * 8 + D8 : 0;
  }
};

template <typename R, int DX, int DU>
DONK_D void fit_warp_cta(const WCtx& w, const FitArgs<R>& a, int cta, R* smem) {
  using D = FitWarpDims<DX, DU>;
  constexpr int P = D::P, d = D::D, P4 = D::P4, PA = D::PA, DXP = D::DXP;
  const int N = a.N, T = a.T;
  const long long nfits = (long long)a.B * T;
  const R eps = sizeof(R) == 8 ? R(2.220446049250313e-16) : R(1.1920929e-7);
  [[maybe_unused]] const bool env_flag_no_prefetch = a.no_prefetch != 0;  // experiments: profiles/exp

  TEAMS_DO(tm) {
    const long long fit = (long long)cta * w.G + tm;
    if (fit >= nfits) continue;
    const int b = (int)(fit / T), t = (int)(fit % T);
    const bool pre_wts = a.wts_in != nullptr;
    const int team_sz = D::TEAM + D::prior_tail(N, a.K, pre_wts);
    R* tp = smem + (size_t)tm * team_sz;
    R* Spp = tp + D::oSpp;  // upper triangle + diagonal: Sigma_pp; strict lower triangle: work space for factors
    R* Spx = tp + D::oSpx;  // Sigma_px, then Y = L^-1 Sigma_px, then Z = F^T
    R* Sxx = tp + D::oSxx;
    R* mu = tp + D::oMu;
    R* td = tp + D::oTd;
    R* te = tp + D::oTe;
    R* vv = tp + D::oV;
    R* ww = tp + D::oWw;
    R* scal = tp + D::oScal;
    LANE_FOR(idx, team_sz) tp[idx] = R(0);
    w.team_sync();

    // ---- 1. S = sum_n (x_n - x_0)(x_n - x_0)^T (upper blocks) and column sums, fragments straight from global ----
    auto store_sigma = [&](int i, int j, R val) {  // element (i <= j) of the joint covariance, blocked storage
      if (j < P) { Spp[i * PA + j] = val; Spp[j * PA + i] = val; }
      else if (i < P) Spx[i * DXP + (j - P)] = val;
      else { Sxx[(i - P) * DXP + (j - P)] = val; Sxx[(j - P) * DXP + (i - P)] = val; }
    };
#if !defined(DONK_EMU)
    {
      constexpr int NBD = D::NBD, NTRI = NBD * (NBD + 1) / 2;
      const int g = w.lane >> 2, tq = w.lane & 3;
      // all sample rows of the fit into L2 at once: the accumulation loop below keeps only two rows per lane in flight,
      // and at DRAM latency that chain was half of the kernel (ncu: 51 % of the stall samples on its loads)
      if (!env_flag_no_prefetch) {
        for (int n = w.lane; n < N; n += 32) {
          const char* xr = reinterpret_cast<const char*>(a.X + (((size_t)b * N + n) * (T + 1) + t) * DX);
          const char* ur = reinterpret_cast<const char*>(a.U + (((size_t)b * N + n) * T + t) * DU);
          const char* xl = reinterpret_cast<const char*>(reinterpret_cast<size_t>(xr) & ~(size_t)127);
#pragma unroll
          for (int o = 0; o < (int)(2 * DX * sizeof(R)) + 128; o += 128)
            if (xl + o < xr + 2 * DX * sizeof(R)) prefetch_l2(xl + o);
          prefetch_l2(ur);
          if (((reinterpret_cast<size_t>(ur) + DU * sizeof(R) - 1) >> 7) != (reinterpret_cast<size_t>(ur) >> 7))
            prefetch_l2(ur + DU * sizeof(R) - 1);
        }
      }
      const R invN = R(1) / R(N);
      // Register budget (3 CTAs x 4 warps per SM = 170 per thread): 84 for the accumulators, 36 for three sample rows in
      // rotation, 12 for the column sums; addresses are a 32-bit element offset per column block from two row bases
      // (64-bit pointers + strides per block spilled: 176 B of stack inside the loop, and a load that is spilled waits
      // for its data), and the shift lives in shared memory (the Sxx area is free until the covariance is stored).
      static_assert(32 * NBD <= DX * D::DXP + 8, "warp fit kernel: the shift does not fit the Sxx scratch");
      R* shs = Sxx + w.lane;  // shs[32 ib]: shift of column 8 ib + g = the first sample
      const R* Xb = a.X + (((size_t)b * N) * (T + 1) + t) * DX;  // sample 0: [x_t | x_{t+1}] (adjacent), then stride sx
      const R* Ub = a.U + (((size_t)b * N) * T + t) * DU;
      const int sx4 = 4 * (T + 1) * DX, su4 = 4 * T * DU;  // the launcher keeps N (T+1) dX below 2^31
      // column 8 ib + g: offset in its row and which base; padding columns (>= d) read column 0 of X - their products land
      // in rows / columns of the accumulators that are never stored
      auto col_is_u = [&](int ib) { const int j = 8 * ib + g; return j >= DX && j < P; };
      auto col_off = [&](int ib) { const int j = 8 * ib + g; return j < DX ? j : (j < P ? j - DX : (j < d ? j - DU : 0)); };
      bool ok[NBD];
      int off[NBD];  // element offset of row 4 s + tq, column 8 ib + g
      R acc[NTRI][2], csum[NBD], cur[NBD], nxt[NBD];
#pragma unroll
      for (int ib = 0; ib < NBD; ++ib) {
        ok[ib] = 8 * ib + g < d;
        const bool iu = col_is_u(ib);
        shs[32 * ib] = ldg((iu ? Ub : Xb) + col_off(ib));
        off[ib] = col_off(ib) + tq * ((iu ? su4 : sx4) / 4);
        csum[ib] = R(0);
      }
#pragma unroll
      for (int q = 0; q < NTRI; ++q) { acc[q][0] = R(0); acc[q][1] = R(0); }
      // Rows 4 s + tq of step s; the offsets advance by four samples per call.  Rows beyond N re-read sample 0 (= the
      // shift: they contribute zeros); the choice is made on the ADDRESS - a select behind the load would make it wait for
      // its data at once.  Three register sets in rotation by unrolling, not by copying: `nxt = nx2` moves waited for
      // the newest load every step and left a single row in flight (ncu: 13 % of all stall samples on that move, 51 % on
      // the loads of the loop).
      auto load = [&](int n, R* dst) {
#pragma unroll
        for (int ib = 0; ib < NBD; ++ib) {
          const bool iu = col_is_u(ib);
          dst[ib] = ldg((iu ? Ub : Xb) + (n < N ? off[ib] : col_off(ib)));
          off[ib] += iu ? su4 : sx4;
        }
      };
      auto accumulate = [&](R* cur) {
#pragma unroll
        for (int ib = 0; ib < NBD; ++ib) { cur[ib] -= shs[32 * ib]; csum[ib] += cur[ib]; }
        int q = 0;
#pragma unroll
        for (int ib = 0; ib < NBD; ++ib)
#pragma unroll
          for (int jb = ib; jb < NBD; ++jb, ++q) {
            double d0 = (double)acc[q][0], d1 = (double)acc[q][1];
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(d0), "+d"(d1)
                         : "d"((double)cur[ib]), "d"((double)cur[jb]));
            acc[q][0] = (R)d0;
            acc[q][1] = (R)d1;
          }
      };
      const int nsteps = (N + 3) / 4;
      R nx2[NBD];
      load(tq, cur);
      load(4 + tq, nxt);
      for (int k4 = 0; k4 < nsteps; k4 += 3) {  // two sample rows are in flight while a third is multiplied
        load(4 * (k4 + 2) + tq, nx2);
        accumulate(cur);
        if (k4 + 1 < nsteps) {
          load(4 * (k4 + 3) + tq, cur);
          accumulate(nxt);
        }
        if (k4 + 2 < nsteps) {
          load(4 * (k4 + 4) + tq, nxt);
          accumulate(nx2);
        }
      }
#pragma unroll
      for (int ib = 0; ib < NBD; ++ib) {
        R s = csum[ib];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        csum[ib] = s * invN;  // every lane of the quad now holds the shifted mean of column 8 ib + g
        if (tq == 0 && ok[ib]) mu[8 * ib + g] = shs[32 * ib] + csum[ib];
      }
      // the shift area becomes Sxx: its padding must read as zeros again (operand fragments over-read into it)
#pragma unroll
      for (int ib = 0; ib < NBD; ++ib) shs[32 * ib] = R(0);
      __syncwarp();
      // Sigma = S / N - m m^T, biased (linear_dynamics.py:165): the lane holds row 8 ib + g, columns 8 jb + 2 tq + h;
      // the shifted mean of column j lives in the quad g' = j % 8 and is fetched by shuffle
      int q = 0;
#pragma unroll
      for (int ib = 0; ib < NBD; ++ib)
#pragma unroll
        for (int jb = ib; jb < NBD; ++jb, ++q) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const R mj = __shfl_sync(0xffffffffu, csum[jb], 4 * (2 * tq + h));
            const int i = 8 * ib + g, j = 8 * jb + 2 * tq + h;
            if (i <= j && j < d) store_sigma(i, j, fma(acc[q][h], invN, -csum[ib] * mj));
          }
        }
    }
#else
    {
      std::vector<R> S((size_t)d * d, R(0)), cs(d, R(0)), x0(d), xr(d);
      for (int j = 0; j < d; ++j) x0[j] = xux_load(a, b, 0, t, j);
      for (int n = 0; n < N; ++n) {
        for (int j = 0; j < d; ++j) xr[j] = xux_load(a, b, n, t, j) - x0[j];
        for (int i = 0; i < d; ++i) {
          cs[i] += xr[i];
          for (int j = i; j < d; ++j) S[(size_t)i * d + j] = fma(xr[i], xr[j], S[(size_t)i * d + j]);
        }
      }
      for (int i = 0; i < d; ++i) { cs[i] /= R(N); mu[i] = x0[i] + cs[i]; }
      for (int i = 0; i < d; ++i)
        for (int j = i; j < d; ++j) store_sigma(i, j, S[(size_t)i * d + j] / R(N) - cs[i] * cs[j]);
    }
#endif
    w.team_sync();

    // ---- 1b. GMM prior -> NIW posterior MAP covariance (prior_gmm.py:26-41, prior.py:18-45) -------------------------
    if (a.K > 0) {
      const int K = a.K, N8 = (N + 7) / 8 * 8, K8 = (K + 7) / 8 * 8;
      const int KP = K | 1;  // odd pitch of the per-sample rows: the softmax runs a lane per sample
      const size_t dd = (size_t)d * d;
      R* resp = tp + D::TEAM;  // [N8][K]: log-probabilities, then responsibilities (absent with precomputed weights)
      R* wts = resp + (pre_wts ? 0 : (size_t)N8 * KP);
      R* mup = wts + K8;
      auto load_sigma = [&](int i, int j) {  // element (i <= j) of the blocked joint covariance
        if (j < P) return Spp[i * PA + j];
        if (i < P) return Spx[i * DXP + (j - P)];
        return Sxx[(i - P) * DXP + (j - P)];
      };
      if (pre_wts) {
        LANE_FOR(k, K) wts[k] = ldg(a.wts_in + (size_t)fit * K + k);
      } else {
#if !defined(DONK_EMU)
      {
        constexpr int NBD = D::NBD, KX4 = D::KX4, KT4 = D::KT4, NSB = D::NSB_MAX;
        const int g = w.lane >> 2, tq = w.lane & 3;
        const R* Xb = a.X + (((size_t)b * N) * (T + 1) + t) * DX;  // sample n: [x_t | x_{t+1}] = Xb[n * sx + 0 .. 2 DX)
        const R* Ub = a.U + (((size_t)b * N) * T + t) * DU;
        const size_t sx = (size_t)(T + 1) * DX, su = (size_t)T * DU;
        for (int k = 0; k < K; ++k) {
          const R* Pk = a.gpc + k * dd;
          R ssp[NSB][2];
#pragma unroll
          for (int rb = 0; rb < NSB; ++rb) { ssp[rb][0] = R(0); ssp[rb][1] = R(0); }
          R cj[NBD];  // -(mu_k P_k) at the lane's output features 8 ra + g
#pragma unroll
          for (int ra = 0; ra < NBD; ++ra) cj[ra] = 8 * ra + g < d ? ldg(a.gconst + (size_t)k * d + 8 * ra + g) : R(0);
          // Two sample blocks at a time against ALL feature blocks: a sample fragment is read once per cluster (it is
          // this fit's own data, an L2 access), the P_k fragments it meets are shared by every fit on the SM (L1).
          // k-order of the product: [x_t | x_{t+1} | u_t]; block k4 holds the original rows o(4 k4 + tq).  P_k is upper
          // triangular: a block whose smallest original row index exceeds 8 ra + 7 contributes nothing to feature block
          // ra and is skipped (compile-time test)
#pragma unroll
          for (int rb = 0; rb < NSB; rb += 2) {
            if (8 * rb < N) {  // warp-uniform
              const int n0 = 8 * rb + g, n1 = n0 + 8;
              const bool v0 = n0 < N, v1 = n1 < N;
              const R* x0p = Xb + (size_t)n0 * sx + tq;
              const R* x1p = Xb + (size_t)n1 * sx + tq;
              const R* u0p = Ub + (size_t)n0 * su + tq;
              const R* u1p = Ub + (size_t)n1 * su + tq;
              double c0[NBD][2], c1[NBD][2];
#pragma unroll
              for (int ra = 0; ra < NBD; ++ra) { c0[ra][0] = c0[ra][1] = c1[ra][0] = c1[ra][1] = 0; }
#pragma unroll
              for (int k4 = 0; k4 < KT4; ++k4) {
                const int kk = 4 * k4 + tq;
                const int orig = kk < DX ? kk : (kk < 2 * DX ? kk + DU : kk - DX);
                const int omin = 4 * k4 < DX ? 4 * k4 : (4 * k4 < 2 * DX ? 4 * k4 + DU : 4 * k4 - DX);
                R b0, b1;
                if (k4 < KX4) {
                  b0 = v0 ? ldg(x0p + 4 * k4) : R(0);
                  b1 = v1 ? ldg(x1p + 4 * k4) : R(0);
                } else {
                  const bool fv = 4 * (k4 - KX4) + tq < DU;
                  b0 = (v0 && fv) ? ldg(u0p + 4 * (k4 - KX4)) : R(0);
                  b1 = (v1 && fv) ? ldg(u1p + 4 * (k4 - KX4)) : R(0);
                }
#pragma unroll
                for (int ra = 0; ra < NBD; ++ra) {
                  if (omin <= 8 * ra + 7) {
                    const int col = 8 * ra + g;
                    const R pa = (orig < d && col < d && orig <= col) ? ldg(Pk + (size_t)orig * d + col) : R(0);
                    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                        : "+d"(c0[ra][0]), "+d"(c0[ra][1]) : "d"((double)pa), "d"((double)b0));
                    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                        : "+d"(c1[ra][0]), "+d"(c1[ra][1]) : "d"((double)pa), "d"((double)b1));
                  }
                }
              }
              // the lane holds y[j = 8 ra + g][n = 8 rb + 2 tq + h] (and the same for block rb + 1)
#pragma unroll
              for (int ra = 0; ra < NBD; ++ra) {
                const R y00 = (R)c0[ra][0] + cj[ra], y01 = (R)c0[ra][1] + cj[ra];
                const R y10 = (R)c1[ra][0] + cj[ra], y11 = (R)c1[ra][1] + cj[ra];
                ssp[rb][0] = fma(y00, y00, ssp[rb][0]);
                ssp[rb][1] = fma(y01, y01, ssp[rb][1]);
                if (rb + 1 < NSB) {
                  ssp[rb + 1][0] = fma(y10, y10, ssp[rb + 1][0]);
                  ssp[rb + 1][1] = fma(y11, y11, ssp[rb + 1][1]);
                }
              }
            }
          }
          const R lc = ldg(a.gconst + (size_t)K * d + k);
#pragma unroll
          for (int rb = 0; rb < NSB; ++rb) {
            if (8 * rb < N) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                R s = ssp[rb][h];
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                s += __shfl_xor_sync(0xffffffffu, s, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 16);
                const int n = 8 * rb + 2 * tq + h;
                if (g == 0 && n < N) resp[(size_t)n * KP + k] = lc - s / 2;
              }
            }
          }
        }
      }
#else
      for (int k = 0; k < K; ++k) {
        const R* Pk = a.gpc + k * dd;
        for (int n = 0; n < N; ++n) {
          R ss = R(0);
          for (int j = 0; j < d; ++j) {
            R y = a.gconst[(size_t)k * d + j];
            for (int i = 0; i <= j; ++i) y = fma(xux_load(a, b, n, t, i), Pk[(size_t)i * d + j], y);
            ss = fma(y, y, ss);
          }
          resp[(size_t)n * KP + k] = a.gconst[(size_t)K * d + k] - ss / 2;
        }
      }
#endif
      w.team_sync();
      LANE_FOR(n, N) {  // responsibilities: softmax over the clusters (sklearn/mixture/_base.py:552-582)
        R m = resp[(size_t)n * KP];
        for (int k = 1; k < K; ++k) m = resp[(size_t)n * KP + k] > m ? resp[(size_t)n * KP + k] : m;
        R s = R(0);
        for (int k = 0; k < K; ++k) s += exp(resp[(size_t)n * KP + k] - m);
        const R lse = m + log(s);
        for (int k = 0; k < K; ++k) resp[(size_t)n * KP + k] = exp(resp[(size_t)n * KP + k] - lse);
      }
      w.team_sync();
      LANE_FOR(k, K) {
        R acc = R(0);
        for (int n = 0; n < N; ++n) acc += resp[(size_t)n * KP + k];
        wts[k] = acc / R(N);  // prior_gmm.py:33
      }
      }  // !pre_wts
      w.team_sync();
      // Mixture moments: sum_k wts_k means_k and sum_k wts_k covariances_k.  The cluster parameters live in global
      // memory (L2); the loads of 4 clusters x 4 entries are issued together before any of them is used - a
      // cluster-by-cluster accumulation waits out one L2 latency per term (ncu: 130 k of the 210 k cycles of a fit).
      LANE_FOR(j, d) {
        R acc = R(0);
        for (int k0 = 0; k0 < K; k0 += 4) {
          R v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] = k0 + q < K ? ldg(a.gmeans + (size_t)(k0 + q) * d + j) : R(0);
#pragma unroll
          for (int q = 0; q < 4; ++q) acc = fma(k0 + q < K ? wts[k0 + q] : R(0), v[q], acc);
        }
        mup[j] = acc;  // prior_gmm.py:35
      }
      LANE_FOR(one, 1) {
        R acc = R(0);
        for (int k = 0; k < K; ++k) acc = fma(wts[k], ldg(a.gw + k), acc);
        scal[4] = acc * a.n_pool;  // prior_gmm.py:39: strength of the prior in samples
      }
      w.team_sync();
      {
        const R Np = scal[4];
        const R nn = R(N) / a.prior_weight;  // linear_dynamics.py:172
        const R coef = Np * nn / (Np + nn);
        const int total = d * d;
        LANE_FOR(base, (total + 3) / 4) {  // entries base, base + Q, base + 2Q, base + 3Q with Q = ceil(total / 4)
          const int Q = (total + 3) / 4;
          int ii[4], jj[4];
          bool on[4];
          R phi[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int idx = base + u * Q;
            ii[u] = idx / d; jj[u] = idx % d;
            on[u] = idx < total && ii[u] <= jj[u];
            phi[u] = R(0);
          }
          for (int k0 = 0; k0 < K; k0 += 4) {
            R v[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
              for (int q = 0; q < 4; ++q)
                v[u][q] = (on[u] && k0 + q < K) ? ldg(a.gcov + (k0 + q) * dd + (size_t)(base + u * Q)) : R(0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const R wq = k0 + q < K ? wts[k0 + q] : R(0);
#pragma unroll
              for (int u = 0; u < 4; ++u) phi[u] = fma(wq, v[u][q], phi[u]);  // prior_gmm.py:36
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (!on[u]) continue;
            const int i = ii[u], j = jj[u];
            const R dm = (mu[i] - mup[i]) * (mu[j] - mup[j]);
            // Phi' / (N_covar' + d + 1), Phi' = Np Phi_p + n S + Np n / (Np + n) dm dm^T  (prior.py:29-45)
            store_sigma(i, j, (Np * phi[u] + nn * load_sigma(i, j) + coef * dm) / (Np + nn));
          }
        }
      }
      w.team_sync();
    }

    // In-place Cholesky of (A - shift I): A is read from the upper triangle + diagonal of Spp, the factor goes to
    // the strict lower triangle, its diagonal to dg.  *flag (shared) becomes 1 if a pivot is not positive.
#if !defined(DONK_EMU)
    // Device form: lane i owns row i of the factor and keeps it in registers (compile-time indices after unrolling);
    // column j needs row j, which lane j has already published to shared memory - one broadcast load per term instead
    // of two dependent loads, no loop overhead.  Every lane recomputes the pivot (no cross-lane dependency in a column).
    auto chol = [&](R shift, R* dg, R* flag) {
      R Lr[P];
      bool bad = false;
      const int li = w.lane;
#pragma unroll
      for (int j = 0; j < P; ++j) {
        R sj = Spp[j * PA + j] - shift;
        R si = (li > j && li < P) ? Spp[j * PA + li] : R(0);  // A[i][j] from the upper triangle
#pragma unroll
        for (int k = 0; k < j; ++k) {
          const R ljk = Spp[j * PA + k];
          sj = fma(-ljk, ljk, sj);
          si = fma(-Lr[k], ljk, si);
        }
        bad = bad || !(sj > R(0));
        const R rs = rsqrt_r(sj);
        Lr[j] = si * rs;
        if (li == j) dg[j] = rs;  // the device form stores the RECIPROCAL pivots: the substitutions multiply
        if (li > j && li < P) Spp[li * PA + j] = Lr[j];
        w.team_sync();
      }
      if (li == 0) *flag = bad ? R(1) : R(0);
      w.team_sync();
    };
#else
    auto chol = [&](R shift, R* dg, R* flag) {
      LANE_FOR(one, 1) *flag = R(0);
      w.team_sync();
      for (int j = 0; j < P; ++j) {
        LANE_FOR(r, P - j) {
          const int i = j + r;
          R sj = Spp[j * PA + j] - shift, si = Spp[j * PA + i];  // A[i][j] from the upper triangle
          for (int k = 0; k < j; ++k) {
            const R ljk = Spp[j * PA + k];
            sj = fma(-ljk, ljk, sj);
            si = fma(-Spp[i * PA + k], ljk, si);
          }
          if (i == j) {
            dg[j] = sqrt(sj);
            if (!(sj > R(0))) *flag = R(1);
          } else {
            Spp[i * PA + j] = si * rsqrt_r(sj);
          }
        }
        w.team_sync();
      }
    };
#endif

    // ---- 2. does the reference regularise?  lambda_min(Sigma_pp) < reg  <=>  Sigma_pp - reg I is not PD ------------
    chol(a.reg, ww, scal + 0);
    if (scal[0] != R(0)) {  // warp-uniform slow path: smallest eigenvalue of Sigma_pp (linear_dynamics.py:178)
      // Householder tridiagonalisation in the strict lower triangle; td is the working (finally the tridiagonal)
      // diagonal, te the sub-diagonal, vv the Householder vector, ww = tau A v
      LANE_FOR(idx, P * P) {
        const int i = idx / P, j = idx % P;
        if (j < i) Spp[i * PA + j] = Spp[j * PA + i];
      }
      LANE_FOR(i, P) td[i] = Spp[i * PA + i];
      w.team_sync();
      for (int k = 0; k + 2 < P; ++k) {
        LANE_FOR(one, 1) { scal[1] = R(0); scal[2] = R(0); }
        w.team_sync();
        team_reduce_add(w, scal + 1, P - k - 2, [&](int r) { const R x = Spp[(k + 2 + r) * PA + k]; return x * x; });
        w.team_sync();
        const R xnorm2 = scal[1], alpha = Spp[(k + 1) * PA + k];
        R tau = R(0), beta = alpha, scale = R(0);
        if (xnorm2 > R(0)) {
          const R nrm = sqrt(alpha * alpha + xnorm2);
          beta = alpha >= R(0) ? -nrm : nrm;
          tau = (beta - alpha) / beta;
          scale = R(1) / (alpha - beta);
        }
        LANE_FOR(i, P - k - 1) vv[k + 1 + i] = i == 0 ? R(1) : Spp[(k + 1 + i) * PA + k] * scale;
        LANE_FOR(one, 1) te[k] = beta;
        w.team_sync();
        if (tau != R(0)) {
          team_reduce_add(w, scal + 2, P - k - 1, [&](int i) {
            const int ri = k + 1 + i;
            R acc = R(0);
            for (int j = k + 1; j < P; ++j) {
              const R aij = j == ri ? td[ri] : (j < ri ? Spp[ri * PA + j] : Spp[j * PA + ri]);
              acc = fma(aij, vv[j], acc);
            }
            ww[ri] = tau * acc;
            return tau * acc * vv[ri];
          });
          w.team_sync();
          const R al2 = -tau * scal[2] / 2;
          LANE_FOR(i, P - k - 1) {
            const int ri = k + 1 + i;
            const R vi = vv[ri], wi = fma(al2, vi, ww[ri]);
            for (int j = k + 1; j < ri; ++j) Spp[ri * PA + j] -= vi * fma(al2, vv[j], ww[j]) + wi * vv[j];
            td[ri] -= R(2) * vi * wi;
          }
          w.team_sync();
        }
      }
      LANE_FOR(one, 1) {
        if (P >= 2) te[P - 2] = Spp[(P - 1) * PA + P - 2];
      }
      w.team_sync();
      R lo = td[0], hi = td[0], emax = R(0);
      for (int i = 0; i < P; ++i) {
        const R el = i > 0 ? fabs(te[i - 1]) : R(0), er = i + 1 < P ? fabs(te[i]) : R(0);
        const R gsh = td[i] - el - er;
        lo = gsh < lo ? gsh : lo;
        hi = td[i] < hi ? td[i] : hi;
        emax = el > emax ? el : emax;
      }
      const R tnorm = fabs(lo) > fabs(hi) ? fabs(lo) : fabs(hi);
      const R pivmin = (sizeof(R) == 8 ? R(2.3e-308) : R(1.2e-38)) * (emax * emax > R(1) ? emax * emax : R(1));
      lo -= R(2) * eps * tnorm * R(P) + R(2) * pivmin;
      hi += R(2) * eps * tnorm * R(P) + R(2) * pivmin;
      R* cnt = vv;  // eigenvalue counts of the 32 section points
      for (int round = 0; round < 16; ++round) {
        const R width = hi - lo;
        LANE_FOR(q, 32) {
          const R x = lo + width * R(q + 1) / R(33);
          int c = 0;
          R piv = td[0] - x;
          if (fabs(piv) < pivmin) piv = -pivmin;
          c += piv < R(0);
          for (int i = 1; i < P; ++i) {
            piv = td[i] - x - te[i - 1] * te[i - 1] / piv;
            if (fabs(piv) < pivmin) piv = -pivmin;
            c += piv < R(0);
          }
          cnt[q] = R(c);
        }
        w.team_sync();
        int first = 32;
        for (int q = 31; q >= 0; --q)
          if (cnt[q] > R(0)) first = q;
        const R nlo = first == 0 ? lo : lo + width * R(first) / R(33);
        const R nhi = first == 32 ? hi : lo + width * R(first + 1) / R(33);
        lo = nlo; hi = nhi;
        w.team_sync();
        if (hi - lo <= eps * tnorm) break;
      }
      const R min_eig = (lo + hi) / 2;
      if (min_eig < a.reg) {
        LANE_FOR(i, P) Spp[i * PA + i] += a.reg - min_eig;
      }
      w.team_sync();
    }

    // ---- 3. Cholesky of Sigma_pp; Y = L^-1 Sigma_px; covar = Sigma_x'x' - Y^T Y; Z = L^-T Y = F^T ----------------------
    chol(R(0), td, scal + 3);
#if !defined(DONK_EMU)
    if (w.lane < DX) {  // forward substitution, one right-hand side per lane, the solution in registers
      const int c = w.lane;
      R y[P];
#pragma unroll
      for (int i = 0; i < P; ++i) {
        R a0 = Spx[i * DXP + c], a1 = R(0);
#pragma unroll
        for (int k = 0; k < i; k += 2) {
          a0 = fma(-Spp[i * PA + k], y[k], a0);
          if (k + 1 < i) a1 = fma(-Spp[i * PA + k + 1], y[k + 1], a1);
        }
        y[i] = (a0 + a1) * td[i];
        Spx[i * DXP + c] = y[i];
      }
    }
#else
    LANE_FOR(c, DX) {  // forward substitution, one right-hand side per lane
      for (int i = 0; i < P; ++i) {
        R acc = Spx[i * DXP + c];
        for (int k = 0; k < i; ++k) acc = fma(-Spp[i * PA + k], Spx[k * DXP + c], acc);
        Spx[i * DXP + c] = acc / td[i];
      }
    }
#endif
    w.team_sync();
    const size_t bt = (size_t)b * T + t;
    {
      // Sigma_xp Sigma_pp^-1 Sigma_px = Y^T Y: symmetric and PSD by construction (linear_dynamics.py:186-188)
      auto epi = [&](int i, int j, R v0, R v1) {
        if (i < DX && i <= j + 1) {
          if (i <= j && j < DX) {
            const R val = Sxx[i * DXP + j] - v0;
            a.covar[bt * DX * DX + (size_t)i * DX + j] = val;
            a.covar[bt * DX * DX + (size_t)j * DX + i] = val;
          }
          if (j + 1 < DX) {
            const R val = Sxx[i * DXP + j + 1] - v1;
            a.covar[bt * DX * DX + (size_t)i * DX + j + 1] = val;
            a.covar[bt * DX * DX + (size_t)(j + 1) * DX + i] = val;
          }
        }
      };
#define DONK_FIT_ROW(RB_ROW)                                                                              \
  if (RB_ROW < D::NBX)                                                                                    \
    warp_mma<R, 1, (D::NBX - RB_ROW > 0 ? D::NBX - RB_ROW : 1)>(                                           \
        w, Spx + 8 * RB_ROW, DXP, Spx + 8 * RB_ROW, DXP, P4 / 4,                                           \
        [&](int i, int j, R v0, R v1) { epi(8 * RB_ROW + i, 8 * RB_ROW + j, v0, v1); });
      DONK_FIT_ROW(0) DONK_FIT_ROW(1) DONK_FIT_ROW(2) DONK_FIT_ROW(3)
#undef DONK_FIT_ROW
    }
    w.team_sync();
#if !defined(DONK_EMU)
    if (w.lane < DX) {  // backward substitution
      const int c = w.lane;
      R z[P];
#pragma unroll
      for (int i = P - 1; i >= 0; --i) {
        R a0 = Spx[i * DXP + c], a1 = R(0);
#pragma unroll
        for (int k = i + 1; k < P; k += 2) {
          a0 = fma(-Spp[k * PA + i], z[k], a0);
          if (k + 1 < P) a1 = fma(-Spp[(k + 1) * PA + i], z[k + 1], a1);
        }
        z[i] = (a0 + a1) * td[i];
        Spx[i * DXP + c] = z[i];
      }
    }
#else
    LANE_FOR(c, DX) {  // backward substitution
      for (int i = P - 1; i >= 0; --i) {
        R acc = Spx[i * DXP + c];
        for (int k = i + 1; k < P; ++k) acc = fma(-Spp[k * PA + i], Spx[k * DXP + c], acc);
        Spx[i * DXP + c] = acc / td[i];
      }
    }
#endif
    w.team_sync();
    // outputs F[i][j] = Z[j][i], f = mu_x' - F mu_p   (linear_dynamics.py:184-185)
    LANE_FOR(idx, DX * P) {
      const int i = idx / P, j = idx % P;
      a.F[bt * DX * P + idx] = Spx[j * DXP + i];
    }
    LANE_FOR(i, DX) {
      R acc = R(0);
      for (int j = 0; j < P; ++j) acc = fma(Spx[j * DXP + i], mu[j], acc);
      a.f[bt * DX + i] = mu[P + i] - acc;
    }
    LANE_FOR(one, 1) {
      if (a.info) a.info[bt] = scal[3] != R(0);
    }
    w.team_sync();
  }
}

template <typename R, int DX, int DU>
inline int fit_warp_plan(size_t smem_cap, int force_G, int N, int K, int& G, size_t& bytes, bool pre_wts = false) {
  using D = FitWarpDims<DX, DU>;
  const size_t per = (size_t)(D::TEAM + D::prior_tail(N, K, pre_wts)) * sizeof(R);
  G = force_G > 0 ? force_G : 4;
  while (G > 1 && (size_t)G * per > smem_cap) --G;
  bytes = (size_t)G * per;
  return bytes <= smem_cap ? DONK_OK : DONK_SMEM_LIMIT;
}

}  // namespace donk
