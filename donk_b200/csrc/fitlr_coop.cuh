// donk_b200 — fit_lr, CTA-per-fit variant for large dimensions (BASELINE config 5: dX = 64, dU = 16, d = 144, N = 128).
//
// Same math and reference lines as fitlr.cuh (fit_lr, donk/dynamics/linear_dynamics.py:137-189; NIW posterior
// prior.py:18-45; GMM mixing prior_gmm.py:35-41).  The generic kernel spends ~1.3 M cycles per fit at d = 144 (the
// Householder / Sturm eigenvalue step and column-by-column factorisations with a CTA barrier each dominate); here
//   1. the moments are one pass of DMMA tiles over sample chunks staged with cp.async (shifted by the first sample:
//      no cancellation), accumulated into the blocked joint covariance in shared memory;
//   2. the prior enters through per-fit cluster weights computed beforehand (wts = mean_n predict_proba, the E pass of
//      the EM kernels over all samples of the call) or through explicit NIW parameters; the mixture mean / covariance
//      and the NIW posterior MAP covariance (prior.py:29-45) are formed in place;
//   3. the eigenvalue test of linear_dynamics.py:178-181 is decided by a Cholesky attempt on Sigma_pp - reg I, run
//      in the same panel steps as the factorisation of Sigma_pp itself (two matrices, one set of barriers): it
//      succeeds iff the reference would not regularise.  A fit whose attempt fails is flagged (info = 2) and redone by
//      the generic kernel, which computes the eigenvalue (rare: needs N <= p or collinear data);
//   4. blocked right-looking Cholesky (8-column panels, trailing updates on DMMA), blocked forward substitution
//      Y = L^-1 Sigma_px, Schur complement Sigma_x'x' - Y^T Y on DMMA (one Gram matrix: PSD up to rounding), blocked
//      backward substitution F^T = L^-T Y.
// Dimensions: dX, dU multiples of 8 (whole fragments).
#pragma once
#include "common.cuh"
#include "dmma_tiles.cuh"
#include "fitlr.cuh"

namespace donk {

template <int DX, int DU>
struct FitCoopDims {
  static_assert(DX % 8 == 0 && DU % 8 == 0, "cooperative fit kernel: dX, dU multiples of 8");
  static constexpr int P = DX + DU, D = P + DX;
  static constexpr int NBP = P / 8, NBX = DX / 8, NBD = D / 8;
  static constexpr int PL = ld_pad(P), XL = ld_pad(DX), DL = ld_pad(D);
  static constexpr int NS = 32;  // samples per staged chunk
  static constexpr int SPP = P * PL + 8, SPX = P * XL + 8, SXX = DX * XL + 8;
  static constexpr int STG1 = NS * DL + 8;
  static constexpr int STG = 2 * STG1 > SPP ? 2 * STG1 : SPP;  // two sample chunks; later the matrix of the attempt
  static constexpr int oSpp = 0, oSpx = oSpp + SPP, oSxx = oSpx + SPX, oStg = oSxx + SXX;
  static constexpr int oMu = oStg + STG;        // mean (D)
  static constexpr int oSh = oMu + D + 8;       // shift = first sample (D)
  static constexpr int oSum = oSh + D + 8;      // column sums of the shifted samples (D)
  static constexpr int oRs = oSum + D + 8;      // 1 / L_jj of Sigma_pp (P) | of the attempt (P)
  static constexpr int oMup = oRs + 2 * P + 8;  // prior mean (D)
  static constexpr int oWts = oMup + D + 8;     // cluster weights (<= 64)
  static constexpr int oScal = oWts + 64;       // 0 Np  1 bad factor  2 attempt failed
  static constexpr int REALS = oScal + 8;
  static constexpr int M_RA = NBD % 2 == 0 ? 2 : 1, M_RB = 3;  // moment tiles (whole tiles per block row)
  static constexpr int S_RA = 2, S_RB = 3;      // Schur tiles
  static_assert(NBX % 2 == 0, "cooperative fit kernel: two block rows per Schur tile (dX multiple of 16)");
};

template <typename R, int DX, int DU>
inline size_t fit_coop_bytes() {
  return (size_t)FitCoopDims<DX, DU>::REALS * sizeof(R) + sizeof(CoopSched) + 16;
}

template <typename R, int DX, int DU>
DONK_D void fit_coop_cta(const WCtx& w, const FitArgs<R>& a, int cta, int ncta, R* smem) {
  using D = FitCoopDims<DX, DU>;
  constexpr int P = D::P, d = D::D, PL = D::PL, XL = D::XL, DL = D::DL, NS = D::NS;
  constexpr int NBP = D::NBP, NBX = D::NBX, NBD = D::NBD;
  const int G = w.G, N = a.N, T = a.T, K = a.K;
  const long long nfits = (long long)a.B * T;
  R* Spp = smem + D::oSpp;  // Sigma_pp, full symmetric storage; its lower triangle becomes L
  R* Spx = smem + D::oSpx;  // Sigma_px, then Y = L^-1 Sigma_px, then Z = F^T
  R* Sxx = smem + D::oSxx;
  R* stg = smem + D::oStg;
  R* Ap = stg;              // after the moments: Sigma_pp - reg I (the attempt)
  R* mu = smem + D::oMu;
  R* sh = smem + D::oSh;
  R* sums = smem + D::oSum;
  R* rs = smem + D::oRs;
  R* rsp = rs + P;
  R* mup = smem + D::oMup;
  R* wts = smem + D::oWts;
  R* scal = smem + D::oScal;
  CoopSched& sched = *reinterpret_cast<CoopSched*>(smem + ((D::REALS + 1) / 2) * 2);
  enum { PH_M = 0, PH_S = 1 };

  CTA_FOR(one, 1) {
    coop_schedule(sched, PH_M, G, D::M_RA, D::M_RB, NBD, NBD, true, false);
    coop_schedule(sched, PH_S, G, D::S_RA, D::S_RB, NBX, NBX, true, false);
  }
  w.cta_sync();

  // element (i <= j) of the joint covariance in the blocked storage
  auto sig = [&](int i, int j) -> R* {
    if (j < P) return Spp + i * PL + j;
    if (i < P) return Spx + i * XL + (j - P);
    return Sxx + (i - P) * XL + (j - P);
  };

  for (long long fit = cta; fit < nfits; fit += ncta) {
    const int b = (int)(fit / T), t = (int)(fit % T);
    const size_t bt = (size_t)b * T + t;
    const R* Xb = a.X + ((size_t)b * N * (T + 1) + t) * DX;   // sample n: + n (T+1) DX; x_{t+1} follows x_t
    const R* Ub = a.U + ((size_t)b * N * T + t) * DU;         // sample n: + n T DU
    const size_t xs_n = (size_t)(T + 1) * DX, us_n = (size_t)T * DU;

    // ---- 1. moments -------------------------------------------------------------------------------------------------
    CTA_FOR(idx, D::SPP + D::SPX + D::SXX) smem[D::oSpp + idx] = R(0);
    CTA_FOR(j, d) {
      sh[j] = j < DX ? ldg(Xb + j) : (j < P ? ldg(Ub + (j - DX)) : ldg(Xb + DX + (j - P)));
      sums[j] = R(0);
    }
    CTA_FOR(one, 1) { scal[0] = R(0); scal[1] = R(0); scal[2] = R(0); }
    auto stage = [&](int ch, int buf) {  // rows [x_t | u_t | x_{t+1}] of samples ch NS .. ch NS + NS - 1 (asynchronously)
      R* xs = stg + buf * D::STG1;
      const int n0 = ch * NS;
      if (a.aligned16) {
        constexpr int HX = DX / 2, HU = DU / 2, HR = 2 * HX + HU;
        CTA_FOR(idx, NS * HR) {
          const int r = idx / HR, q = idx % HR, n = n0 + r;
          if (n >= N) continue;
          if (q < HX) cp_async_pair(xs + r * DL + 2 * q, Xb + n * xs_n + 2 * q);
          else if (q < HX + HU) cp_async_pair(xs + r * DL + DX + 2 * (q - HX), Ub + n * us_n + 2 * (q - HX));
          else cp_async_pair(xs + r * DL + P + 2 * (q - HX - HU), Xb + n * xs_n + DX + 2 * (q - HX - HU));
        }
      } else {
        CTA_FOR(idx, NS * d) {
          const int r = idx / d, j = idx % d, n = n0 + r;
          if (n >= N) continue;
          const R* src = j < DX ? Xb + n * xs_n + j : (j < P ? Ub + n * us_n + (j - DX) : Xb + n * xs_n + DX + (j - P));
          cp_async(xs + r * DL + j, src);
        }
      }
    };
    const int nch = (N + NS - 1) / NS;
    w.cta_sync();
    stage(0, 0);
    for (int ch = 0; ch < nch; ++ch) {
      R* xs = stg + (ch & 1) * D::STG1;
      cp_async_wait_all();
      w.cta_sync();  // chunk ch has landed; everyone is done with the tiles of chunk ch - 1
      if (ch + 1 < nch) stage(ch + 1, (ch + 1) & 1);
      CTA_FOR(j, d) {  // shift by the first sample (rows beyond N become zero) and column sums
        const R s0 = sh[j];
        R acc = R(0);
        for (int r = 0; r < NS; ++r) {
          const R v = ch * NS + r < N ? xs[r * DL + j] - s0 : R(0);
          xs[r * DL + j] = v;
          acc += v;
        }
        sums[j] += acc;
      }
      w.cta_sync();
      TEAMS_DO(tm) {
        for (int s = 0; s < sched.n[PH_M][tm]; ++s) {
          const int ra0 = sched.t[PH_M][tm][s][0], rb0 = sched.t[PH_M][tm][s][1];
          tile_mma_n<R, D::M_RA, D::M_RB, 0, false, false, true>(
              w, NBD - rb0, rb0 == ra0, xs + 8 * ra0, DL, xs + 8 * rb0, DL, NS / 4, [](int, int, R*) {},
              [&](int il, int jl, R v0, R v1, const R*) {
                const int i = 8 * ra0 + il, j = 8 * rb0 + jl;
                if (i <= j) *sig(i, j) += v0;
                if (i <= j + 1) *sig(i, j + 1) += v1;
              });
        }
      }
    }
    w.cta_sync();
    // mean and centred, biased covariance (linear_dynamics.py:164-165): S/N - delta delta^T, delta = mean - shift
    CTA_FOR(j, d) mu[j] = sh[j] + sums[j] / R(N);
    w.cta_sync();
    {
      const R invN = R(1) / R(N);
      CTA_FOR(idx, d * d) {
        const int i = idx / d, j = idx % d;
        if (i > j) continue;
        const R di = sums[i] * invN, dj = sums[j] * invN;
        const R v = fma(-di, dj, *sig(i, j) * invN);
        *sig(i, j) = v;
      }
    }
    w.cta_sync();

    // ---- 2. prior -> NIW posterior MAP covariance (prior_gmm.py:35-41, prior.py:18-45) ------------------------------
    if (K > 0 && a.wts_in != nullptr) {
      const size_t dd = (size_t)d * d;
      CTA_FOR(k, K) wts[k] = ldg(a.wts_in + bt * K + k);
      w.cta_sync();
      CTA_FOR(j, d) {
        R acc = R(0);
        for (int k = 0; k < K; ++k) acc = fma(wts[k], ldg(a.gmeans + (size_t)k * d + j), acc);
        mup[j] = acc;
      }
      CTA_FOR(one, 1) {
        R acc = R(0);
        for (int k = 0; k < K; ++k) acc = fma(wts[k], ldg(a.gw + k), acc);
        scal[0] = acc * a.n_pool;
      }
      w.cta_sync();
      const R Np = scal[0], n = R(N) / a.prior_weight, coef = Np * n / (Np + n), den = Np + n;
      CTA_FOR(idx, d * d) {
        const int i = idx / d, j = idx % d;
        if (i > j) continue;
        R phi = R(0);
        for (int k = 0; k < K; ++k) phi = fma(wts[k], ldg(a.gcov + (size_t)k * dd + idx), phi);
        const R dm = (mu[i] - mup[i]) * (mu[j] - mup[j]);
        *sig(i, j) = (Np * phi + n * *sig(i, j) + coef * dm) / den;
      }
      w.cta_sync();
    } else if (K == 0 && a.niw_Phi != nullptr) {
      const R Nm = a.niw_Nmean[bt], Nc = a.niw_Ncovar[bt];
      const R n = R(N) / a.prior_weight, coef = Nm * n / (Nm + n), den = Nc + n + R(d) + R(1);
      const R* mu0 = a.niw_mu0 + bt * d;
      const R* Phi = a.niw_Phi + bt * (size_t)d * d;
      CTA_FOR(idx, d * d) {
        const int i = idx / d, j = idx % d;
        if (i > j) continue;
        const R dm = (mu[i] - ldg(mu0 + i)) * (mu[j] - ldg(mu0 + j));
        *sig(i, j) = (ldg(Phi + idx) + n * *sig(i, j) + coef * dm) / den;
      }
      w.cta_sync();
    }
    // full symmetric storage of Sigma_pp and Sigma_x'x'; the attempt matrix Sigma_pp - reg I (over the dead sample chunks)
    CTA_FOR(idx, P * P) {
      const int i = idx / P, j = idx % P;
      const R v = Spp[(i < j ? i : j) * PL + (i < j ? j : i)];
      if (i > j) Spp[i * PL + j] = v;
      Ap[i * PL + j] = i == j ? v - a.reg : v;
    }
    CTA_FOR(idx, DX * DX) {
      const int i = idx / DX, j = idx % DX;
      if (i > j) Sxx[i * XL + j] = Sxx[j * XL + i];
    }
    w.cta_sync();

    // ---- 3./4. blocked right-looking Cholesky of Sigma_pp (and of the attempt), lower triangles --------------------
    for (int jb = 0; jb < NBP; ++jb) {
      const int j0 = 8 * jb;
      for (int jj = 0; jj < 8; ++jj) {  // unblocked inside the panel: one barrier per column
        const int j = j0 + jj, nk = 7 - jj;
        const R dj = Spp[j * PL + j], djp = Ap[j * PL + j];
        const R inv = R(1) / dj, invp = R(1) / djp;
        if (nk > 0) {
          CTA_FOR(idx, (P - j - 1) * nk) {
            const int i = j + 1 + idx / nk, k = j + 1 + idx % nk;
            if (i < k) continue;
            Spp[i * PL + k] = fma(-Spp[i * PL + j] * inv, Spp[k * PL + j], Spp[i * PL + k]);
            Ap[i * PL + k] = fma(-Ap[i * PL + j] * invp, Ap[k * PL + j], Ap[i * PL + k]);
          }
        }
        CTA_FOR(one, 1) {
          rs[j] = rsqrt_r(dj);
          rsp[j] = rsqrt_r(djp);
          if (!(dj > R(0))) scal[1] = R(1);
          if (!(djp > R(0))) scal[2] = R(1);
        }
        w.cta_sync();
      }
      CTA_FOR(idx, (P - j0) * 8) {  // scale the panel: L[i][j] = a[i][j] / sqrt(d_j)
        const int i = j0 + idx / 8, j = j0 + idx % 8;
        if (i < j) continue;
        Spp[i * PL + j] *= rs[j];
        Ap[i * PL + j] *= rsp[j];
      }
      w.cta_sync();
      // trailing update A[ib][kb] -= L[ib][jb] L[kb][jb]^T for jb < kb <= ib (one fragment per warp, round robin)
      TEAMS_DO(tm) {
        int q = 0;
        for (int kb = jb + 1; kb < NBP; ++kb)
          for (int ib = kb; ib < NBP; ++ib, ++q) {
            if (q % G != tm) continue;
#pragma unroll
            for (int m = 0; m < 2; ++m) {
              R* Am = m == 0 ? Spp : Ap;
              tile_mma<R, 1, 1, 0, true, true, false>(
                  w, Am + 8 * ib * PL + j0, PL, Am + 8 * kb * PL + j0, PL, 2, [](int, int, R*) {},
                  [&](int il, int kl, R v0, R v1, const R*) {
                    R* dst = Am + (8 * ib + il) * PL + 8 * kb + kl;
                    dst[0] -= v0;
                    dst[1] -= v1;
                  });
            }
          }
      }
      w.cta_sync();
    }

    // ---- forward substitution Y = L^-1 Sigma_px, blocked (in place) -----------------------------------------------------
    for (int jb = 0; jb < NBP; ++jb) {
      const int j0 = 8 * jb;
      CTA_FOR(c, DX) {
        R y[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          R acc = Spx[(j0 + jj) * XL + c];
#pragma unroll
          for (int k = 0; k < jj; ++k) acc = fma(-Spp[(j0 + jj) * PL + j0 + k], y[k], acc);
          y[jj] = acc * rs[j0 + jj];
          Spx[(j0 + jj) * XL + c] = y[jj];
        }
      }
      w.cta_sync();
      TEAMS_DO(tm) {  // Sigma_px[ib] -= L[ib][jb] Y[jb] for ib > jb
        int q = 0;
        for (int ib = jb + 1; ib < NBP; ++ib)
          for (int cb = 0; cb < NBX; ++cb, ++q) {
            if (q % G != tm) continue;
            tile_mma<R, 1, 1, 0, true, false, false>(
                w, Spp + 8 * ib * PL + j0, PL, Spx + j0 * XL + 8 * cb, XL, 2, [](int, int, R*) {},
                [&](int il, int cl, R v0, R v1, const R*) {
                  R* dst = Spx + (8 * ib + il) * XL + 8 * cb + cl;
                  dst[0] -= v0;
                  dst[1] -= v1;
                });
          }
      }
      w.cta_sync();
    }

    // ---- Schur complement covar = Sigma_x'x' - Y^T Y (linear_dynamics.py:186,188) ------------------------------------------
    TEAMS_DO(tm) {
      for (int s = 0; s < sched.n[PH_S][tm]; ++s) {
        const int ra0 = sched.t[PH_S][tm][s][0], rb0 = sched.t[PH_S][tm][s][1];
        tile_mma_n<R, D::S_RA, D::S_RB, 0, false, false, true>(
            w, NBX - rb0, rb0 == ra0, Spx + 8 * ra0, XL, Spx + 8 * rb0, XL, P / 4, [](int, int, R*) {},
            [&](int il, int jl, R v0, R v1, const R*) {
              const int i = 8 * ra0 + il, j = 8 * rb0 + jl;
              R* out = a.covar + bt * DX * DX;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (i <= j + h) {
                  const R val = Sxx[i * XL + j + h] - (h ? v1 : v0);
                  out[(size_t)i * DX + j + h] = val;  // symmetric by construction
                  out[(size_t)(j + h) * DX + i] = val;
                }
              }
            });
      }
    }
    w.cta_sync();

    // ---- backward substitution Z = L^-T Y, blocked (in place): F^T ---------------------------------------------------------
    for (int jb = NBP - 1; jb >= 0; --jb) {
      const int j0 = 8 * jb;
      CTA_FOR(c, DX) {
        R z[8];
#pragma unroll
        for (int jj = 7; jj >= 0; --jj) {
          R acc = Spx[(j0 + jj) * XL + c];
#pragma unroll
          for (int k = jj + 1; k < 8; ++k) acc = fma(-Spp[(j0 + k) * PL + j0 + jj], z[k], acc);
          z[jj] = acc * rs[j0 + jj];
          Spx[(j0 + jj) * XL + c] = z[jj];
        }
      }
      w.cta_sync();
      TEAMS_DO(tm) {  // Y[ib] -= L[jb][ib]^T Z[jb] for ib < jb
        int q = 0;
        for (int ib = 0; ib < jb; ++ib)
          for (int cb = 0; cb < NBX; ++cb, ++q) {
            if (q % G != tm) continue;
            tile_mma<R, 1, 1, 0, false, false, false>(
                w, Spp + j0 * PL + 8 * ib, PL, Spx + j0 * XL + 8 * cb, XL, 2, [](int, int, R*) {},
                [&](int il, int cl, R v0, R v1, const R*) {
                  R* dst = Spx + (8 * ib + il) * XL + 8 * cb + cl;
                  dst[0] -= v0;
                  dst[1] -= v1;
                });
          }
      }
      w.cta_sync();
    }

    // ---- outputs: F[i][j] = Z[j][i]; f = mu_x' - F mu_p (linear_dynamics.py:184-185) --------------------------------
    CTA_FOR(idx, DX * P) {
      const int i = idx / P, j = idx % P;
      a.F[bt * DX * P + idx] = Spx[j * XL + i];
    }
    CTA_FOR(i, DX) {
      R acc = R(0);
      for (int j = 0; j < P; ++j) acc = fma(Spx[j * XL + i], mu[j], acc);
      a.f[bt * DX + i] = mu[P + i] - acc;
    }
    CTA_FOR(one, 1) {
      if (a.info) a.info[bt] = scal[2] != R(0) ? 2 : (scal[1] != R(0) ? 1 : 0);
    }
    w.cta_sync();
  }
}

}  // namespace donk
