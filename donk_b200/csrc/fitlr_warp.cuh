// donk_b200 — fit_lr, warp-per-fit variant with compile-time dimensions (fast path, no prior).
//
// Same math and reference lines as fitlr.cuh (fit_lr, donk/dynamics/linear_dynamics.py:137-189).  The generic
// kernel spends a CTA and ~200 __syncthreads on every (condition, t); here one warp owns a fit:
//   1. moments in one pass on the FP64 tensor path: samples are staged in chunks (cp.async, double buffered),
//      shifted by the first sample (keeps the one-pass covariance free of cancellation) and multiplied as DMMA
//      fragments S += Xs^T Xs with the sample index as k dimension; A and B fragments are the same registers;
//   2. the eigenvalue test of linear_dynamics.py:178-181 is decided by a Cholesky attempt on Sigma_pp - reg I:
//      it succeeds iff lambda_min > reg, i.e. iff the reference would not regularise.  Only if it fails the
//      smallest eigenvalue is computed (Householder tridiagonalisation + Sturm multisection, warp level) and
//      the diagonal lifted by reg - lambda_min;
//   3. Cholesky of Sigma_pp, triangular solves for F^T (one right-hand side per lane), Sigma_pp F^T and the Schur
//      complement as DMMA block products.
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
  static constexpr int D8 = (D + 7) / 8 * 8, NBD = D8 / 8, LDX = ld_pad(D8);
  static constexpr int P4 = (P + 3) / 4 * 4, PA = ld_pad(P), DXP = ld_pad(DX);
  static constexpr int NBP = (P + 7) / 8, NBX = (DX + 7) / 8;
  static constexpr int NS = 16;  // samples per staged chunk (4 k4-steps)
  static constexpr int oSpp = 0, oSpx = oSpp + P4 * PA + 8, oSxx = oSpx + P4 * DXP + 8;
  static constexpr int oW1 = oSxx + DX * DXP + 8, oW2 = oW1 + P4 * PA + 8;
  static constexpr int CH = 2 * NS * LDX > 2 * (P4 * DXP + 8) ? 2 * NS * LDX : 2 * (P4 * DXP + 8);
  static constexpr int oCh = oW2 + P4 * PA + 8;       // sample chunks (2 buffers); later Z | Gt
  static constexpr int oMu = oCh + CH + 8;            // shifted mean (D8) | shift x0 (D8)
  static constexpr int oTd = oMu + 2 * D8, oTe = oTd + PA, oV = oTe + PA, oWw = oV + PA;
  static constexpr int oScal = oWw + PA, TEAM = oScal + 8;
};

template <typename R, int DX, int DU>
DONK_D void fit_warp_cta(const WCtx& w, const FitArgs<R>& a, int cta, R* smem) {
  using D = FitWarpDims<DX, DU>;
  constexpr int P = D::P, d = D::D, D8 = D::D8, NBD = D::NBD, LDX = D::LDX, P4 = D::P4, PA = D::PA, DXP = D::DXP;
  constexpr int NS = D::NS;
  const int N = a.N, T = a.T;
  const long long nfits = (long long)a.B * T;
  const R eps = sizeof(R) == 8 ? R(2.220446049250313e-16) : R(1.1920929e-7);

  TEAMS_DO(tm) {
    const long long fit = (long long)cta * w.G + tm;
    if (fit >= nfits) continue;
    const int b = (int)(fit / T), t = (int)(fit % T);
    R* tp = smem + (size_t)tm * D::TEAM;
    R* Spp = tp + D::oSpp;
    R* Spx = tp + D::oSpx;
    R* Sxx = tp + D::oSxx;
    R* W1 = tp + D::oW1;
    R* W2 = tp + D::oW2;
    R* mu = tp + D::oMu;
    R* x0 = mu + D8;
    LANE_FOR(idx, D::TEAM) tp[idx] = R(0);
    w.team_sync();
    LANE_FOR(j, d) x0[j] = xux_load(a, b, 0, t, j);  // shift: the first sample of this (condition, t)
    auto stage = [&](int ch, int buf) {
      R* xs = tp + D::oCh + buf * NS * LDX;
      LANE_FOR(idx, NS * d) {
        const int r = idx / d, j = idx % d, n = ch * NS + r;
        if (n < N) {
          const R* src = j < DX ? a.X + (((size_t)b * N + n) * (T + 1) + t) * DX + j
                       : j < P ? a.U + (((size_t)b * N + n) * T + t) * DU + (j - DX)
                               : a.X + (((size_t)b * N + n) * (T + 1) + t + 1) * DX + (j - P);
          cp_async(xs + r * LDX + j, src);
        } else {
          xs[r * LDX + j] = x0[j];  // shifted value 0: contributes nothing
        }
      }
    };
    const int nchunk = (N + NS - 1) / NS;
    w.team_sync();
    stage(0, 0);

    // ---- 1. S = sum_n (x_n - x0)(x_n - x0)^T (upper blocks), column sums ---------------------------------------
#if !defined(DONK_EMU)
    {
      const int g = w.lane >> 2, tq = w.lane & 3;
      constexpr int NTRI = NBD * (NBD + 1) / 2;
      R acc[NTRI][2];
      R csum[NBD], sh[NBD];
#pragma unroll
      for (int q = 0; q < NTRI; ++q) { acc[q][0] = R(0); acc[q][1] = R(0); }
#pragma unroll
      for (int ib = 0; ib < NBD; ++ib) { csum[ib] = R(0); sh[ib] = x0[8 * ib + g]; }
      for (int ch = 0; ch < nchunk; ++ch) {
        cp_async_wait_all();
        w.team_sync();
        if (ch + 1 < nchunk) stage(ch + 1, (ch + 1) & 1);
        const R* xs = tp + D::oCh + (ch & 1) * NS * LDX;
#pragma unroll
        for (int k4 = 0; k4 < NS / 4; ++k4) {
          R fr[NBD];
#pragma unroll
          for (int ib = 0; ib < NBD; ++ib) {
            fr[ib] = xs[(4 * k4 + tq) * LDX + 8 * ib + g] - sh[ib];
            csum[ib] += fr[ib];
          }
          int q = 0;
#pragma unroll
          for (int ib = 0; ib < NBD; ++ib)
#pragma unroll
            for (int jb = ib; jb < NBD; ++jb, ++q) {
              double d0 = (double)acc[q][0], d1 = (double)acc[q][1];
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(d0), "+d"(d1)
                           : "d"((double)fr[ib]), "d"((double)fr[jb]));
              acc[q][0] = (R)d0;
              acc[q][1] = (R)d1;
            }
        }
        w.team_sync();
      }
      // mean of the shifted samples
#pragma unroll
      for (int ib = 0; ib < NBD; ++ib) {
        R s = csum[ib];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (tq == 0) mu[8 * ib + g] = s / R(N);
      }
      w.team_sync();
      // Sigma = S / N - m m^T, biased (linear_dynamics.py:165), into the blocked storage (both halves)
      int q = 0;
#pragma unroll
      for (int ib = 0; ib < NBD; ++ib)
#pragma unroll
        for (int jb = ib; jb < NBD; ++jb, ++q) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int i = 8 * ib + g, j = 8 * jb + 2 * tq + h;
            if (i <= j && j < d) {
              const R val = acc[q][h] / R(N) - mu[i] * mu[j];
              if (j < P) { Spp[i * PA + j] = val; Spp[j * PA + i] = val; }
              else if (i < P) Spx[i * DXP + (j - P)] = val;
              else { Sxx[(i - P) * DXP + (j - P)] = val; Sxx[(j - P) * DXP + (i - P)] = val; }
            }
          }
        }
    }
#else
    {
      std::vector<R> S((size_t)d * d, R(0)), cs(d, R(0));
      for (int ch = 0; ch < nchunk; ++ch) {
        if (ch + 1 < nchunk) stage(ch + 1, (ch + 1) & 1);
        const R* xs = tp + D::oCh + (ch & 1) * NS * LDX;
        for (int r = 0; r < NS; ++r)
          for (int i = 0; i < d; ++i) {
            const R xi = xs[r * LDX + i] - x0[i];
            cs[i] += xi;
            for (int j = i; j < d; ++j) S[(size_t)i * d + j] = fma(xi, xs[r * LDX + j] - x0[j], S[(size_t)i * d + j]);
          }
      }
      for (int i = 0; i < d; ++i) mu[i] = cs[i] / R(N);
      for (int i = 0; i < d; ++i)
        for (int j = i; j < d; ++j) {
          const R val = S[(size_t)i * d + j] / R(N) - mu[i] * mu[j];
          if (j < P) { Spp[i * PA + j] = val; Spp[j * PA + i] = val; }
          else if (i < P) Spx[i * DXP + (j - P)] = val;
          else { Sxx[(i - P) * DXP + (j - P)] = val; Sxx[(j - P) * DXP + (i - P)] = val; }
        }
    }
#endif
    w.team_sync();

    // warp-level Cholesky of A - shift*I (A symmetric P x P, ld PA) into L (strict lower) and dg (diagonal);
    // returns through `flag` (shared) whether a pivot was not positive
    auto chol = [&](const R* A, R shift, R* L, R* dg, R* flag) {
      LANE_FOR(one, 1) *flag = R(0);
      w.team_sync();
      for (int j = 0; j < P; ++j) {
        LANE_FOR(r, P - j) {
          const int i = j + r;
          R sj = A[j * PA + j] - shift, si = A[i * PA + j];
          for (int k = 0; k < j; ++k) {
            const R ljk = L[j * PA + k];
            sj = fma(-ljk, ljk, sj);
            si = fma(-L[i * PA + k], ljk, si);
          }
          if (i == j) {
            dg[j] = sqrt(sj);
            if (!(sj > R(0))) *flag = R(1);
          } else {
            L[i * PA + j] = si * rsqrt_r(sj);
          }
        }
        w.team_sync();
      }
    };
    R* td = tp + D::oTd;
    R* te = tp + D::oTe;
    R* vv = tp + D::oV;
    R* ww = tp + D::oWw;
    R* scal = tp + D::oScal;

    // ---- 2. does the reference regularise?  lambda_min(Sigma_pp) < reg  <=>  Sigma_pp - reg I is not PD ------------
    chol(Spp, a.reg, W1, td, scal + 0);
    if (scal[0] != R(0)) {  // warp-uniform: slow path, smallest eigenvalue of Sigma_pp (linear_dynamics.py:178)
      R* A = W2;
      LANE_FOR(idx, P * PA) A[idx] = Spp[idx];
      w.team_sync();
      for (int k = 0; k + 2 < P; ++k) {
        LANE_FOR(one, 1) { scal[1] = R(0); scal[2] = R(0); }
        w.team_sync();
        team_reduce_add(w, scal + 1, P - k - 2, [&](int r) { return A[(k + 2 + r) * PA + k] * A[(k + 2 + r) * PA + k]; });
        w.team_sync();
        const R xnorm2 = scal[1], alpha = A[(k + 1) * PA + k];
        R tau = R(0), beta = alpha, scale = R(0);
        if (xnorm2 > R(0)) {
          const R nrm = sqrt(alpha * alpha + xnorm2);
          beta = alpha >= R(0) ? -nrm : nrm;
          tau = (beta - alpha) / beta;
          scale = R(1) / (alpha - beta);
        }
        LANE_FOR(i, P - k - 1) vv[k + 1 + i] = i == 0 ? R(1) : A[(k + 1 + i) * PA + k] * scale;
        w.team_sync();
        if (tau != R(0)) {
          team_reduce_add(w, scal + 2, P - k - 1, [&](int i) {
            const R* row = A + (k + 1 + i) * PA;
            R acc = R(0);
            for (int j = k + 1; j < P; ++j) acc = fma(row[j], vv[j], acc);
            ww[k + 1 + i] = tau * acc;
            return tau * acc * vv[k + 1 + i];
          });
          w.team_sync();
          const R al2 = -tau * scal[2] / 2;
          LANE_FOR(i, P - k - 1) {
            const int ri = k + 1 + i;
            const R vi = vv[ri], wi = fma(al2, vi, ww[ri]);
            for (int j = k + 1; j < P; ++j) A[ri * PA + j] -= vi * fma(al2, vv[j], ww[j]) + wi * vv[j];
          }
        }
        LANE_FOR(one, 1) { td[k] = A[k * PA + k]; te[k] = beta; }
        w.team_sync();
      }
      LANE_FOR(one, 1) {
        if (P >= 2) { td[P - 2] = A[(P - 2) * PA + P - 2]; te[P - 2] = A[(P - 1) * PA + P - 2]; }
        td[P - 1] = A[(P - 1) * PA + P - 1];
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
      R* cnt = vv;  // counts as reals (P <= 32 sections fit)
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

    // ---- 3. Cholesky of Sigma_pp and the two triangular solves for Z = Sigma_pp^-1 Sigma_px = F^T (one right-hand
    //         side per lane).  Substitution keeps the residual Sigma_pp Z - Sigma_px at rounding level, which the Schur
    //         complement below needs when the fit is (nearly) exact; an explicit inverse does not. --------------------
    chol(Spp, R(0), W1, td, scal + 3);
    R* Z = tp + D::oCh;
    R* Gt = Z + P4 * DXP + 8;
    LANE_FOR(idx, 2 * (P4 * DXP + 8)) Z[idx] = R(0);
    w.team_sync();
    LANE_FOR(c, DX) {
      for (int i = 0; i < P; ++i) {
        R acc = Spx[i * DXP + c];
        for (int k = 0; k < i; ++k) acc = fma(-W1[i * PA + k], Z[k * DXP + c], acc);
        Z[i * DXP + c] = acc / td[i];
      }
      for (int i = P - 1; i >= 0; --i) {
        R acc = Z[i * DXP + c];
        for (int k = i + 1; k < P; ++k) acc = fma(-W1[k * PA + i], Z[k * DXP + c], acc);
        Z[i * DXP + c] = acc / td[i];
      }
    }
    w.team_sync();
    // Gt = Sigma_pp Z
    warp_mma<R, D::NBP, D::NBX>(w, Spp, PA, Z, DXP, P4 / 4, [&](int i, int c, R v0, R v1) {
      if (i < P) {
        if (c < DX) Gt[i * DXP + c] = v0;
        if (c + 1 < DX) Gt[i * DXP + c + 1] = v1;
      }
    });
    // outputs F[i][j] = Z[j][i], f = mu_x' - F mu_p (mu = x0 + shifted mean)   (linear_dynamics.py:184-185)
    const size_t bt = (size_t)b * T + t;
    LANE_FOR(idx, DX * P) {
      const int i = idx / P, j = idx % P;
      a.F[bt * DX * P + idx] = Z[j * DXP + i];
    }
    LANE_FOR(i, DX) {
      R acc = R(0);
      for (int j = 0; j < P; ++j) acc = fma(Z[j * DXP + i], x0[j] + mu[j], acc);
      a.f[bt * DX + i] = (x0[P + i] + mu[P + i]) - acc;
    }
    w.team_sync();
    // covar = Sigma_x'x' - Z^T Gt (upper blocks, mirrored)   (linear_dynamics.py:186-188)
    {
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
        w, Gt + 8 * RB_ROW, DXP, Z + 8 * RB_ROW, DXP, P4 / 4,                                              \
        [&](int i, int j, R v0, R v1) { epi(8 * RB_ROW + i, 8 * RB_ROW + j, v0, v1); });
      DONK_FIT_ROW(0) DONK_FIT_ROW(1) DONK_FIT_ROW(2) DONK_FIT_ROW(3)
#undef DONK_FIT_ROW
    }
    LANE_FOR(one, 1) {
      if (a.info) a.info[bt] = scal[3] != R(0);
    }
    w.team_sync();
  }
}

template <typename R, int DX, int DU>
inline int fit_warp_plan(size_t smem_cap, int force_G, int& G, size_t& bytes) {
  using D = FitWarpDims<DX, DU>;
  const size_t per = (size_t)D::TEAM * sizeof(R);
  G = force_G > 0 ? force_G : 4;
  while (G > 1 && (size_t)G * per > smem_cap) --G;
  bytes = (size_t)G * per;
  return bytes <= smem_cap ? DONK_OK : DONK_SMEM_LIMIT;
}

}  // namespace donk
