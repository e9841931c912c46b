// donk_b200 — fit_lr: per-timestep linear-Gaussian dynamics fit, one CTA per (condition, t).
//
// Replaces the Python loop of the reference (file:line in /root/reference):
//   fit_lr                      donk/dynamics/linear_dynamics.py:137-189
//   GMMPrior.eval               donk/dynamics/prior/prior_gmm.py:26-41  (sklearn predict_proba)
//   NormalInverseWishart        donk/dynamics/prior/prior.py:18-45
//
// Stages inside the CTA (all in shared memory, fp64):
//   1. samples [x_t|u_t|x_{t+1}] staged in chunks; mean, then centred second moments as 4x4
//      register tiles over the upper triangle (two passes, like the reference's centred formula);
//   2. optional prior: responsibilities of the samples under the K-cluster GMM (triangular
//      precision-Cholesky products), mixture mean/covariance, closed-form NIW posterior MAP;
//   3. smallest eigenvalue of Sigma[:p,:p] by Householder tridiagonalisation + parallel Sturm
//      multisection, lifted to `regularization` (linear_dynamics.py:178-181);
//   4. column-parallel Cholesky of Sigma_pp, triangular solves for F^T, f, Schur complement.
// The joint covariance is kept in blocks (Sigma_pp, Sigma_px, Sigma_x'x') so that d = 144
// (dX=64, dU=16) fits in 227 KB.
#pragma once
#include "common.cuh"

namespace donk {

template <typename R>
struct FitArgs {
  int B, N, T, dX, dU;
  int NS;  // samples per staged chunk (multiple of 4)
  const R *X, *U;  // (B,N,T+1,dX) (B,N,T,dU)
  int K;           // GMM clusters (0: no prior)
  const R *gw, *gmeans, *gcov, *gpc;  // (K) (K,d) (K,d,d) (K,d,d)
  R n_pool, prior_weight, reg;
  const R* gconst;   // warp kernel with prior: per-cluster constants from gmm_prior_consts_item, (K*d + K)
  // explicit normal-inverse-Wishart prior per (condition, t) (any DynamicsPrior evaluated by the caller, K == 0):
  const R *niw_mu0, *niw_Phi, *niw_Nmean, *niw_Ncovar;  // (B,T,d) (B,T,d,d) (B,T) (B,T) or null
  const R* wts_in;   // cooperative kernel with GMM prior: per-fit cluster weights mean_n predict_proba, (B,T,K)
  int aligned16;     // set by the launcher: X and U are 16-byte aligned
  int redo;          // generic kernel as second pass: only the fits another kernel flagged with info == 2
  int no_prefetch;   // warp kernel: skip the L2 prefetch of the fit's sample rows (DONK_FIT_NO_PREFETCH=1, experiments)
  R *F, *f, *covar;  // (B,T,dX,p) (B,T,dX) (B,T,dX,dX)
  int* info;         // (B,T)
};

struct FitLayout {
  int dX, dU, p, d, dp, pa, dXp, K, Np;
  int oMu, oSpp, oSpx, oSxx, oWork, oTd, oTe, oV, oW, oPart, oLogp, oSP, oWts, oMup, oScal, total;
  enum { NPART = 32, NSEC = 64 };
  DONK_HD FitLayout(int dX_, int dU_, int K_, int N, int NS) {
    dX = dX_; dU = dU_; K = K_; p = dX + dU; d = p + dX;
    dp = round_up(d, 4); pa = round_up(p, 4); dXp = round_up(dX, 4); Np = round_up(N, 4);
    int o = 0;
    oMu = o; o += dp;
    oSpp = o; o += p * pa;
    oSpx = o; o += p * dXp;          // Sigma_px, overwritten by Z = Sigma_pp^-1 Sigma_px
    oSxx = o; o += dX * dXp;
    int work = p * pa;               // tridiagonalisation copy / Cholesky factor / (Sigma_pp Z)
    if (NS * dp > work) work = NS * dp;  // ... and the staged sample chunk
    oWork = o; o += round_up(work, 4);
    oTd = o; o += pa; oTe = o; o += pa; oV = o; o += pa; oW = o; o += pa;
    oPart = o; o += 2 * NSEC;
    oLogp = o; o += K > 0 ? Np * K : 0;
    oSP = o; o += K > 0 ? K * dp : 0;
    oWts = o; o += round_up(K + 4, 4);
    oMup = o; o += K > 0 ? dp : 0;
    oScal = o; o += 8;
    total = o;
  }
};

// Store element (i <= j) of the joint covariance into the blocked layout (both halves).
template <typename R>
DONK_HD void sig_add(R* sm, const FitLayout& L, int i, int j, R val, bool accumulate) {
  const int p = L.p;
  if (j < p) {
    R* a = sm + L.oSpp + i * L.pa + j;
    R* b = sm + L.oSpp + j * L.pa + i;
    const R v = accumulate ? *a + val : val;
    *a = v; *b = v;
  } else if (i < p) {
    R* a = sm + L.oSpx + i * L.dXp + (j - p);
    *a = accumulate ? *a + val : val;
  } else {
    R* a = sm + L.oSxx + (i - p) * L.dXp + (j - p);
    R* b = sm + L.oSxx + (j - p) * L.dXp + (i - p);
    const R v = accumulate ? *a + val : val;
    *a = v; *b = v;
  }
}
template <typename R>
DONK_HD R sig_get(const R* sm, const FitLayout& L, int i, int j) {  // i <= j
  const int p = L.p;
  if (j < p) return sm[L.oSpp + i * L.pa + j];
  if (i < p) return sm[L.oSpx + i * L.dXp + (j - p)];
  return sm[L.oSxx + (i - p) * L.dXp + (j - p)];
}

// One sample value of the joint vector [x_t | u_t | x_{t+1}] from the roll-out arrays.
// Per-cluster constants of the sample log-probabilities (sklearn _estimate_log_gaussian_prob, _gaussian_mixture.py:
// 530-545: y = X P_k - mu_k P_k, logp = -(d log 2pi + |y|^2)/2 + sum log diag P_k + log w_k): they depend on the
// mixture only, not on the fit, so one tiny launch computes them for all (condition, t) fits of a call.
//   out[k*d + j] = -sum_{i<=j} mu_k[i] P_k[i][j]        (k < K, j < d)
//   out[K*d + k] = -d/2 log(2 pi) + sum_i log P_k[i][i] + log w_k
template <typename R>
DONK_HD void gmm_prior_consts_item(int d, int K, const R* gw, const R* gmeans, const R* gpc, R* out, int idx) {
  const size_t dd = (size_t)d * d;
  if (idx < K * d) {
    const int k = idx / d, j = idx % d;
    R acc = R(0);
    for (int i = 0; i <= j; ++i) acc = fma(gmeans[(size_t)k * d + i], gpc[k * dd + (size_t)i * d + j], acc);
    out[idx] = -acc;
  } else if (idx < K * d + K) {
    const int k = idx - K * d;
    R ldet = R(0);
    for (int i = 0; i < d; ++i) ldet += log(gpc[k * dd + (size_t)i * d + i]);
    out[idx] = -R(d) * R(1.8378770664093453) / 2 + ldet + log(gw[k]);
  }
}

template <typename R>
DONK_HD R xux_load(const FitArgs<R>& a, int b, int n, int t, int j) {
  const int dX = a.dX, dU = a.dU, p = dX + dU;
  if (j < dX) return ldg(a.X + (((size_t)b * a.N + n) * (a.T + 1) + t) * dX + j);
  if (j < p) return ldg(a.U + (((size_t)b * a.N + n) * a.T + t) * dU + (j - dX));
  return ldg(a.X + (((size_t)b * a.N + n) * (a.T + 1) + t + 1) * dX + (j - p));
}

template <typename R>
DONK_D void fit_lr_cta(const Ctx& ctx, const FitArgs<R>& a, int cta, R* sm) {
  const FitLayout L(a.dX, a.dU, a.K, a.N, a.NS);
  const int dX = L.dX, p = L.p, d = L.d, dp = L.dp, pa = L.pa, dXp = L.dXp, N = a.N, NS = a.NS, K = a.K;
  const int b = cta / a.T, t = cta % a.T;
  if (a.redo && a.info[cta] != 2) return;  // CTA-uniform: this fit was completed by the cooperative kernel
  const int nchunk = (N + NS - 1) / NS;
  R* xs = sm + L.oWork;
  R* mu = sm + L.oMu;

  // ---- stage a chunk of samples (rows beyond N are zero) ----------------------------------------
  auto stage = [&](int ch, bool centre) {
    const int n0 = ch * NS;
    TEAM_FOR(idx, NS * dp) {
      const int r = idx / dp, j = idx % dp, n = n0 + r;
      R v = R(0);
      if (n < N && j < d) {
        v = xux_load(a, b, n, t, j);
        if (centre) v -= mu[j];
      }
      xs[r * dp + j] = v;
    }
  };

  // ---- 1a. mean (linear_dynamics.py:164) ---------------------------------------------------------
  TEAM_FOR(j, dp) mu[j] = R(0);
  ctx.sync();
  for (int ch = 0; ch < nchunk; ++ch) {
    stage(ch, false);
    ctx.sync();
    TEAM_FOR(j, d) {
      R acc = R(0);
      for (int r = 0; r < NS; ++r) acc += xs[r * dp + j];
      mu[j] += acc;
    }
    ctx.sync();
  }
  TEAM_FOR(j, d) mu[j] = mu[j] / R(N);
  TEAM_FOR(idx, p * pa + p * dXp + dX * dXp) sm[L.oSpp + idx] = R(0);  // the three blocks are contiguous
  ctx.sync();

  // ---- 1b. centred second moments, biased (/N) (linear_dynamics.py:165) ---------------------------
  {
    const int nt = dp / 4, ntri = nt * (nt + 1) / 2;
    for (int ch = 0; ch < nchunk; ++ch) {
      if (nchunk > 1 || ch == 0) {
        if (nchunk > 1) {
          stage(ch, true);
        } else {
          TEAM_FOR(idx, NS * dp) {  // single chunk: still resident, centre in place
            const int r = idx / dp, j = idx % dp;
            if (r < N && j < d) xs[idx] -= mu[j];
          }
        }
        ctx.sync();
      }
      TEAM_FOR(idx, ntri) {
        int ti, tj;
        tri_tile(idx, nt, ti, tj);
        R acc[4][4];
        tile4x4(xs, dp, 4 * ti, xs, dp, 4 * tj, NS, acc);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int i = 4 * ti + r, j = 4 * tj + c;
            if (i <= j && j < d) sig_add(sm, L, i, j, acc[r][c], true);
          }
      }
      ctx.sync();
    }
    TEAM_FOR(idx, p * pa + p * dXp + dX * dXp) sm[L.oSpp + idx] = sm[L.oSpp + idx] / R(N);
    ctx.sync();
  }

  // ---- 2. GMM prior -> NIW posterior MAP covariance (prior_gmm.py:26-41, prior.py:18-45) -------------
  if (K > 0) {
    R* logp = sm + L.oLogp;
    R* sP = sm + L.oSP;
    R* wts = sm + L.oWts;
    R* mup = sm + L.oMup;
    const size_t dd = (size_t)d * d;
    // sP[k][j] = sum_i (mu_emp - mu_k)[i] P_k[i][j]; samples are centred on mu_emp, so
    // (x - mu_k) P_k = xs P_k + sP[k]   (sklearn: X @ P - mu @ P, _gaussian_mixture.py:537-538)
    TEAM_FOR(idx, K * d) {
      const int k = idx / d, j = idx % d;
      const R* P = a.gpc + (size_t)k * dd;
      R acc = R(0);
      for (int i = 0; i <= j; ++i) acc = fma(mu[i] - ldg(a.gmeans + (size_t)k * d + i), ldg(P + (size_t)i * d + j), acc);
      sP[k * dp + j] = acc;
    }
    ctx.sync();
    const int ntj = (d + 3) / 4;
    for (int ch = 0; ch < nchunk; ++ch) {
      if (nchunk > 1) {
        stage(ch, true);
        ctx.sync();
      }
      const int nrt = NS / 4;  // sample tiles in this chunk
      TEAM_FOR(idx, K * nrt) {
        const int k = idx / nrt, rt = idx % nrt;
        const R* P = a.gpc + (size_t)k * dd;
        R ss[4] = {R(0), R(0), R(0), R(0)};
        for (int jt = 0; jt < ntj; ++jt) {
          const int j0 = 4 * jt;
          R acc[4][4];
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = R(0);
          const int kmax = j0 + 4 < d ? j0 + 4 : d;  // P upper triangular: rows i <= j only
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
        R ldet = R(0);
        for (int i = 0; i < d; ++i) ldet += log(ldg(P + (size_t)i * d + i));
        const R lw = log(ldg(a.gw + k));
        const R cst = R(d) * R(1.8378770664093453);  // d * log(2 pi)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int n = ch * NS + 4 * rt + r;
          if (n < N) logp[n * K + k] = -(cst + ss[r]) / 2 + ldet + lw;
        }
      }
      ctx.sync();
    }
    // responsibilities: softmax over clusters (sklearn/mixture/_base.py:552-582)
    TEAM_FOR(n, N) {
      R m = logp[n * K];
      for (int k = 1; k < K; ++k) m = logp[n * K + k] > m ? logp[n * K + k] : m;
      R s = R(0);
      for (int k = 0; k < K; ++k) s += exp(logp[n * K + k] - m);
      const R lse = m + log(s);
      for (int k = 0; k < K; ++k) logp[n * K + k] = exp(logp[n * K + k] - lse);
    }
    ctx.sync();
    TEAM_FOR(k, K) {
      R acc = R(0);
      for (int n = 0; n < N; ++n) acc += logp[n * K + k];
      wts[k] = acc / R(N);  // prior_gmm.py:33
    }
    ctx.sync();
    TEAM_FOR(j, d) {
      R acc = R(0);
      for (int k = 0; k < K; ++k) acc = fma(wts[k], ldg(a.gmeans + (size_t)k * d + j), acc);
      mup[j] = acc;  // prior_gmm.py:35
    }
    if (ctx.tid == 0) {
      R acc = R(0);
      for (int k = 0; k < K; ++k) acc = fma(wts[k], ldg(a.gw + k), acc);
      sm[L.oScal + 0] = acc * a.n_pool;  // prior_gmm.py:39: strength of the prior in samples
    }
    ctx.sync();
    {
      const R Np = sm[L.oScal + 0];
      const R n = R(N) / a.prior_weight;  // linear_dynamics.py:172
      const R coef = Np * n / (Np + n);
      TEAM_FOR(idx, d * d) {
        const int i = idx / d, j = idx % d;
        if (i > j) continue;
        R phi = R(0);
        for (int k = 0; k < K; ++k) phi = fma(wts[k], ldg(a.gcov + (size_t)k * dd + idx), phi);  // prior_gmm.py:36
        const R S = sig_get(sm, L, i, j);
        const R dm = (mu[i] - mup[i]) * (mu[j] - mup[j]);
        // Phi' / (N_covar' + d + 1) with Phi' = Np Phi_p + n S + Np n/(Np+n) dm dm^T  (prior.py:29-45)
        sig_add(sm, L, i, j, (Np * phi + n * S + coef * dm) / (Np + n), false);
      }
    }
    ctx.sync();
  }

  // ---- 2'. explicit NIW prior of this (condition, t): sigma = posterior(mu, S, N / prior_weight).map_covar()
  //          (linear_dynamics.py:171-175, prior.py:18-45), for priors the caller evaluated itself ----------------
  if (K == 0 && a.niw_Phi != nullptr) {
    const size_t bt = (size_t)b * a.T + t;
    const R Nm = a.niw_Nmean[bt], Nc = a.niw_Ncovar[bt];
    const R n = R(N) / a.prior_weight;
    const R coef = Nm * n / (Nm + n), den = Nc + n + R(d) + R(1);
    const R* mu0 = a.niw_mu0 + bt * d;
    const R* Phi = a.niw_Phi + bt * (size_t)d * d;
    TEAM_FOR(idx, d * d) {
      const int i = idx / d, j = idx % d;
      if (i > j) continue;
      const R S = sig_get(sm, L, i, j);
      const R dm = (mu[i] - ldg(mu0 + i)) * (mu[j] - ldg(mu0 + j));
      sig_add(sm, L, i, j, (ldg(Phi + idx) + n * S + coef * dm) / den, false);
    }
    ctx.sync();
  }

  // ---- 3. smallest eigenvalue of Sigma_pp (linear_dynamics.py:178) -------------------------------
  R* A = sm + L.oWork;
  R* td = sm + L.oTd;
  R* te = sm + L.oTe;
  R* vv = sm + L.oV;
  R* ww = sm + L.oW;
  R* part = sm + L.oPart;
  const int NPART = FitLayout::NPART, NSEC = FitLayout::NSEC;
  TEAM_FOR(idx, p * pa) A[idx] = sm[L.oSpp + idx];
  ctx.sync();
  for (int k = 0; k + 2 < p; ++k) {
    // Householder reflector annihilating A[k+2:, k] (LAPACK dsytd2 / dlarfg, lower variant)
    TEAM_FOR(q, NPART) {
      R acc = R(0);
      for (int i = k + 2 + q; i < p; i += NPART) acc = fma(A[i * pa + k], A[i * pa + k], acc);
      part[q] = acc;
    }
    ctx.sync();
    R xnorm2 = R(0);
    for (int q = 0; q < NPART; ++q) xnorm2 += part[q];
    const R alpha = A[(k + 1) * pa + k];
    R tau = R(0), beta = alpha, scale = R(0);
    if (xnorm2 > R(0)) {
      const R nrm = sqrt(alpha * alpha + xnorm2);
      beta = alpha >= R(0) ? -nrm : nrm;
      tau = (beta - alpha) / beta;
      scale = R(1) / (alpha - beta);
    }
    TEAM_FOR(i, p - k - 1) vv[k + 1 + i] = i == 0 ? R(1) : A[(k + 1 + i) * pa + k] * scale;
    ctx.sync();
    if (tau != R(0)) {  // CTA-uniform
      TEAM_FOR(i, p - k - 1) {  // w = tau * A22 v
        const R* row = A + (k + 1 + i) * pa;
        R acc = R(0);
        for (int j = k + 1; j < p; ++j) acc = fma(row[j], vv[j], acc);
        ww[k + 1 + i] = tau * acc;
      }
      ctx.sync();
      TEAM_FOR(q, NPART) {
        R acc = R(0);
        for (int i = k + 1 + q; i < p; i += NPART) acc = fma(ww[i], vv[i], acc);
        part[NSEC + q] = acc;
      }
      ctx.sync();
      R dot = R(0);
      for (int q = 0; q < NPART; ++q) dot += part[NSEC + q];
      const R al2 = -tau * dot / 2;
      const int m = p - k - 1;
      TEAM_FOR(idx, m * m) {  // A22 -= v w'^T + w' v^T,  w' = w + al2 v
        const int i = k + 1 + idx / m, j = k + 1 + idx % m;
        const R wi = fma(al2, vv[i], ww[i]), wj = fma(al2, vv[j], ww[j]);
        A[i * pa + j] -= vv[i] * wj + wi * vv[j];
      }
    }
    if (ctx.tid == 0) { td[k] = A[k * pa + k]; te[k] = beta; }
    ctx.sync();
  }
  if (ctx.tid == 0) {
    if (p >= 2) {
      td[p - 2] = A[(p - 2) * pa + p - 2];
      te[p - 2] = A[(p - 1) * pa + p - 2];
    }
    td[p - 1] = A[(p - 1) * pa + p - 1];
  }
  ctx.sync();
  // Sturm multisection for the smallest eigenvalue of tridiag(td, te)
  R lo, hi;
  {
    lo = td[0]; hi = td[0];
    R emax = R(0);
    for (int i = 0; i < p; ++i) {
      const R el = i > 0 ? fabs(te[i - 1]) : R(0), er = i + 1 < p ? fabs(te[i]) : R(0);
      const R g = td[i] - el - er;
      lo = g < lo ? g : lo;
      hi = td[i] < hi ? td[i] : hi;  // lambda_min <= min diagonal
      emax = el > emax ? el : emax;
    }
    const R tnorm = fabs(lo) > fabs(hi) ? fabs(lo) : fabs(hi);
    const R pivmin = (sizeof(R) == 8 ? R(2.3e-308) : R(1.2e-38)) * (emax * emax > R(1) ? emax * emax : R(1));
    const R eps = sizeof(R) == 8 ? R(2.220446049250313e-16) : R(1.1920929e-7);
    lo -= R(2) * eps * tnorm * R(p) + R(2) * pivmin;
    hi += R(2) * eps * tnorm * R(p) + R(2) * pivmin;
    int* cnt = reinterpret_cast<int*>(part);
    for (int round = 0; round < 12; ++round) {
      const R width = hi - lo;
      TEAM_FOR(q, NSEC) {
        const R x = lo + width * R(q + 1) / R(NSEC + 1);
        int c = 0;
        R piv = td[0] - x;
        if (fabs(piv) < pivmin) piv = -pivmin;
        c += piv < R(0);
        for (int i = 1; i < p; ++i) {
          piv = td[i] - x - te[i - 1] * te[i - 1] / piv;
          if (fabs(piv) < pivmin) piv = -pivmin;
          c += piv < R(0);
        }
        cnt[q] = c;  // eigenvalues below x
      }
      ctx.sync();
      int first = NSEC;  // first section point with at least one eigenvalue below it
      for (int q = NSEC - 1; q >= 0; --q)
        if (cnt[q] > 0) first = q;
      const R nlo = first == 0 ? lo : lo + width * R(first) / R(NSEC + 1);
      const R nhi = first == NSEC ? hi : lo + width * R(first + 1) / R(NSEC + 1);
      lo = nlo; hi = nhi;
      ctx.sync();
      if (hi - lo <= eps * tnorm) break;  // CTA-uniform
    }
  }
  const R min_eig = (lo + hi) / 2;
  if (min_eig < a.reg) {  // lift the smallest eigenvalue to `regularization` (linear_dynamics.py:179-181)
    const R shift = a.reg - min_eig;
    TEAM_FOR(i, p) sm[L.oSpp + i * pa + i] += shift;
  }
  ctx.sync();

  // ---- 4. F^T = Sigma_pp^-1 Sigma_px by Cholesky (the reference uses LU, linear_dynamics.py:184) ----
  R* Lm = sm + L.oWork;
  if (ctx.tid == 0) sm[L.oScal + 1] = R(0);  // not-PD flag
  ctx.sync();
  for (int j = 0; j < p; ++j) {
    TEAM_FOR(r, p - j) {
      const int i = j + r;
      const R* Sj = sm + L.oSpp + j * pa;
      R sj = Sj[j], si = sm[L.oSpp + i * pa + j];
      for (int k = 0; k < j; ++k) {
        const R ljk = Lm[j * pa + k];
        sj = fma(-ljk, ljk, sj);
        si = fma(-Lm[i * pa + k], ljk, si);
      }
      const R dj = sqrt(sj);
      if (i == j) {
        td[j] = dj;  // diagonal of L kept apart: rows i > j read only the strict lower part
        if (!(sj > R(0))) sm[L.oScal + 1] = R(1);
      } else {
        Lm[i * pa + j] = si / dj;
      }
    }
    ctx.sync();
  }
  {
    R* Z = sm + L.oSpx;
    const size_t bt = (size_t)b * a.T + t;
    TEAM_FOR(c, dX) {  // one right-hand side per item: L y = b
      for (int i = 0; i < p; ++i) {
        R acc = Z[i * dXp + c];
        for (int k = 0; k < i; ++k) acc = fma(-Lm[i * pa + k], Z[k * dXp + c], acc);
        Z[i * dXp + c] = acc / td[i];
      }
    }
    ctx.sync();
    // covar = Sigma_x'x' - Y^T Y with Y = L^-1 Sigma_px: the Schur complement of linear_dynamics.py:186
    // (F Sigma_pp F^T = Sigma_xp Sigma_pp^-1 Sigma_px = Y^T Y) as the difference of the sample block and ONE Gram
    // matrix - no second product through F, whose rounding pushed the smallest eigenvalue of a singular result
    // to -1.8e-16 (the reference's own test tolerates -1e-16, tests/dynamics_test.py:85-90).
    {
      const int nt = dXp / 4, ntri = nt * (nt + 1) / 2;
      TEAM_FOR(idx, ntri) {
        int ti, tj;
        tri_tile(idx, nt, ti, tj);
        R acc[4][4];
        tile4x4(Z, dXp, 4 * ti, Z, dXp, 4 * tj, p, acc);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int i = 4 * ti + r, j = 4 * tj + c;
            if (i <= j && j < dX) {
              const R val = sm[L.oSxx + i * dXp + j] - acc[r][c];
              a.covar[bt * dX * dX + (size_t)i * dX + j] = val;  // symmetric by construction (linear_dynamics.py:188)
              a.covar[bt * dX * dX + (size_t)j * dX + i] = val;
            }
          }
      }
    }
    ctx.sync();
    TEAM_FOR(c, dX) {  // L^T z = y: F^T = Sigma_pp^-1 Sigma_px
      for (int i = p - 1; i >= 0; --i) {
        R acc = Z[i * dXp + c];
        for (int k = i + 1; k < p; ++k) acc = fma(-Lm[k * pa + i], Z[k * dXp + c], acc);
        Z[i * dXp + c] = acc / td[i];
      }
    }
    ctx.sync();
    // outputs: F[i][j] = Z[j][i]; f = mu_x' - F mu_p   (linear_dynamics.py:184-185)
    TEAM_FOR(idx, dX * p) {
      const int i = idx / p, j = idx % p;
      a.F[bt * dX * p + idx] = Z[j * dXp + i];
    }
    TEAM_FOR(i, dX) {
      R acc = R(0);
      for (int j = 0; j < p; ++j) acc = fma(Z[j * dXp + i], mu[j], acc);
      a.f[bt * dX + i] = mu[p + i] - acc;
    }
    ctx.sync();
    if (ctx.tid == 0 && a.info) a.info[bt] = sm[L.oScal + 1] != R(0);
  }
}

// Samples per chunk so that the layout fits `cap` bytes; 0 if even NS = 4 does not fit.
inline int fit_plan(int dX, int dU, int K, int N, size_t sz, size_t cap, size_t& bytes) {
  int NS = round_up(N, 4);
  while (NS >= 4) {
    FitLayout L(dX, dU, K, N, NS);
    bytes = (size_t)L.total * sz;
    if (bytes <= cap) return NS;
    NS = NS > 64 ? round_up(NS / 2, 4) : NS - 4;
  }
  return 0;
}

// to_transitions (donk/dynamics/utils.py:4-17): row n*T+t = [x_t | u_t | x_{t+1}].
template <typename R>
DONK_HD R transitions_item(int T, int dX, int dU, const R* X, const R* U, size_t idx) {
  const int d = 2 * dX + dU, p = dX + dU;
  const size_t row = idx / d;
  const int j = (int)(idx % d);
  const size_t n = row / T;
  const int t = (int)(row % T);
  if (j < dX) return X[(n * (T + 1) + t) * dX + j];
  if (j < p) return U[(n * T + t) * dU + (j - dX)];
  return X[(n * (T + 1) + t + 1) * dX + (j - p)];
}

// StateDistribution.fit (donk/samples/state_dist.py:27-37): mean and unbiased covariance; item = (b,i,j).
template <typename R>
DONK_HD void state_fit_item(int N, int dX, const R* X0, size_t sample_stride, size_t cond_stride, size_t idx, R* mean,
                            R* covar) {
  const size_t b = idx / ((size_t)dX * dX);
  const int r = (int)(idx % ((size_t)dX * dX)), i = r / dX, j = r % dX;
  const R* x = X0 + b * cond_stride;
  R mi = R(0), mj = R(0);
  for (int n = 0; n < N; ++n) { mi += x[n * sample_stride + i]; mj += x[n * sample_stride + j]; }
  mi /= R(N); mj /= R(N);
  R acc = R(0);
  for (int n = 0; n < N; ++n) acc = fma(x[n * sample_stride + i] - mi, x[n * sample_stride + j] - mj, acc);
  covar[idx] = acc / R(N - 1);
  if (j == 0) mean[b * dX + i] = mi;
}

}  // namespace donk
