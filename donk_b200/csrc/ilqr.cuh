// donk_b200 — KL-constrained iLQR step: backward Riccati recursion, forward Gaussian
// pass, KL divergence and expected cost, fused in one persistent CTA.
//
// Replaces the Python loops of the reference (all file:line in /root/reference):
//   ILQR.step              donk/traj_opt/lqg.py:55-75
//   backward               donk/traj_opt/lqg.py:126-182
//   forward                donk/traj_opt/lqg.py:185-244
//   kl_divergence_action   donk/traj_opt/lqg.py:276-306
//   expected_costs         donk/costs/quadratic_costs.py:55-92,106-119
//
// Work decomposition.  A CTA owns S "slots"; a slot is one (condition b, eta e)
// problem and keeps its whole working set in shared memory for all T steps:
//   M  (pp x pp)  Q_t / V_t during backward, Sigma_t during forward
//   Rb            W = V F (dX x pp) during backward, (F Sigma)^T (p x dXp) forward
//   Kb, kb, Sinv  K_t, k_t, Quu^-1 of the current step
// Slots of one condition share the staged F_t (the eta candidates of a condition
// read the dynamics once).  The two O(dX p^2) products per step run as 4x4 register
// tiles with both operands k-major in shared memory (4 LDS.128 per 16 DFMA); the
// dU x dU Cholesky/inverse runs column-parallel.  T is sequential, as in the
// reference; parallelism comes from slots x tiles.
#pragma once
#include "common.cuh"

namespace donk {

enum { ILQR_BACKWARD = 1, ILQR_FORWARD = 2, ILQR_KL = 4, ILQR_COST = 8 };

template <typename R>
struct IlqrArgs {
  int B, E, T, dX, dU;
  int S;   // slots per CTA
  int ES;  // eta candidates of one condition handled per CTA (S % ES == 0)
  int flags;
  const R *F, *f, *dyn_covar;                    // (B,T,dX,p) (B,T,dX) (B,T,dX,dX)
  const R *C, *c, *cc;                           // (B,T+1,p,p) (B,T+1,p) (B,T+1)
  const R *C_kl, *c_kl;                          // (B,T,p,p) (B,T,p) or null
  const R *prevK, *prevk, *prevLam, *prev_hld;   // (B,T,dU,dX) (B,T,dU) (B,T,dU,dU) (B,T)
  const R *x0m, *x0S;                            // (B,dX) (B,dX,dX)
  const R *eta;                                  // (B,E) or null (eta = 1)
  R gamma, fwd_reg;
  R *K, *k, *pcov, *pinv;                        // (B,E,T,dU,dX) (B,E,T,dU) (B,E,T,dU,dU) x2
  R *kl, *cost;                                  // (B,E)
  R *tmean, *tcov;                               // optional (B,E,T+1,p) (B,E,T+1,p,p)
  int *info;                                     // (B,E): 0 ok, 1+t = Quu not PD at step t
  const int *active;                             // optional (B): 0 = skip this condition
  int aligned16;                                 // set by the launcher: F, C_kl, dyn_covar are 16-byte aligned
};

// Shared-memory layout, in units of R.  Every sub-array starts 16-byte aligned.
struct IlqrLayout {
  int dX, dU, p, dXp, pp;
  int oM, oR, oK, okb, oSinv, oL, oLi, oLd, oq, ow, ov, omun, odelta, orow, oscal, slot;
  int oFs, ofs, cond;
  DONK_HD IlqrLayout(int dX_, int dU_) {
    dX = dX_; dU = dU_; p = dX + dU;
    dXp = round_up(dX, 4); pp = round_up(p, 4);
    int o = 0;
    oM = o; o += pp * pp;
    int rsz = dX * pp > p * dXp ? dX * pp : p * dXp;
    oR = o; o += round_up(rsz, 4);
    oK = o; o += round_up(dU * dXp, 4);
    okb = o; o += round_up(dU, 4);
    oSinv = o; o += round_up(dU * dU, 4);
    oL = o; o += round_up(dU * dU, 4);
    oLi = o; o += round_up(dU * dU, 4);
    oLd = o; o += round_up(dU, 4);
    oq = o; o += pp;      // q (backward) / mu (forward)
    ow = o; o += dXp;     // w = V f + v
    ov = o; o += dXp;     // v
    omun = o; o += dXp;   // next state mean
    odelta = o; o += round_up(dU, 4);
    orow = o; o += pp;    // per-row cost partials
    oscal = o; o += 8;    // [0] eta [1] kl_acc [2] cost_acc [3] logdet_acc
    slot = o;
    int fsz = dX * pp > p * dXp ? dX * pp : p * dXp;
    oFs = 0; ofs = round_up(fsz, 4); cond = ofs + dXp;
  }
  DONK_HD size_t reals(int S, int NC) const { return (size_t)S * slot + (size_t)NC * cond; }
  // meta ints: per slot {b, e, valid, notpd}; per condition {b}
  DONK_HD size_t bytes(int S, int NC, size_t sz) const {
    const size_t b = reals(S, NC) * sz + (size_t)(4 * S + NC) * sizeof(int);
    return (b + 15) / 16 * 16;  // regions of consecutive teams stay 16-byte aligned
  }
};

template <typename R>
DONK_D void ilqr_cta(const Ctx& ctx, const IlqrArgs<R>& a, int cta, R* smem) {
  const IlqrLayout L(a.dX, a.dU);
  const int dX = L.dX, dU = L.dU, p = L.p, dXp = L.dXp, pp = L.pp, T = a.T, E = a.E;
  const int S = a.S, ES = a.ES, NC = S / ES;
  const int n_ech = (E + ES - 1) / ES;
  R* slots = smem;
  R* conds = smem + (size_t)S * L.slot;
  int* meta = reinterpret_cast<int*>(conds + (size_t)NC * L.cond);
  int* cmeta = meta + 4 * S;
  const R gamma = a.gamma;

  // ---- slot assignment -------------------------------------------------------------------
  {
    const int cg = cta / n_ech, ec = cta % n_ech;
    TEAM_FOR(s, S) {
      const int ci = s / ES, ei = s % ES;
      const int b = cg * NC + ci, e = ec * ES + ei;
      const bool valid = b < a.B && e < E && (a.active == nullptr || a.active[b] != 0);
      meta[4 * s + 0] = b; meta[4 * s + 1] = e; meta[4 * s + 2] = valid ? 1 : 0; meta[4 * s + 3] = 0;
      if (ei == 0) cmeta[ci] = valid ? b : -1;
    }
    ctx.sync();
    int any = 0;
    for (int s = 0; s < S; ++s) any |= meta[4 * s + 2];
    if (!any) return;  // CTA-uniform
  }
#define SLOT(s) (slots + (size_t)(s) * L.slot)
#define SB(s) meta[4 * (s) + 0]
#define SE(s) meta[4 * (s) + 1]
#define SVALID(s) meta[4 * (s) + 2]
#define PIDX(s) ((size_t)SB(s) * E + SE(s))  // problem index into (B,E,...) outputs

  const int ntx = dXp / 4, ntp = pp / 4;
  const size_t szF = (size_t)dX * p, szC = (size_t)p * p, szK = (size_t)dU * dX, szS = (size_t)dU * dU;

  TEAM_FOR(s, S) {
    R* sc = SLOT(s) + L.oscal;
    sc[0] = (SVALID(s) && a.eta) ? a.eta[PIDX(s)] : R(1);
    sc[1] = R(0); sc[2] = R(0); sc[3] = R(0);
  }
  ctx.sync();

  // =========================================================================================
  // Backward pass (lqg.py:126-182) under C_ext = C/eta + C_kl, c_ext = c/eta + c_kl (lqg.py:62-65)
  // =========================================================================================
  if (a.flags & ILQR_BACKWARD) {
    // terminal value: V = C_ext[T,:dX,:dX], v = c_ext[T,:dX]  (lqg.py:151-152; C_kl has no row T)
    TEAM_FOR(idx, S * dX * dX) {
      const int s = idx / (dX * dX), r = idx % (dX * dX), i = r / dX, j = r % dX;
      if (!SVALID(s)) continue;
      R* sl = SLOT(s);
      const R eta = sl[L.oscal];
      R val = ldg(a.C + ((size_t)SB(s) * (T + 1) + T) * szC + (size_t)i * p + j);
      if (a.eta) val = val / eta;
      sl[L.oM + i * pp + j] = val;
    }
    TEAM_FOR(idx, S * dX) {
      const int s = idx / dX, i = idx % dX;
      if (!SVALID(s)) continue;
      R* sl = SLOT(s);
      R val = ldg(a.c + ((size_t)SB(s) * (T + 1) + T) * p + i);
      if (a.eta) val = val / sl[L.oscal];
      sl[L.ov + i] = val;
    }
    ctx.sync();

    for (int t = T - 1; t >= 0; --t) {
      // ---- stage F_t (dX x pp, zero padded) and f_t per condition --------------------------
      TEAM_FOR(idx, NC * dX * pp) {
        const int ci = idx / (dX * pp), r = idx % (dX * pp), k = r / pp, j = r % pp;
        const int b = cmeta[ci];
        if (b < 0) continue;
        R* cd = conds + (size_t)ci * L.cond;
        cd[L.oFs + k * pp + j] = j < p ? ldg(a.F + ((size_t)b * T + t) * szF + (size_t)k * p + j) : R(0);
      }
      TEAM_FOR(idx, NC * dX) {
        const int ci = idx / dX, k = idx % dX;
        const int b = cmeta[ci];
        if (b < 0) continue;
        conds[(size_t)ci * L.cond + L.ofs + k] = ldg(a.f + ((size_t)b * T + t) * dX + k);
      }
      ctx.sync();

      // ---- B1: W = V F_t (dX x p);  w = V f_t + v ------------------------------------------
      {
        const int ntile = ntx * ntp;
        TEAM_FOR(idx, S * ntile + S * dX) {
          if (idx < S * ntile) {
            const int s = idx / ntile, tl = idx % ntile, ti = tl / ntp, tj = tl % ntp;
            if (!SVALID(s)) continue;
            R* sl = SLOT(s);
            const R* Fs = conds + (size_t)(s / ES) * L.cond + L.oFs;
            R acc[4][4];
            tile4x4(sl + L.oM, pp, 4 * ti, Fs, pp, 4 * tj, dX, acc);  // V symmetric: V[i][k] = M[k][i]
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int i = 4 * ti + r;
              if (i < dX) {
#pragma unroll
                for (int c = 0; c < 4; ++c) sl[L.oR + i * pp + 4 * tj + c] = acc[r][c];
              }
            }
          } else {
            const int v = idx - S * ntile, s = v / dX, i = v % dX;
            if (!SVALID(s)) continue;
            R* sl = SLOT(s);
            const R* fs = conds + (size_t)(s / ES) * L.cond + L.ofs;
            R acc = R(0);
            for (int k = 0; k < dX; ++k) acc = fma(sl[L.oM + k * pp + i], fs[k], acc);
            sl[L.ow + i] = acc + sl[L.ov + i];
          }
        }
      }
      ctx.sync();

      // ---- B2: Q = C_ext[t] + gamma F^T W (upper tiles, mirrored);  q = c_ext[t] + gamma F^T w
      {
        const int ntri = ntp * (ntp + 1) / 2;
        TEAM_FOR(idx, S * ntri + S * p) {
          if (idx < S * ntri) {
            const int s = idx / ntri;
            if (!SVALID(s)) continue;
            int ti, tj;
            tri_tile(idx % ntri, ntp, ti, tj);
            R* sl = SLOT(s);
            const R* Fs = conds + (size_t)(s / ES) * L.cond + L.oFs;
            R acc[4][4];
            tile4x4(Fs, pp, 4 * ti, sl + L.oR, pp, 4 * tj, dX, acc);
            const R eta = sl[L.oscal];
            const size_t cb = ((size_t)SB(s) * (T + 1) + t) * szC, kb = ((size_t)SB(s) * T + t) * szC;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const int i = 4 * ti + r, j = 4 * tj + c;
                if (i <= j && j < p) {
                  R cv = ldg(a.C + cb + (size_t)i * p + j);
                  if (a.eta) cv = cv / eta;
                  if (a.C_kl) cv += ldg(a.C_kl + kb + (size_t)i * p + j);
                  const R val = cv + gamma * acc[r][c];
                  sl[L.oM + i * pp + j] = val;
                  sl[L.oM + j * pp + i] = val;
                }
              }
            }
          } else {
            const int v = idx - S * ntri, s = v / p, i = v % p;
            if (!SVALID(s)) continue;
            R* sl = SLOT(s);
            const R* Fs = conds + (size_t)(s / ES) * L.cond + L.oFs;
            R acc = R(0);
            for (int k = 0; k < dX; ++k) acc = fma(Fs[k * pp + i], sl[L.ow + k], acc);
            R cv = ldg(a.c + ((size_t)SB(s) * (T + 1) + t) * p + i);
            if (a.eta) cv = cv / sl[L.oscal];
            if (a.c_kl) cv += ldg(a.c_kl + ((size_t)SB(s) * T + t) * p + i);
            sl[L.oq + i] = cv + gamma * acc;
          }
        }
      }
      ctx.sync();

      // ---- B3: Quu = L L^T column by column; Quu^-1 = L^-T L^-1 (lqg.py:165-166) ------------
      for (int j = 0; j < dU; ++j) {
        const int rows = dU - j;
        TEAM_FOR(idx, S * rows) {
          const int s = idx / rows, i = j + idx % rows;
          if (!SVALID(s)) continue;
          R* sl = SLOT(s);
          const R* Quu = sl + L.oM + dX * pp + dX;
          R* Lm = sl + L.oL;
          R sj = Quu[j * pp + j], si = Quu[i * pp + j];
          for (int k = 0; k < j; ++k) {
            const R ljk = Lm[j * dU + k];
            sj = fma(-ljk, ljk, sj);
            si = fma(-Lm[i * dU + k], ljk, si);
          }
          const R d = sqrt(sj);
          if (i == j) {
            sl[L.oLd + j] = d;
            if (!(sj > R(0)) && meta[4 * s + 3] == 0) meta[4 * s + 3] = t + 1;
          } else {
            Lm[i * dU + j] = si / d;
          }
        }
        ctx.sync();
      }
      TEAM_FOR(idx, S * dU) {  // column c of L^-1 by forward substitution
        const int s = idx / dU, c = idx % dU;
        if (!SVALID(s)) continue;
        R* sl = SLOT(s);
        const R* Lm = sl + L.oL;
        const R* Ld = sl + L.oLd;
        R* Li = sl + L.oLi;
        Li[c * dU + c] = R(1) / Ld[c];
        for (int i = c + 1; i < dU; ++i) {
          R acc = R(0);
          for (int k = c; k < i; ++k) acc = fma(Lm[i * dU + k], Li[k * dU + c], acc);
          Li[i * dU + c] = -acc / Ld[i];
        }
        if (c == 0) {
          R ld = R(0);
          for (int j = 0; j < dU; ++j) ld += log(Ld[j]);
          sl[L.oscal + 3] += ld;  // sum_t sum_j log diag chol(Quu_t) = -sum_t log diag chol(pol_covar_t)
        }
      }
      ctx.sync();
      TEAM_FOR(idx, S * dU * dU) {
        const int s = idx / (dU * dU), r = idx % (dU * dU), i = r / dU, j = r % dU;
        if (!SVALID(s)) continue;
        R* sl = SLOT(s);
        const size_t ob = (PIDX(s) * T + t) * szS;
        a.pinv[ob + r] = sl[L.oM + (dX + i) * pp + dX + j];
        if (j <= i) {
          const R* Li = sl + L.oLi;
          R acc = R(0);
          for (int k = i; k < dU; ++k) acc = fma(Li[k * dU + i], Li[k * dU + j], acc);
          sl[L.oSinv + i * dU + j] = acc;
          sl[L.oSinv + j * dU + i] = acc;
          a.pcov[ob + i * dU + j] = acc;
          a.pcov[ob + j * dU + i] = acc;
        }
      }
      ctx.sync();

      // ---- B4: K = -Quu^-1 Qux, k = -Quu^-1 q_u (lqg.py:170-172) ---------------------------
      TEAM_FOR(idx, S * dU * (dX + 1)) {
        const int s = idx / (dU * (dX + 1)), r = idx % (dU * (dX + 1));
        if (!SVALID(s)) continue;
        R* sl = SLOT(s);
        const R* Si = sl + L.oSinv;
        if (r < dU * dX) {
          const int ai = r / dX, j = r % dX;
          R acc = R(0);
          for (int b2 = 0; b2 < dU; ++b2) acc = fma(Si[ai * dU + b2], sl[L.oM + (dX + b2) * pp + j], acc);
          sl[L.oK + ai * dXp + j] = -acc;
          a.K[(PIDX(s) * T + t) * szK + r] = -acc;
        } else {
          const int ai = r - dU * dX;
          R acc = R(0);
          for (int b2 = 0; b2 < dU; ++b2) acc = fma(Si[ai * dU + b2], sl[L.oq + dX + b2], acc);
          sl[L.okb + ai] = -acc;
          a.k[(PIDX(s) * T + t) * dU + ai] = -acc;
        }
      }
      ctx.sync();

      // ---- B5: V = Qxx + Qxu K (upper tiles, mirrored);  v = q_x + Qxu k (lqg.py:178-180) ----
      {
        const int ntri = ntx * (ntx + 1) / 2;
        TEAM_FOR(idx, S * ntri + S * dX) {
          if (idx < S * ntri) {
            const int s = idx / ntri;
            if (!SVALID(s)) continue;
            int ti, tj;
            tri_tile(idx % ntri, ntx, ti, tj);
            R* sl = SLOT(s);
            R acc[4][4];
            tile4x4(sl + L.oM + dX * pp, pp, 4 * ti, sl + L.oK, dXp, 4 * tj, dU, acc);  // Qxu[i][a] = M[dX+a][i]
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const int i = 4 * ti + r, j = 4 * tj + c;
                if (i <= j && j < dX) {
                  const R val = sl[L.oM + i * pp + j] + acc[r][c];
                  sl[L.oM + i * pp + j] = val;
                  sl[L.oM + j * pp + i] = val;
                }
              }
            }
          } else {
            const int v = idx - S * ntri, s = v / dX, i = v % dX;
            if (!SVALID(s)) continue;
            R* sl = SLOT(s);
            R acc = R(0);
            for (int b2 = 0; b2 < dU; ++b2) acc = fma(sl[L.oM + (dX + b2) * pp + i], sl[L.okb + b2], acc);
            sl[L.ov + i] = sl[L.oq + i] + acc;
          }
        }
      }
      ctx.sync();
    }
  }

  // =========================================================================================
  // Forward pass (lqg.py:185-244) with KL (lqg.py:276-306) and expected cost accumulated on the fly
  // =========================================================================================
  if (a.flags & ILQR_FORWARD) {
    const R reg = a.fwd_reg;
    const bool do_kl = (a.flags & ILQR_KL) != 0, do_cost = (a.flags & ILQR_COST) != 0;
    TEAM_FOR(idx, S * dX * dX) {
      const int s = idx / (dX * dX), r = idx % (dX * dX), i = r / dX, j = r % dX;
      if (!SVALID(s)) continue;
      SLOT(s)[L.oM + i * pp + j] = ldg(a.x0S + (size_t)SB(s) * dX * dX + r);
    }
    TEAM_FOR(idx, S * dX) {
      const int s = idx / dX, i = idx % dX;
      if (!SVALID(s)) continue;
      SLOT(s)[L.oq + i] = ldg(a.x0m + (size_t)SB(s) * dX + i);
    }
    ctx.sync();

    for (int t = 0; t < T; ++t) {
      // ---- stage F_t^T (p x dXp), f_t per condition; K_t, k_t, pol_covar_t per slot ---------
      TEAM_FOR(idx, NC * dX * p) {
        const int ci = idx / (dX * p), r = idx % (dX * p), i = r / p, k = r % p;
        const int b = cmeta[ci];
        if (b < 0) continue;
        conds[(size_t)ci * L.cond + L.oFs + k * dXp + i] = ldg(a.F + ((size_t)b * T + t) * szF + r);
      }
      TEAM_FOR(idx, NC * dX) {
        const int ci = idx / dX, k = idx % dX;
        const int b = cmeta[ci];
        if (b < 0) continue;
        conds[(size_t)ci * L.cond + L.ofs + k] = ldg(a.f + ((size_t)b * T + t) * dX + k);
      }
      {
        const int per = dU * dX + dU + dU * dU;
        TEAM_FOR(idx, S * per) {
          const int s = idx / per, r = idx % per;
          if (!SVALID(s)) continue;
          R* sl = SLOT(s);
          const size_t pt = PIDX(s) * T + t;
          if (r < dU * dX) sl[L.oK + (r / dX) * dXp + r % dX] = a.K[pt * szK + r];  // plain loads: may be
          else if (r < dU * dX + dU) sl[L.okb + r - dU * dX] = a.k[pt * dU + r - dU * dX];  // written by this CTA
          else sl[L.oSinv + r - dU * dX - dU] = a.pcov[pt * szS + r - dU * dX - dU];
        }
      }
      ctx.sync();

      // ---- F1: Sigma_ux = K Sigma_xx (and its transpose); mu_u = K mu_x + k (lqg.py:227-230) -
      TEAM_FOR(idx, S * dU * (dX + 1)) {
        const int s = idx / (dU * (dX + 1)), r = idx % (dU * (dX + 1));
        if (!SVALID(s)) continue;
        R* sl = SLOT(s);
        const int ai = r / (dX + 1), j = r % (dX + 1);
        const R* Kr = sl + L.oK + ai * dXp;
        R acc = R(0);
        if (j < dX) {
          for (int k = 0; k < dX; ++k) acc = fma(Kr[k], sl[L.oM + k * pp + j], acc);
          sl[L.oM + (dX + ai) * pp + j] = acc;
          sl[L.oM + j * pp + dX + ai] = acc;
        } else {
          for (int k = 0; k < dX; ++k) acc = fma(Kr[k], sl[L.oq + k], acc);
          sl[L.oq + dX + ai] = acc + sl[L.okb + ai];
        }
      }
      ctx.sync();
      // ---- F2: Sigma_uu = Sigma_ux K^T + pol_covar (lqg.py:231) ----------------------------
      TEAM_FOR(idx, S * dU * dU) {
        const int s = idx / (dU * dU), r = idx % (dU * dU), ai = r / dU, bi = r % dU;
        if (!SVALID(s) || bi < ai) continue;
        R* sl = SLOT(s);
        R acc = R(0);
        for (int k = 0; k < dX; ++k) acc = fma(sl[L.oM + (dX + ai) * pp + k], sl[L.oK + bi * dXp + k], acc);
        acc += sl[L.oSinv + ai * dU + bi];
        sl[L.oM + (dX + ai) * pp + dX + bi] = acc;
        sl[L.oM + (dX + bi) * pp + dX + ai] = acc;
      }
      ctx.sync();

      // ---- F3: Rt = (F Sigma_t)^T (p x dX); mu' = F mu + f; cost rows; KL delta; outputs -----
      {
        const int ntile = ntp * ntx;
        const int n_cost = do_cost ? S * p : 0, n_kl = do_kl ? S * dU : 0;
        const int n_out = a.tcov ? S * p * p : 0, n_outm = a.tmean ? S * p : 0;
        const int o1 = S * ntile, o2 = o1 + S * dX, o3 = o2 + n_cost, o4 = o3 + n_kl, o5 = o4 + n_out, o6 = o5 + n_outm;
        TEAM_FOR(idx, o6) {
          if (idx < o1) {
            const int s = idx / ntile, tl = idx % ntile, ti = tl / ntx, tj = tl % ntx;
            if (!SVALID(s)) continue;
            R* sl = SLOT(s);
            const R* Ft = conds + (size_t)(s / ES) * L.cond + L.oFs;
            R acc[4][4];
            tile4x4(sl + L.oM, pp, 4 * ti, Ft, dXp, 4 * tj, p, acc);  // Rt[j][i] = sum_k Sigma[k][j] F[i][k]
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int j = 4 * ti + r;
              if (j < p) {
#pragma unroll
                for (int c = 0; c < 4; ++c) sl[L.oR + j * dXp + 4 * tj + c] = acc[r][c];
              }
            }
          } else if (idx < o2) {
            const int v = idx - o1, s = v / dX, i = v % dX;
            if (!SVALID(s)) continue;
            R* sl = SLOT(s);
            const R* cd = conds + (size_t)(s / ES) * L.cond;
            R acc = R(0);
            for (int k = 0; k < p; ++k) acc = fma(cd[L.oFs + k * dXp + i], sl[L.oq + k], acc);
            sl[L.omun + i] = acc + cd[L.ofs + i];
          } else if (idx < o3) {
            // row i of  1/2 mu^T C mu + mu^T c + 1/2 tr(C (Sigma + reg I))   (quadratic_costs.py:84-91,118)
            const int v = idx - o2, s = v / p, i = v % p;
            if (!SVALID(s)) continue;
            R* sl = SLOT(s);
            const R* Cr = a.C + ((size_t)SB(s) * (T + 1) + t) * szC + (size_t)i * p;
            const R mi = sl[L.oq + i];
            R acc = R(0);
            for (int j = 0; j < p; ++j) acc = fma(ldg(Cr + j), fma(mi, sl[L.oq + j], sl[L.oM + j * pp + i]), acc);
            acc = fma(ldg(Cr + i), reg, acc);
            sl[L.orow + i] = acc / 2 + mi * ldg(a.c + ((size_t)SB(s) * (T + 1) + t) * p + i);
          } else if (idx < o4) {
            // delta_u = (K_prev - K) mu_x + k_prev - k   (lqg.py:294)
            const int v = idx - o3, s = v / dU, ai = v % dU;
            if (!SVALID(s)) continue;
            R* sl = SLOT(s);
            const size_t bt = (size_t)SB(s) * T + t;
            const R* pK = a.prevK + bt * szK + (size_t)ai * dX;
            R acc = R(0);
            for (int k = 0; k < dX; ++k) acc = fma(ldg(pK + k) - sl[L.oK + ai * dXp + k], sl[L.oq + k], acc);
            sl[L.odelta + ai] = acc + ldg(a.prevk + bt * dU + ai) - sl[L.okb + ai];
          } else if (idx < o5) {
            const int v = idx - o4, s = v / (p * p), r = v % (p * p), i = r / p, j = r % p;
            if (!SVALID(s)) continue;
            const R* sl = SLOT(s);
            a.tcov[(PIDX(s) * (T + 1) + t) * szC + r] = sl[L.oM + i * pp + j] + (i == j ? reg : R(0));
          } else {
            const int v = idx - o5, s = v / p, i = v % p;
            if (!SVALID(s)) continue;
            a.tmean[(PIDX(s) * (T + 1) + t) * p + i] = SLOT(s)[L.oq + i];
          }
        }
      }
      ctx.sync();

      // ---- F4: Sigma'_xx = (F Sigma) F^T + dyn_covar (upper tiles, mirrored) (lqg.py:236) -----
      {
        const int ntri = ntx * (ntx + 1) / 2;
        const int o1 = S * ntri, o2 = o1 + S * dX, o3 = o2 + S;
        TEAM_FOR(idx, o3) {
          if (idx < o1) {
            const int s = idx / ntri;
            if (!SVALID(s)) continue;
            int ti, tj;
            tri_tile(idx % ntri, ntx, ti, tj);
            R* sl = SLOT(s);
            const R* Ft = conds + (size_t)(s / ES) * L.cond + L.oFs;
            R acc[4][4];
            tile4x4(sl + L.oR, dXp, 4 * ti, Ft, dXp, 4 * tj, p, acc);
            const R* D = a.dyn_covar + ((size_t)SB(s) * T + t) * dX * dX;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const int i = 4 * ti + r, j = 4 * tj + c;
                if (i <= j && j < dX) {
                  const R val = acc[r][c] + ldg(D + (size_t)i * dX + j);
                  sl[L.oM + i * pp + j] = val;
                  sl[L.oM + j * pp + i] = val;
                }
              }
            }
          } else if (idx < o2) {
            const int v = idx - o1, s = v / dX, i = v % dX;
            if (!SVALID(s)) continue;
            R* sl = SLOT(s);
            sl[L.oq + i] = sl[L.omun + i];
          } else {
            const int s = idx - o2;
            if (!SVALID(s)) continue;
            R* sl = SLOT(s);
            const size_t bt = (size_t)SB(s) * T + t;
            if (do_cost) {
              R acc = R(0);
              for (int i = 0; i < p; ++i) acc += sl[L.orow + i];
              sl[L.oscal + 2] += acc + ldg(a.cc + (size_t)SB(s) * (T + 1) + t);
            }
            if (do_kl) {
              // 1/2 [ tr(Lam_prev Sigma_new) + delta^T Lam_prev delta - dU ] + log|chol prev|   (lqg.py:295-304)
              const R* Lam = a.prevLam + bt * szS;
              R tr = R(0), quad = R(0);
              for (int i = 0; i < dU; ++i) {
                R row = R(0);
                for (int j = 0; j < dU; ++j) {
                  const R lij = ldg(Lam + i * dU + j);
                  tr = fma(lij, sl[L.oSinv + j * dU + i], tr);
                  row = fma(lij, sl[L.odelta + j], row);
                }
                quad = fma(sl[L.odelta + i], row, quad);
              }
              sl[L.oscal + 1] += (tr + quad - R(dU)) / 2 + ldg(a.prev_hld + bt);
            }
          }
        }
      }
      ctx.sync();
    }

    // ---- final row T: state block only (lqg.py:221-222,242) ---------------------------------
    {
      const int n_cost = do_cost ? S * dX : 0, n_out = a.tcov ? S * p * p : 0, n_outm = a.tmean ? S * p : 0;
      const int o1 = n_cost, o2 = o1 + n_out, o3 = o2 + n_outm;
      TEAM_FOR(idx, o3) {
        if (idx < o1) {
          const int s = idx / dX, i = idx % dX;
          if (!SVALID(s)) continue;
          R* sl = SLOT(s);
          const R* Cr = a.C + ((size_t)SB(s) * (T + 1) + T) * szC + (size_t)i * p;
          const R mi = sl[L.oq + i];
          R acc = R(0);
          for (int j = 0; j < dX; ++j) acc = fma(ldg(Cr + j), fma(mi, sl[L.oq + j], sl[L.oM + j * pp + i]), acc);
          sl[L.orow + i] = acc / 2 + mi * ldg(a.c + ((size_t)SB(s) * (T + 1) + T) * p + i);
        } else if (idx < o2) {
          const int v = idx - o1, s = v / (p * p), r = v % (p * p), i = r / p, j = r % p;
          if (!SVALID(s)) continue;
          a.tcov[(PIDX(s) * (T + 1) + T) * szC + r] = (i < dX && j < dX) ? SLOT(s)[L.oM + i * pp + j] : R(0);
        } else {
          const int v = idx - o2, s = v / p, i = v % p;
          if (!SVALID(s)) continue;
          a.tmean[(PIDX(s) * (T + 1) + T) * p + i] = i < dX ? SLOT(s)[L.oq + i] : R(0);
        }
      }
      ctx.sync();
    }
  }

  // ---- results --------------------------------------------------------------------------------
  TEAM_FOR(s, S) {
    if (!SVALID(s)) continue;
    R* sl = SLOT(s);
    if ((a.flags & ILQR_FORWARD) && (a.flags & ILQR_COST) && a.cost) {
      R acc = R(0);
      for (int i = 0; i < dX; ++i) acc += sl[L.orow + i];
      a.cost[PIDX(s)] = sl[L.oscal + 2] + acc + ldg(a.cc + (size_t)SB(s) * (T + 1) + T);
    }
    if ((a.flags & ILQR_FORWARD) && (a.flags & ILQR_KL) && a.kl) a.kl[PIDX(s)] = (sl[L.oscal + 1] + sl[L.oscal + 3]) / R(T);
    if (a.info) a.info[PIDX(s)] = meta[4 * s + 3];
  }
#undef SLOT
#undef SB
#undef SE
#undef SVALID
#undef PIDX
}

// ---------------------------------------------------------------------------------------------
// Small per-(b,t) helpers around the step kernel (one thread per matrix; not hot).
// ---------------------------------------------------------------------------------------------

// extended_costs_kl (lqg.py:247-273): C_kl = [[K^T L K, -K^T L],[-L K, L]], c_kl = [K^T L k; -L k].
// One item per (b,t,row i of the p x p block).
template <typename R>
DONK_HD void extended_costs_item(size_t bt, int i, int dX, int dU, const R* K, const R* k, const R* Lam, R* C, R* c) {
  const int p = dX + dU;
  const R* Kt = K + bt * dU * dX;
  const R* kt = k + bt * dU;
  const R* Lt = Lam + bt * dU * dU;
  R* Ct = C + bt * p * p + (size_t)i * p;
  if (i < dX) {
    // row i of K^T Lam: g[b] = sum_a K[a][i] Lam[a][b]
    R ci = R(0);
    for (int j = 0; j < dX; ++j) Ct[j] = R(0);
    for (int b = 0; b < dU; ++b) {
      R g = R(0);
      for (int a2 = 0; a2 < dU; ++a2) g = fma(Kt[a2 * dX + i], Lt[a2 * dU + b], g);
      Ct[dX + b] = -g;
      for (int j = 0; j < dX; ++j) Ct[j] = fma(g, Kt[b * dX + j], Ct[j]);
      ci = fma(g, kt[b], ci);
    }
    c[bt * p + i] = ci;
  } else {
    const int a2 = i - dX;
    R ci = R(0);
    for (int j = 0; j < dX; ++j) {
      R g = R(0);
      for (int b = 0; b < dU; ++b) g = fma(Lt[a2 * dU + b], Kt[b * dX + j], g);
      Ct[j] = -g;
    }
    for (int b = 0; b < dU; ++b) {
      Ct[dX + b] = Lt[a2 * dU + b];
      ci = fma(Lt[a2 * dU + b], kt[b], ci);
    }
    c[bt * p + i] = -ci;
  }
}

// extended_costs_kl for a group of (b,t) pairs per CTA: inputs staged in shared memory, outputs written with
// consecutive threads on consecutive addresses (the one-thread-per-row version above wrote with stride p and took as
// long as the whole iLQR step at config 3).  shared per pair: K (dU*dX) | k (dU) | Lam (dU*dU) | G = Lam K (dU*dX) | g = Lam k (dU)
DONK_HD size_t extended_costs_smem(int dX, int dU) { return (size_t)2 * dU * dX + 2 * dU + (size_t)dU * dU; }
template <typename R>
DONK_D void extended_costs_cta(const Ctx& ctx, int n, int npairs, int cta, int dX, int dU, const R* K, const R* k,
                               const R* Lam, R* C, R* c, R* sm) {
  const int p = dX + dU, per = (int)extended_costs_smem(dX, dU);
  const size_t first = (size_t)cta * npairs;
  TEAM_FOR(idx, npairs * per) {
    const int q = idx / per, r = idx % per;
    const size_t bt = first + q;
    if (bt >= (size_t)n) continue;
    R v = R(0);
    if (r < dU * dX) v = ldg(K + bt * dU * dX + r);
    else if (r < dU * dX + dU) v = ldg(k + bt * dU + (r - dU * dX));
    else if (r < dU * dX + dU + dU * dU) v = ldg(Lam + bt * dU * dU + (r - dU * dX - dU));
    sm[idx] = v;
  }
  ctx.sync();
  TEAM_FOR(idx, npairs * dU * (dX + 1)) {  // G = Lam K, g = Lam k
    const int q = idx / (dU * (dX + 1)), r = idx % (dU * (dX + 1)), a2 = r / (dX + 1), j = r % (dX + 1);
    if (first + q >= (size_t)n) continue;
    R* s = sm + (size_t)q * per;
    const R* Ks = s;
    const R* ks = s + dU * dX;
    const R* Ls = ks + dU;
    R acc = R(0);
    for (int b2 = 0; b2 < dU; ++b2) acc = fma(Ls[a2 * dU + b2], j < dX ? Ks[b2 * dX + j] : ks[b2], acc);
    if (j < dX) s[dU * dX + dU + dU * dU + a2 * dX + j] = acc;
    else s[2 * dU * dX + dU + dU * dU + a2] = acc;
  }
  ctx.sync();
  TEAM_FOR(idx, npairs * (p * p + p)) {
    const int q = idx / (p * p + p), r = idx % (p * p + p);
    const size_t bt = first + q;
    if (bt >= (size_t)n) continue;
    const R* s = sm + (size_t)q * per;
    const R* Ks = s;
    const R* Ls = s + dU * dX + dU;
    const R* Gs = Ls + dU * dU;
    const R* gs = Gs + dU * dX;
    if (r < p * p) {
      const int i = r / p, j = r % p;
      R v;
      if (i < dX && j < dX) {  // K^T Lam K
        v = R(0);
        for (int a2 = 0; a2 < dU; ++a2) v = fma(Ks[a2 * dX + i], Gs[a2 * dX + j], v);
      } else if (i < dX) {  // -K^T Lam = -(Lam K)^T
        v = -Gs[(j - dX) * dX + i];
      } else if (j < dX) {  // -Lam K
        v = -Gs[(i - dX) * dX + j];
      } else {
        v = Ls[(i - dX) * dU + (j - dX)];
      }
      C[bt * p * p + r] = v;
    } else {
      const int i = r - p * p;
      R v;
      if (i < dX) {  // K^T Lam k
        v = R(0);
        for (int a2 = 0; a2 < dU; ++a2) v = fma(Ks[a2 * dX + i], gs[a2], v);
      } else {
        v = -gs[i - dX];
      }
      c[bt * p + i] = v;
    }
  }
}

// Cholesky-based inverse and half log-determinant of one small SPD matrix held in global memory
// (TimeVaryingLinearGaussian lazy chol_covar / inv_covar, donk/models/tvlg.py:62-99,
// donk/utils/batched.py:6-49).  ws: 2*d*d scratch.  Returns 0 or 1 (not PD).
template <typename R>
DONK_HD int spd_inverse_item(int d, const R* A, R* Ainv, R* chol, R* half_logdet, R* ws) {
  R* Lm = ws;
  R* Li = ws + d * d;
  int bad = 0;
  R hld = R(0);
  for (int j = 0; j < d; ++j) {
    R sj = A[j * d + j];
    for (int k = 0; k < j; ++k) sj = fma(-Lm[j * d + k], Lm[j * d + k], sj);
    if (!(sj > R(0))) bad = 1;
    const R dj = sqrt(sj);
    Lm[j * d + j] = dj;
    hld += log(dj);
    for (int i = j + 1; i < d; ++i) {
      R si = A[i * d + j];
      for (int k = 0; k < j; ++k) si = fma(-Lm[i * d + k], Lm[j * d + k], si);
      Lm[i * d + j] = si / dj;
      Lm[j * d + i] = R(0);
    }
  }
  if (chol)
    for (int i = 0; i < d * d; ++i) chol[i] = Lm[i];
  if (half_logdet) *half_logdet = hld;
  if (Ainv) {
    for (int c = 0; c < d; ++c) {
      for (int i = 0; i < c; ++i) Li[i * d + c] = R(0);
      Li[c * d + c] = R(1) / Lm[c * d + c];
      for (int i = c + 1; i < d; ++i) {
        R acc = R(0);
        for (int k = c; k < i; ++k) acc = fma(Lm[i * d + k], Li[k * d + c], acc);
        Li[i * d + c] = -acc / Lm[i * d + i];
      }
    }
    for (int i = 0; i < d; ++i)
      for (int j = 0; j <= i; ++j) {
        R acc = R(0);
        for (int k = i; k < d; ++k) acc = fma(Li[k * d + i], Li[k * d + j], acc);
        Ainv[i * d + j] = acc;
        Ainv[j * d + i] = acc;
      }
  }
  return bad;
}

// kl_divergence_action for an arbitrary state sequence X (lqg.py:276-306): term of one (b,t).
// hld_* = sum log diag chol(covar).  Returns 1/2[tr + quad - dU] + hld_prev - hld_new.
template <typename R>
DONK_HD R kl_term_item(int dX, int dU, const R* x, const R* K, const R* k, const R* covar, R hld_new, const R* pK,
                       const R* pk, const R* pLam, R hld_prev, R* delta) {
  for (int a2 = 0; a2 < dU; ++a2) {
    R acc = R(0);
    for (int j = 0; j < dX; ++j) acc = fma(pK[a2 * dX + j] - K[a2 * dX + j], x[j], acc);
    delta[a2] = acc + pk[a2] - k[a2];
  }
  R tr = R(0), quad = R(0);
  for (int i = 0; i < dU; ++i) {
    R row = R(0);
    for (int j = 0; j < dU; ++j) {
      tr = fma(pLam[i * dU + j], covar[j * dU + i], tr);
      row = fma(pLam[i * dU + j], delta[j], row);
    }
    quad = fma(delta[i], row, quad);
  }
  return (tr + quad - R(dU)) / 2 + hld_prev - hld_new;
}

// Lock-step Brent root bracketing on g(x) = kl(step(exp x)) - kl_step, one state machine per
// condition (ILQR.optimize, lqg.py:116-123; scipy.optimize.brentq restated in SURVEY.md 8(a-14)).
// state (B,12): xpre xcur xblk fpre fcur fblk spre scur | status iter  (status: 0 running,
// 1 converged, 2 sign error, 3 maxiter).  Called once per evaluation round with the fresh kl.
template <typename R>
struct BrentState {
  R xpre, xcur, xblk, fpre, fcur, fblk, spre, scur;
  R status, iter, root, pad;
};

template <typename R>
DONK_HD R sgn(R x) { return x > R(0) ? R(1) : (x < R(0) ? R(-1) : R(0)); }

// phase 0: initialise with f(xa), f(xb); later phases: consume f(xcur).  Writes next eta.
template <typename R>
DONK_HD void brent_item(BrentState<R>& st, int phase, R xa, R xb, R fa, R fb, R fnew, R xtol, R rtol, int maxiter,
                        R* eta_out, int* active_out) {
  if (phase == 0) {
    st.xpre = xa; st.xcur = xb; st.fpre = fa; st.fcur = fb;
    st.xblk = R(0); st.fblk = R(0); st.spre = R(0); st.scur = R(0);
    st.iter = R(0); st.status = R(0); st.root = R(0);
    if (fa == R(0)) { st.status = R(1); st.root = xa; }
    else if (fb == R(0)) { st.status = R(1); st.root = xb; }
    else if (sgn(fa) == sgn(fb)) { st.status = R(2); }
  } else {
    if (st.status != R(0)) return;
    st.fcur = fnew;
  }
  if (st.status == R(0)) {
    if (st.iter >= R(maxiter)) { st.status = R(3); }
    else {
      st.iter += R(1);
      if (st.fpre != R(0) && st.fcur != R(0) && sgn(st.fpre) != sgn(st.fcur)) {
        st.xblk = st.xpre; st.fblk = st.fpre;
        st.spre = st.scur = st.xcur - st.xpre;
      }
      if (fabs(st.fblk) < fabs(st.fcur)) {
        st.xpre = st.xcur; st.xcur = st.xblk; st.xblk = st.xpre;
        st.fpre = st.fcur; st.fcur = st.fblk; st.fblk = st.fpre;
      }
      const R delta = (xtol + rtol * fabs(st.xcur)) / 2;
      const R sbis = (st.xblk - st.xcur) / 2;
      if (st.fcur == R(0) || fabs(sbis) < delta) {
        st.status = R(1); st.root = st.xcur;
      } else {
        if (fabs(st.spre) > delta && fabs(st.fcur) < fabs(st.fpre)) {
          R stry;
          if (st.xpre == st.xblk) {
            stry = -st.fcur * (st.xcur - st.xpre) / (st.fcur - st.fpre);
          } else {
            const R dpre = (st.fpre - st.fcur) / (st.xpre - st.xcur);
            const R dblk = (st.fblk - st.fcur) / (st.xblk - st.xcur);
            stry = -st.fcur * (st.fblk * dblk - st.fpre * dpre) / (dblk * dpre * (st.fblk - st.fpre));
          }
          const R lim1 = fabs(st.spre), lim2 = 3 * fabs(sbis) - delta;
          if (2 * fabs(stry) < (lim1 < lim2 ? lim1 : lim2)) { st.spre = st.scur; st.scur = stry; }
          else { st.spre = sbis; st.scur = sbis; }
        } else {
          st.spre = sbis; st.scur = sbis;
        }
        st.xpre = st.xcur; st.fpre = st.fcur;
        if (fabs(st.scur) > delta) st.xcur += st.scur;
        else st.xcur += (sbis > R(0) ? delta : -delta);
      }
    }
  }
  if (st.status == R(0)) { *eta_out = exp(st.xcur); *active_out = 1; }
  else { if (st.status == R(1)) *eta_out = exp(st.root); *active_out = 0; }
}

}  // namespace donk

// ---------------------------------------------------------------------------------------------
// Launch planning (host side, shared by the CUDA library and the emulation build).
// ---------------------------------------------------------------------------------------------
namespace donk {

struct IlqrPlan {
  int S, ES, NC, threads, grid;
  size_t smem;
};

// Choose slots per CTA.  Prefer a footprint that lets two CTAs share an SM (their sync stalls
// overlap); fall back to the full 227 KB for large state dimensions.  force_S/force_threads > 0
// override the heuristic (tuning knobs DONK_ILQR_SLOTS / DONK_ILQR_THREADS).
inline int ilqr_plan(int B, int E, int dX, int dU, size_t sz, size_t smem_cap, int force_S, int force_threads,
                     IlqrPlan& pl) {
  const IlqrLayout L(dX, dU);
  auto shape = [&](int S, int& ES, int& NC) {
    if (E >= S) { ES = S; NC = 1; }
    else { ES = E; NC = S / E; }
    return NC * ES;
  };
  auto fits = [&](int S, size_t cap) {
    int ES, NC;
    const int S2 = shape(S, ES, NC);
    return S2 >= 1 && L.bytes(S2, NC, sz) <= cap;
  };
  const size_t half = (smem_cap + 1024) / 2 - 1024;  // two CTAs per SM: 228 KB total, 1 KB reserved each
  int S = 0;
  if (force_S > 0) {
    S = force_S;
    if (!fits(S, smem_cap)) return DONK_SMEM_LIMIT;
  } else {
    for (int cand = 8; cand >= 4 && !S; --cand)
      if (fits(cand, half)) S = cand;
    if (S == 8) {  // small problems: more slots per CTA while staying under a quarter of the SM
      for (int cand = 16; cand > 8; --cand)
        if (fits(cand, half / 2)) { S = cand; break; }
    }
    for (int cand = 16; cand >= 1 && !S; --cand)
      if (fits(cand, smem_cap)) S = cand;
    if (!S) return DONK_SMEM_LIMIT;
  }
  S = shape(S, pl.ES, pl.NC);
  pl.S = S;
  pl.smem = L.bytes(S, pl.NC, sz);
  const int n_ech = (E + pl.ES - 1) / pl.ES;
  pl.grid = ((B + pl.NC - 1) / pl.NC) * n_ech;
  int items = S * (L.dXp / 4) * (L.pp / 4);
  int th = round_up(items, 32);
  if (th < 64) th = 64;
  if (th > 512) th = 512;
  pl.threads = force_threads > 0 ? force_threads : th;
  return DONK_OK;
}

}  // namespace donk
