// donk_b200 — iLQR step, warp-per-problem variant with compile-time state/action dimensions.
//
// Same math and reference lines as ilqr.cuh (ILQR.step, donk/traj_opt/lqg.py:55-75; backward :126-182;
// forward :185-244; kl_divergence_action :276-306; expected_costs quadratic_costs.py:106-119); this is the
// fast path for small/medium dimensions (a problem fits one warp), selected by the launcher when an
// instantiation for (dX, dU) exists.  What changed against the generic kernel, and why (ncu, profiles/r01):
// only 14 % of the generic kernel's instructions were tile DFMAs, the rest runtime index arithmetic,
// scalar dot-product loops and global loads inside the phases.  Here
//   * a CTA is G warps = the eta candidates of one condition (or G conditions when E = 1); each warp owns one
//     (condition, eta) problem and synchronises with __syncwarp only;
//   * dimensions are template parameters: tiles are lane-static, all offsets are immediates, no div/mod;
//   * the vectors ride in the padding column P of the matrices: F_ext = [F | f], M[:,P] = q / v / mu, so
//     w = V f + v, q = c + F^T w, mu_u = K mu_x + k and F mu come out of the matrix tiles for free;
//   * the per-condition inputs of step t (F_t, f_t, C_t, C_kl_t, ... ) are staged once per CTA with cp.async,
//     overlapped with the phases that do not read them (Cholesky / gains / value update);
//   * Quu (dU x dU) is factored redundantly in every lane's registers (no shared-memory round trips, no
//     syncs); divisions are replaced by rsqrt / reciprocal multiplies, C/eta by C * (1/eta) (<= 1 ulp).
#pragma once
#include "common.cuh"
#include "ilqr.cuh"

namespace donk {

template <int DX, int DU>
struct WarpDims {
  static constexpr int P = DX + DU, PE = P + 1;
  static constexpr int PP = (PE + 3) / 4 * 4;    // leading dimension of M, W, F_ext (even, 16-byte rows)
  static constexpr int DXP = (DX + 3) / 4 * 4;   // leading dimension of Rt, F^T
  static constexpr int DUP = (DU + 3) / 4 * 4;
  static constexpr int KLD = (DX + 1 + 3) / 4 * 4;  // Kb rows: [K_a | k_a | 0]
  // B1: W_ext (DX x PP) in TM1 x 4 tiles, one per lane if possible
  static constexpr int NTJ1 = PP / 4;
  static constexpr int CAP1 = 32 / NTJ1 > 0 ? 32 / NTJ1 : 1;
  static constexpr int TM1 = (DX + CAP1 - 1) / CAP1;
  static constexpr int NTI1 = (DX + TM1 - 1) / TM1;
  // B2: Q_ext upper triangle in 4 x 4 tiles
  static constexpr int NTP = PP / 4, NTRI2 = NTP * (NTP + 1) / 2;
  // F3: Rt_ext (PE x DX) in 4 x TN3 tiles
  static constexpr int TN3 = TM1, NTI3 = NTI1;
  // B5 / F4: DX x DX upper triangle in 2 x 4 tiles
  static constexpr int NR5 = (DX + 1) / 2, NC5 = DXP / 4;
  // team area (reals)
  static constexpr int RSZ = DX * PP > PE * DXP ? DX * PP : PE * DXP;
  static constexpr int KSZ = DU * KLD > DX * DUP ? DU * KLD : DX * DUP;
  static constexpr int oM = 0, oR = oM + PE * PP, oK = oR + (RSZ + 3) / 4 * 4, oKv = oK + (KSZ + 3) / 4 * 4;
  static constexpr int oPc = oKv + DUP, oLds = oPc + (DU * DU + 3) / 4 * 4, oDel = oLds + DUP, oRow = oDel + DUP;
  static constexpr int oScal = oRow + PP, TEAM = oScal + 8;  // scal: 0 inv_eta 1 kl 2 cost 3 logdet 4 notpd 5 eta
  // stage area per condition (reals): backward and forward views share the storage
  static constexpr int oFx = 0;                                   // bwd: F_ext (DX x PP); fwd: F^T (P x DXP)
  static constexpr int FXS = DX * PP > P * DXP ? DX * PP : P * DXP;
  static constexpr int oCs = oFx + (FXS + 3) / 4 * 4;             // C_t (P*P)
  static constexpr int oCk = oCs + (P * P + 3) / 4 * 4;           // bwd: C_kl_t (P*P); fwd: dyn_covar_t (DX*DX) | prevK | prevLam
  static constexpr int CKS = P * P > DX * DX + DU * DX + DU * DU ? P * P : DX * DX + DU * DX + DU * DU;
  static constexpr int oVec = oCk + (CKS + 3) / 4 * 4;            // cs (P) | ck (P) / fwd: cs (P) | fv (DX) | pk (DU) | cc, phld
  static constexpr int STAGE = oVec + 2 * PP + DXP + DUP + 4;
};

template <int TM, int NR, int NC>
DONK_HD bool tri2_tile(int idx, int& ti, int& tj) {
  // enumerate (ti, tj) tiles of TM x 4 blocks that touch the upper triangle (4 tj + 3 >= TM ti)
  for (int r = 0; r < NR; ++r) {
    const int first = (TM * r) / 4;  // first column tile with 4 tj + 3 >= TM r
    const int cnt = NC - first;
    if (idx < cnt) { ti = r; tj = first + idx; return true; }
    idx -= cnt;
  }
  return false;
}
template <int TM, int NR, int NC>
DONK_HD constexpr int tri2_count() {
  int n = 0;
  for (int r = 0; r < NR; ++r) n += NC - (TM * r) / 4;
  return n;
}

template <typename R, int DX, int DU>
DONK_D void ilqr_warp_cta(const WCtx& w, const IlqrArgs<R>& a, int cta, int ES, R* smem) {
  using D = WarpDims<DX, DU>;
  constexpr int P = D::P, PE = D::PE, PP = D::PP, DXP = D::DXP, DUP = D::DUP, KLD = D::KLD;
  const int G = w.G, T = a.T, E = a.E;
  const int NC = G / ES, n_ech = (E + ES - 1) / ES;
  const int cg = cta / n_ech, ec = cta % n_ech;
  R* teams = smem;
  R* stages = smem + (size_t)G * D::TEAM;
  const R gamma = a.gamma;
  constexpr size_t szF = (size_t)DX * P, szC = (size_t)P * P, szK = (size_t)DU * DX, szS = (size_t)DU * DU;

  // team -> problem (uniform arithmetic, no shared state needed)
  auto team_b = [&](int tm) { return cg * NC + tm / ES; };
  auto team_e = [&](int tm) { return ec * ES + tm % ES; };
  auto team_valid = [&](int tm) {
    const int b = team_b(tm), e = team_e(tm);
    return b < a.B && e < E && (a.active == nullptr || a.active[b] != 0);
  };
  auto cond_b = [&](int ci) {  // condition ci of this CTA, or -1
    const int b = cg * NC + ci;
    return (b < a.B && ec * ES < E && (a.active == nullptr || a.active[b] != 0)) ? b : -1;
  };
  {
    bool any = false;
    for (int ci = 0; ci < NC; ++ci) any = any || cond_b(ci) >= 0;
    if (!any) return;  // CTA-uniform
  }
#define TEAMP(tm) (teams + (size_t)(tm) * D::TEAM)
#define STAGEP(tm) (stages + (size_t)((tm) / ES) * D::STAGE)
#define PIDX(tm) ((size_t)team_b(tm) * E + team_e(tm))

  TEAMS_DO(tm) {
    if (!team_valid(tm)) continue;
    R* tp = TEAMP(tm);
    LANE_FOR(i, 8) {
      const R eta = a.eta ? a.eta[PIDX(tm)] : R(1);
      tp[D::oScal + i] = i == 0 ? R(1) / eta : (i == 5 ? eta : R(0));
    }
  }

  // =========================================================================================
  // Backward pass
  // =========================================================================================
  if (a.flags & ILQR_BACKWARD) {
    // ---- async staging of step t: F_ext = [F_t | f_t | 0], C_t, C_kl_t, c_t, c_kl_t ------------------
    auto stage_bwd = [&](int t) {
      for (int ci = 0; ci < NC; ++ci) {
        const int b = cond_b(ci);
        if (b < 0) continue;
        R* st = stages + (size_t)ci * D::STAGE;
        const size_t bt = (size_t)b * T + t, bt1 = (size_t)b * (T + 1) + t;
        CTA_FOR(idx, DX * PP) {
          const int k = idx / PP, j = idx % PP;  // compile-time divisors
          if (j < P) cp_async(st + D::oFx + idx, a.F + bt * szF + (size_t)k * P + j);
          else if (j == P) cp_async(st + D::oFx + idx, a.f + bt * DX + k);
          else st[D::oFx + idx] = R(0);
        }
        CTA_FOR(idx, P * P) {
          cp_async(st + D::oCs + idx, a.C + bt1 * szC + idx);
          if (a.C_kl) cp_async(st + D::oCk + idx, a.C_kl + bt * szC + idx);
          else st[D::oCk + idx] = R(0);
        }
        CTA_FOR(i, P) {
          cp_async(st + D::oVec + i, a.c + bt1 * P + i);
          if (a.c_kl) cp_async(st + D::oVec + PP + i, a.c_kl + bt * P + i);
          else st[D::oVec + PP + i] = R(0);
        }
      }
    };
    stage_bwd(T - 1);
    // terminal value V = C_ext[T,:DX,:DX], v = c_ext[T,:DX] (lqg.py:151-152)
    TEAMS_DO(tm) {
      if (!team_valid(tm)) continue;
      R* tp = TEAMP(tm);
      const R ie = a.eta ? R(1) / a.eta[PIDX(tm)] : R(1);
      const size_t bT = (size_t)team_b(tm) * (T + 1) + T;
      LANE_FOR(idx, DX * DX) {
        const int i = idx / DX, j = idx % DX;
        tp[D::oM + i * PP + j] = ldg(a.C + bT * szC + (size_t)i * P + j) * ie;
      }
      LANE_FOR(i, DX) tp[D::oM + i * PP + P] = ldg(a.c + bT * P + i) * ie;
    }
    cp_async_wait_all();
    w.cta_sync();

    for (int t = T - 1; t >= 0; --t) {
      TEAMS_DO(tm) {
        if (!team_valid(tm)) continue;
        R* tp = TEAMP(tm);
        const R* st = STAGEP(tm);
        // ---- B1: W_ext = V [F | f] (+ v in column P) ------------------------------------------------
        LANE_FOR(it, D::NTI1 * D::NTJ1) {
          const int ti = it / D::NTJ1, tj = it % D::NTJ1;
          R acc[D::TM1][4];
          tile_km<R, D::TM1, 4, D::TM1 == 4, true>(tp + D::oM + D::TM1 * ti, PP, st + D::oFx + 4 * tj, PP, DX, acc);
#pragma unroll
          for (int r = 0; r < D::TM1; ++r) {
            const int i = D::TM1 * ti + r;
            if (i < DX) {
              if (4 * tj <= P && P < 4 * tj + 4) acc[r][P - 4 * tj] += tp[D::oM + i * PP + P];  // w = V f + v
#pragma unroll
              for (int c = 0; c < 4; ++c) tp[D::oR + i * PP + 4 * tj + c] = acc[r][c];
            }
          }
        }
        w.team_sync();
        // ---- B2: Q_ext = C_ext + gamma [F|f]^T W_ext, upper tiles, mirrored; column P is q ---------------
        {
          const R ie = tp[D::oScal + 0];
          LANE_FOR(it, D::NTRI2) {
            int ti, tj;
            tri_tile(it, D::NTP, ti, tj);
            R acc[4][4];
            tile_km<R, 4, 4, true, true>(st + D::oFx + 4 * ti, PP, tp + D::oR + 4 * tj, PP, DX, acc);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const int i = 4 * ti + r, j = 4 * tj + c;
                if (i <= j && i < P) {
                  if (j < P) {
                    const R val = fma(st[D::oCs + i * P + j], ie, st[D::oCk + i * P + j]) + gamma * acc[r][c];
                    tp[D::oM + i * PP + j] = val;
                    tp[D::oM + j * PP + i] = val;
                  } else if (j == P) {
                    tp[D::oM + i * PP + P] = fma(st[D::oVec + i], ie, st[D::oVec + PP + i]) + gamma * acc[r][c];
                  }
                }
              }
            }
          }
        }
        w.team_sync();
      }
      // all teams are done with the staged inputs of step t: start fetching step t-1 while the
      // factorisation, gains and value update (which only touch team-private data) run
      w.cta_sync();
      if (t > 0) stage_bwd(t - 1);
      TEAMS_DO(tm) {
        if (!team_valid(tm)) continue;
        R* tp = TEAMP(tm);
        const size_t pt = PIDX(tm) * T + t;
        // ---- B34: Quu = L L^T in registers (every lane, redundantly); K = -Quu^-1 Qux; k = -Quu^-1 q_u ----
        LANE_FOR(ln, 32) {
          R Lm[DU][DU];  // lower triangle used
#pragma unroll
          for (int i = 0; i < DU; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) Lm[i][j] = tp[D::oM + (DX + i) * PP + DX + j];
          R idg[DU];
          bool bad = false;
#pragma unroll
          for (int j = 0; j < DU; ++j) {
            R sj = Lm[j][j];
#pragma unroll
            for (int k = 0; k < j; ++k) sj = fma(-Lm[j][k], Lm[j][k], sj);
            bad = bad || !(sj > R(0));
            const R rs = rsqrt_r(sj);
            idg[j] = rs;           // 1 / L_jj
            Lm[j][j] = sj * rs;    // L_jj
#pragma unroll
            for (int i = j + 1; i < DU; ++i) {
              R si = Lm[i][j];
#pragma unroll
              for (int k = 0; k < j; ++k) si = fma(-Lm[i][k], Lm[j][k], si);
              Lm[i][j] = si * rs;
            }
          }
          if (ln < DU) {  // log diag chol(Quu): one value per lane (compile-time selection, no dynamic indexing)
            R dsel = Lm[0][0];
#pragma unroll
            for (int j = 1; j < DU; ++j) dsel = ln == j ? Lm[j][j] : dsel;
            tp[D::oLds + ln] = log(dsel);
          }
          if (ln == 31 && bad && tp[D::oScal + 4] == R(0)) tp[D::oScal + 4] = R(t + 1);
          // in-place inverse of the lower factor: Li = L^-1
#pragma unroll
          for (int c = 0; c < DU; ++c) {
            Lm[c][c] = idg[c];
#pragma unroll
            for (int i = c + 1; i < DU; ++i) {
              R acc = R(0);
#pragma unroll
              for (int k = c; k < i; ++k) acc = fma(Lm[i][k], Lm[k][c], acc);
              Lm[i][c] = -acc * idg[i];
            }
          }
          // pol_covar = Quu^-1 = Li^T Li (symmetric by construction, lqg.py:165-166); lane e stores entry e
          {
            int e = 0;
#pragma unroll
            for (int i = 0; i < DU; ++i)
#pragma unroll
              for (int j = 0; j <= i; ++j, ++e) {
                R acc = R(0);
#pragma unroll
                for (int k = i; k < DU; ++k) acc = fma(Lm[k][i], Lm[k][j], acc);
                if (ln == (e & 31)) {
                  a.pcov[pt * szS + i * DU + j] = acc;
                  a.pcov[pt * szS + j * DU + i] = acc;
                  const R quu = tp[D::oM + (DX + i) * PP + DX + j];
                  a.pinv[pt * szS + i * DU + j] = quu;
                  a.pinv[pt * szS + j * DU + i] = quu;
                }
              }
          }
          // column ln of K (ln < DX), or k (ln == DX): z = -Li^T (Li x), x = Qux[:, ln] / q_u
          if (ln <= DX) {
            const int col = ln < DX ? ln : P;
            R x[DU];
#pragma unroll
            for (int b2 = 0; b2 < DU; ++b2) x[b2] = tp[D::oM + (DX + b2) * PP + col];
#pragma unroll
            for (int i = DU - 1; i >= 0; --i) {  // y = Li x (in place, bottom up)
              R acc = R(0);
#pragma unroll
              for (int k = 0; k <= i; ++k) acc = fma(Lm[i][k], x[k], acc);
              x[i] = acc;
            }
#pragma unroll
            for (int i = 0; i < DU; ++i) {  // z = Li^T y (in place, top down)
              R acc = R(0);
#pragma unroll
              for (int k = i; k < DU; ++k) acc = fma(Lm[k][i], x[k], acc);
              x[i] = -acc;
            }
#pragma unroll
            for (int a2 = 0; a2 < DU; ++a2) {
              tp[D::oK + a2 * KLD + ln] = x[a2];
              if (ln < DX) a.K[pt * szK + a2 * DX + ln] = x[a2];
              else a.k[pt * DU + a2] = x[a2];
            }
          }
        }
        w.team_sync();
        // ---- B5: V = Qxx + Qxu K (upper, mirrored); v = q_x + Qxu k (lqg.py:178-180) ---------------------
        LANE_FOR(it, (tri2_count<2, D::NR5, D::NC5>())) {
          int ti, tj;
          tri2_tile<2, D::NR5, D::NC5>(it, ti, tj);
          R acc[2][4];
          tile_km<R, 2, 4, false, true>(tp + D::oM + DX * PP + 2 * ti, PP, tp + D::oK + 4 * tj, KLD, DU, acc);
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int i = 2 * ti + r, j = 4 * tj + c;
              if (i <= j && j < DX) {
                const R val = tp[D::oM + i * PP + j] + acc[r][c];
                tp[D::oM + i * PP + j] = val;
                tp[D::oM + j * PP + i] = val;
              }
            }
        }
        LANE_FOR(i, DX) {
          R acc = tp[D::oM + i * PP + P];
#pragma unroll
          for (int a2 = 0; a2 < DU; ++a2) acc = fma(tp[D::oM + (DX + a2) * PP + i], tp[D::oK + a2 * KLD + DX], acc);
          tp[D::oM + i * PP + P] = acc;
        }
        LANE_FOR(one, 1) {
          R ld = R(0);
#pragma unroll
          for (int j = 0; j < DU; ++j) ld += tp[D::oLds + j];
          tp[D::oScal + 3] += ld;
        }
        w.team_sync();
      }
      cp_async_wait_all();
      w.cta_sync();
    }
  }

  // =========================================================================================
  // Forward pass with KL and expected cost
  // =========================================================================================
  if (a.flags & ILQR_FORWARD) {
    const R reg = a.fwd_reg;
    const bool do_kl = (a.flags & ILQR_KL) != 0, do_cost = (a.flags & ILQR_COST) != 0;
    constexpr int oD = D::oCk, oPK = D::oCk + DX * DX, oPL = oPK + DU * DX;      // inside the stage area
    constexpr int oCsV = D::oVec, oFv = D::oVec + PP, oPk = oFv + DXP, oCc = oPk + DUP;
    auto stage_fwd = [&](int t) {
      for (int ci = 0; ci < NC; ++ci) {
        const int b = cond_b(ci);
        if (b < 0) continue;
        R* st = stages + (size_t)ci * D::STAGE;
        const size_t bt = (size_t)b * T + t, bt1 = (size_t)b * (T + 1) + t;
        CTA_FOR(idx, DX * P) {  // F^T: element (k, i) <- F[i][k]
          const int i = idx / P, k = idx % P;
          cp_async(st + D::oFx + k * DXP + i, a.F + bt * szF + idx);
        }
        CTA_FOR(idx, DX * DX) cp_async(st + oD + idx, a.dyn_covar + bt * DX * DX + idx);
        CTA_FOR(i, DX) cp_async(st + oFv + i, a.f + bt * DX + i);
        if (do_cost) {
          CTA_FOR(idx, P * P) cp_async(st + D::oCs + idx, a.C + bt1 * szC + idx);
          CTA_FOR(i, P) cp_async(st + oCsV + i, a.c + bt1 * P + i);
          CTA_FOR(one, 1) cp_async(st + oCc, a.cc + bt1);
        }
        if (do_kl) {
          CTA_FOR(idx, DU * DX) cp_async(st + oPK + idx, a.prevK + bt * szK + idx);
          CTA_FOR(idx, DU * DU) cp_async(st + oPL + idx, a.prevLam + bt * szS + idx);
          CTA_FOR(i, DU) cp_async(st + oPk + i, a.prevk + bt * DU + i);
          CTA_FOR(one, 1) cp_async(st + oCc + 1, a.prev_hld + bt);
        }
      }
    };
    auto stage_policy = [&](int tm, int t) {  // K_t^T, k_t, pol_covar_t of this team's problem (plain loads)
      R* tp = TEAMP(tm);
      const size_t pt = PIDX(tm) * T + t;
      LANE_FOR(idx, DU * DX) {
        const int a2 = idx / DX, k = idx % DX;
        tp[D::oK + k * DUP + a2] = a.K[pt * szK + idx];
      }
      LANE_FOR(i, DU) tp[D::oKv + i] = a.k[pt * DU + i];
      LANE_FOR(idx, DU * DU) tp[D::oPc + idx] = a.pcov[pt * szS + idx];
    };
    stage_fwd(0);
    TEAMS_DO(tm) {
      if (!team_valid(tm)) continue;
      R* tp = TEAMP(tm);
      const int b = team_b(tm);
      LANE_FOR(idx, DX * DX) tp[D::oM + (idx / DX) * PP + idx % DX] = ldg(a.x0S + (size_t)b * DX * DX + idx);
      LANE_FOR(i, DX) tp[D::oM + i * PP + P] = ldg(a.x0m + (size_t)b * DX + i);
      stage_policy(tm, 0);
    }
    cp_async_wait_all();
    w.cta_sync();

    for (int t = 0; t < T; ++t) {
      TEAMS_DO(tm) {
        if (!team_valid(tm)) continue;
        R* tp = TEAMP(tm);
        const R* st = STAGEP(tm);
        // ---- F1: [Sigma_ux | mu_u] = K [Sigma_xx | mu_x] (+ k) (lqg.py:227-230) ----------------------------
        {
          constexpr int NA = (DU + 1) / 2;
          constexpr int NJX = (DX + 3) / 4;             // column tiles covering the state block
          constexpr int JP = P / 4;                     // the tile that holds column P (mu)
          constexpr int NJ = NJX + (JP >= NJX ? 1 : 0);
          LANE_FOR(it, NA * NJ) {
            const int ta = it / NJ, q = it % NJ, tj = q < NJX ? q : JP;
            R acc[2][4];
            tile_km<R, 2, 4, false, true>(tp + D::oK + 2 * ta, DUP, tp + D::oM + 4 * tj, PP, DX, acc);
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const int a2 = 2 * ta + r, j = 4 * tj + c;
                if (a2 < DU) {
                  if (j < DX) {
                    tp[D::oM + (DX + a2) * PP + j] = acc[r][c];
                    tp[D::oM + j * PP + DX + a2] = acc[r][c];
                  } else if (j == P) {
                    tp[D::oM + (DX + a2) * PP + P] = acc[r][c] + tp[D::oKv + a2];
                  }
                }
              }
          }
        }
        w.team_sync();
        // ---- F2: Sigma_uu = Sigma_ux K^T + pol_covar (lqg.py:231) --------------------------------------------
        LANE_FOR(it, DU * (DU + 1) / 2) {
          int ai = 0, rem = it;
          while (rem > ai) { rem -= ai + 1; ++ai; }
          const int bi = rem;  // bi <= ai
          R acc = tp[D::oPc + ai * DU + bi];
#pragma unroll 4
          for (int k = 0; k < DX; ++k) acc = fma(tp[D::oM + k * PP + DX + ai], tp[D::oK + k * DUP + bi], acc);
          tp[D::oM + (DX + ai) * PP + DX + bi] = acc;
          tp[D::oM + (DX + bi) * PP + DX + ai] = acc;
        }
        w.team_sync();
        // ---- F3: Rt_ext = ([F] [Sigma | mu])^T (PE x DX); cost rows; KL delta; optional outputs ----------------
        LANE_FOR(it, D::NTP * D::NTI3) {
          const int tj = it / D::NTI3, ti = it % D::NTI3;
          R acc[4][D::TN3];
          tile_km<R, 4, D::TN3, true, D::TN3 == 4>(tp + D::oM + 4 * tj, PP, st + D::oFx + D::TN3 * ti, DXP, P, acc);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int j = 4 * tj + r;
            if (j < PE) {
#pragma unroll
              for (int c = 0; c < D::TN3; ++c)
                if (D::TN3 * ti + c < DX) tp[D::oR + j * DXP + D::TN3 * ti + c] = acc[r][c];
            }
          }
        }
        if (do_cost) {
          LANE_FOR(i, P) {  // row i of 1/2 [mu^T C mu + tr(C (Sigma + reg I))] + mu^T c
            const R mi = tp[D::oM + i * PP + P];
            R acc = R(0);
#pragma unroll 2
            for (int j = 0; j < P; ++j)
              acc = fma(st[D::oCs + i * P + j], fma(mi, tp[D::oM + j * PP + P], tp[D::oM + j * PP + i]), acc);
            acc = fma(st[D::oCs + i * P + i], reg, acc);
            tp[D::oRow + i] = acc * R(0.5) + mi * st[oCsV + i];
          }
        }
        if (do_kl) {
          LANE_FOR(ai, DU) {  // delta_u = (K_prev - K) mu_x + k_prev - k (lqg.py:294)
            R acc = st[oPk + ai] - tp[D::oKv + ai];
#pragma unroll 4
            for (int k = 0; k < DX; ++k) acc = fma(st[oPK + ai * DX + k] - tp[D::oK + k * DUP + ai], tp[D::oM + k * PP + P], acc);
            tp[D::oDel + ai] = acc;
          }
        }
        if (a.tcov) {
          LANE_FOR(idx, P * P) {
            const int i = idx / P, j = idx % P;
            a.tcov[(PIDX(tm) * (T + 1) + t) * szC + idx] = tp[D::oM + i * PP + j] + (i == j ? reg : R(0));
          }
        }
        if (a.tmean) {
          LANE_FOR(i, P) a.tmean[(PIDX(tm) * (T + 1) + t) * P + i] = tp[D::oM + i * PP + P];
        }
        w.team_sync();
        // ---- F4: Sigma'_xx = (F Sigma) F^T + dyn_covar (upper, mirrored); mu' = F mu + f; scalars ----------------
        LANE_FOR(it, (tri2_count<2, D::NR5, D::NC5>())) {
          int ti, tj;
          tri2_tile<2, D::NR5, D::NC5>(it, ti, tj);
          R acc[2][4];
          tile_km<R, 2, 4, false, true>(tp + D::oR + 2 * ti, DXP, st + D::oFx + 4 * tj, DXP, P, acc);
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int i = 2 * ti + r, j = 4 * tj + c;
              if (i <= j && j < DX) {
                const R val = acc[r][c] + st[oD + i * DX + j];
                tp[D::oM + i * PP + j] = val;
                tp[D::oM + j * PP + i] = val;
              }
            }
        }
        LANE_FOR(i, DX) tp[D::oM + i * PP + P] = tp[D::oR + P * DXP + i] + st[oFv + i];
        LANE_FOR(one, 1) {
          if (do_cost) {
            R acc = R(0);
#pragma unroll
            for (int i = 0; i < P; ++i) acc += tp[D::oRow + i];
            tp[D::oScal + 2] += acc + st[oCc];
          }
          if (do_kl) {
            R tr = R(0), quad = R(0);
#pragma unroll
            for (int i = 0; i < DU; ++i) {
              R row = R(0);
#pragma unroll
              for (int j = 0; j < DU; ++j) {
                const R lij = st[oPL + i * DU + j];
                tr = fma(lij, tp[D::oPc + j * DU + i], tr);
                row = fma(lij, tp[D::oDel + j], row);
              }
              quad = fma(tp[D::oDel + i], row, quad);
            }
            tp[D::oScal + 1] += (tr + quad - R(DU)) * R(0.5) + st[oCc + 1];
          }
        }
        w.team_sync();
        if (t + 1 < T) stage_policy(tm, t + 1);
      }
      w.cta_sync();  // everyone is done with the staged inputs of step t
      if (t + 1 < T) {
        stage_fwd(t + 1);
        cp_async_wait_all();
        w.cta_sync();
      }
    }

    // ---- final row T: state block only (lqg.py:221-222,242) ---------------------------------------------------
    TEAMS_DO(tm) {
      if (!team_valid(tm)) continue;
      R* tp = TEAMP(tm);
      const size_t bT = (size_t)team_b(tm) * (T + 1) + T;
      if (do_cost) {
        LANE_FOR(i, DX) {
          const R mi = tp[D::oM + i * PP + P];
          R acc = R(0);
          for (int j = 0; j < DX; ++j)
            acc = fma(ldg(a.C + bT * szC + (size_t)i * P + j), fma(mi, tp[D::oM + j * PP + P], tp[D::oM + j * PP + i]), acc);
          tp[D::oRow + i] = acc * R(0.5) + mi * ldg(a.c + bT * P + i);
        }
      }
      if (a.tcov) {
        LANE_FOR(idx, P * P) {
          const int i = idx / P, j = idx % P;
          a.tcov[(PIDX(tm) * (T + 1) + T) * szC + idx] = (i < DX && j < DX) ? tp[D::oM + i * PP + j] : R(0);
        }
      }
      if (a.tmean) {
        LANE_FOR(i, P) a.tmean[(PIDX(tm) * (T + 1) + T) * P + i] = i < DX ? tp[D::oM + i * PP + P] : R(0);
      }
      w.team_sync();
    }
  }

  // ---- results ----------------------------------------------------------------------------------------------------
  TEAMS_DO(tm) {
    if (!team_valid(tm)) continue;
    R* tp = TEAMP(tm);
    LANE_FOR(one, 1) {
      const size_t pi = PIDX(tm);
      if ((a.flags & ILQR_FORWARD) && (a.flags & ILQR_COST) && a.cost) {
        R acc = R(0);
        for (int i = 0; i < DX; ++i) acc += tp[D::oRow + i];
        a.cost[pi] = tp[D::oScal + 2] + acc + ldg(a.cc + (size_t)team_b(tm) * (T + 1) + T);
      }
      if ((a.flags & ILQR_FORWARD) && (a.flags & ILQR_KL) && a.kl) a.kl[pi] = (tp[D::oScal + 1] + tp[D::oScal + 3]) / R(T);
      if (a.info) a.info[pi] = (int)tp[D::oScal + 4];
    }
  }
#undef TEAMP
#undef STAGEP
#undef PIDX
}

// shared-memory bytes of a warp-kernel CTA with G teams and NC staged conditions
template <typename R, int DX, int DU>
inline size_t ilqr_warp_bytes(int G, int NC) {
  using D = WarpDims<DX, DU>;
  return ((size_t)G * D::TEAM + (size_t)NC * D::STAGE) * sizeof(R);
}

// teams per CTA: as many eta candidates of one condition as fit (<= 16 warps), else several conditions
template <typename R, int DX, int DU>
inline int ilqr_warp_plan(int B, int E, size_t smem_cap, int force_G, int& G, int& ES, int& grid, size_t& bytes) {
  int g = force_G > 0 ? force_G : 16;
  for (; g >= 1; --g) {
    ES = E < g ? E : g;
    G = (g / ES) * ES;
    if (G < 1) continue;
    bytes = ilqr_warp_bytes<R, DX, DU>(G, G / ES);
    if (bytes <= smem_cap) break;
  }
  if (g < 1) return DONK_SMEM_LIMIT;
  const int NC = G / ES, n_ech = (E + ES - 1) / ES;
  grid = ((B + NC - 1) / NC) * n_ech;
  return DONK_OK;
}

}  // namespace donk
