// donk_b200 — iLQR step, warp-per-problem variant with compile-time state/action dimensions (fast path).
//
// Same math and reference lines as ilqr.cuh (ILQR.step, donk/traj_opt/lqg.py:55-75; backward :126-182;
// forward :185-244; kl_divergence_action :276-306; expected_costs quadratic_costs.py:106-119); selected by
// the launcher when an instantiation for (dX, dU) exists.  How it got here (ncu evidence, profiles/r01):
//   v1 generic CTA kernel: 14 % of the instructions were tile DFMAs, the rest runtime index arithmetic,
//      scalar loops, global loads inside phases                                  ->  7 % of FP64 peak
//   v3 warp per problem, compile-time dims, DFMA register tiles: shared-memory wavefronts at 78 % of peak;
//      a warp-wide LDS.64 costs 2 wavefronts whether or not lanes share addresses, so only >= 8x8
//      register tiles would be FP64-bound                                        -> 18 %
//   v4 (this file): the six products of a step run on the FP64 tensor path (mma.sync.m8n8k4.f64, same
//      37 TFLOP/s as DFMA on B200): one LDS.64 per fragment feeds 8 DFMA-equivalents per lane.
// Structure:
//   * a CTA is G warps = the eta candidates of one condition (or G conditions when E = 1); each warp owns
//     one (condition, eta) problem in its own shared-memory area and synchronises with __syncwarp only;
//   * dimensions are template parameters; every buffer has a leading dimension with ld % 16 in {4, 12}
//     (conflict-free fragment loads), K ranges are zero-padded to multiples of 4;
//   * vectors ride in column P of the matrices: F_ext = [F | f], M[:,P] = q / v / mu, K_ext = [K | k], so
//     w = V f + v, q = c + F^T w, v = q_x + Qxu k, mu_u = K mu_x + k, F mu and K_prev mu_x come out of the
//     matrix products for free;
//   * per-condition inputs of step t are staged once per CTA with cp.async (backward: overlapped with the
//     factorisation / gains / value update, which only touch team-private data);
//   * Quu (dU x dU) is factored redundantly in every lane's registers (no shared-memory round trips);
//     divisions become rsqrt / reciprocal multiplies, C/eta becomes C * (1/eta) (<= 1 ulp).
#pragma once
#include "common.cuh"
#include "ilqr.cuh"

namespace donk {

// Per-phase cycle counters of the warp kernel (profiling builds only: profiles/prof_phases.py compiles a separate
// library with -DDONK_PHASE_PROF; the shipped library has no trace of it).
#if defined(DONK_PHASE_PROF) && !defined(DONK_EMU)
__device__ unsigned long long donk_phase_cycles[32];
#endif
#if defined(DONK_PHASE_PROF) && defined(__CUDA_ARCH__)
#define PHASE_INIT long long ph_t0 = clock64();
#define PHASE_MARK(id)                                                                            \
  do {                                                                                            \
    const long long ph_now = clock64();                                                           \
    if (w.lane == 0) atomicAdd(&donk_phase_cycles[id], (unsigned long long)(ph_now - ph_t0));     \
    ph_t0 = ph_now;                                                                               \
  } while (0)
#else
#define PHASE_INIT
#define PHASE_MARK(id)
#endif

template <int DX, int DU>
struct WarpDims {
  static_assert(DU <= 8 && DX + 1 <= 32, "warp kernel: dU <= 8, dX <= 31");
  static constexpr int P = DX + DU, PE = P + 1;
  static constexpr int DX4 = (DX + 3) / 4 * 4, P4 = (P + 3) / 4 * 4, DU4 = (DU + 3) / 4 * 4;
  static constexpr int PP = ld_pad(PE);       // leading dimension of M, W_ext, F_ext
  static constexpr int DXP = ld_pad(DX);      // leading dimension of Rt_ext, F^T
  static constexpr int KLD = ld_pad(DX + 1);  // backward: Kb[a][j] = [K | k]
  static constexpr int KTL = ld_pad(DU);      // forward: Kt[k][a] = K^T, and K_prev^T in the stage area
  static constexpr int NBP = (PE + 7) / 8;    // 8-column blocks covering columns 0..P
  static constexpr int NBX = (DX + 7) / 8;    // ... covering the state block
  static constexpr int NBX1 = (DX + 8) / 8;   // ... covering columns 0..DX (state block + the k column)
  static constexpr int BP = P / 8;            // the block that holds column P
  // team area (reals)
  static constexpr int MSZ = P * PP;
  static constexpr int RSZ = (DX4 * PP > P4 * DXP ? DX4 * PP : P4 * DXP) + 8;
  static constexpr int KSZ = (DU4 * KLD > DX4 * KTL ? DU4 * KLD : DX4 * KTL) + 8;
  static constexpr int oM = 0, oR = oM + MSZ, oK = oR + RSZ, oKv = oK + KSZ;
  static constexpr int oPc = oKv + 8, oLds = oPc + (DU * DU + 3) / 4 * 4, oDel = oLds + 8, oRow = oDel + 8;
  static constexpr int oScal = oRow + 32, TEAM = oScal + 8;  // scal: 0 1/eta 1 kl 2 cost 3 logdet 4 notpd
  // stage area per condition (reals): backward and forward views share the storage.  Only what the products
  // read many times is staged (F_t, K_prev_t); C_t, C_kl_t, dyn_covar_t are read once per problem and step and
  // come straight from global memory, prefetched into registers ahead of the product they are added to.
  static constexpr int oFx = 0;  // bwd: F_ext (DX4 x PP); fwd: F^T (P4 x DXP)
  static constexpr int FXS = (DX4 * PP > P4 * DXP ? DX4 * PP : P4 * DXP) + 8;
  static constexpr int oPK = oFx + FXS;                    // fwd: K_prev^T (DX4 x KTL)
  static constexpr int oPL = oPK + DX4 * KTL + 8;          // fwd: Lam_prev (DU x DU)
  static constexpr int oVec = oPL + (DU * DU + 3) / 4 * 4; // fwd: f (DXP) | k_prev (8) | cc, phld (4)
  static constexpr int STAGE = oVec + DXP + 8 + 4;
};

template <typename R, int DX, int DU>
DONK_D void ilqr_warp_cta(const WCtx& w, const IlqrArgs<R>& a, int cta, int ES, R* smem) {
  using D = WarpDims<DX, DU>;
  constexpr int P = D::P, PE = D::PE, PP = D::PP, DXP = D::DXP, KLD = D::KLD, KTL = D::KTL;
  constexpr int DX4 = D::DX4, P4 = D::P4, DU4 = D::DU4;
  const int G = w.G, T = a.T, E = a.E;
  const int NC = G / ES, n_ech = (E + ES - 1) / ES;
  const int cg = cta / n_ech, ec = cta % n_ech;
  R* teams = smem;
  R* stages = smem + (size_t)G * D::TEAM;
  const R gamma = a.gamma;
  constexpr size_t szF = (size_t)DX * P, szC = (size_t)P * P, szK = (size_t)DU * DX, szS = (size_t)DU * DU;

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

  PHASE_INIT
  // zero everything once: K paddings must be exact zeros, over-read rows/columns must be finite
  CTA_FOR(idx, G * D::TEAM + NC * D::STAGE) smem[idx] = R(0);
  w.cta_sync();
  TEAMS_DO(tm) {
    if (!team_valid(tm)) continue;
    R* tp = TEAMP(tm);
    LANE_FOR(one, 1) tp[D::oScal + 0] = a.eta ? R(1) / a.eta[PIDX(tm)] : R(1);
  }

  // =========================================================================================
  // Backward pass
  // =========================================================================================
  if (a.flags & ILQR_BACKWARD) {
    // async staging of step t: F_ext = [F_t | f_t | 0] (rows DX..DX4-1 stay zero), C_t, C_kl_t, c_t, c_kl_t
    auto stage_bwd = [&](int t) {
      for (int ci = 0; ci < NC; ++ci) {
        const int b = cond_b(ci);
        if (b < 0) continue;
        R* st = stages + (size_t)ci * D::STAGE;
        const size_t bt = (size_t)b * T + t;
        CTA_FOR(idx, DX * PE) {
          const int k = idx / PE, j = idx % PE;  // compile-time divisors
          if (j < P) cp_async(st + D::oFx + k * PP + j, a.F + bt * szF + (size_t)k * P + j);
          else cp_async(st + D::oFx + k * PP + P, a.f + bt * DX + k);
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
    PHASE_MARK(12);

    for (int t = T - 1; t >= 0; --t) {
      TEAMS_DO(tm) {
        if (!team_valid(tm)) continue;
        R* tp = TEAMP(tm);
        const R* st = STAGEP(tm);
        if (t > 0) {  // pull next step's C / C_kl rows into L2 while this step computes (one 128-byte line per lane)
          const int b = team_b(tm);
          constexpr int LINES = (P * P * (int)sizeof(R) + 127) / 128;
          LANE_FOR(ln, LINES) {
            if ((tm % ES) * 2 < ES) prefetch_l2(a.C + ((size_t)b * (T + 1) + t - 1) * szC + (size_t)ln * (128 / sizeof(R)));
            else if (a.C_kl) prefetch_l2(a.C_kl + ((size_t)b * T + t - 1) * szC + (size_t)ln * (128 / sizeof(R)));
          }
        }
        // ---- B1: W_ext = V [F | f] (+ v in column P) ----------------------------------------------------
        warp_mma<R, D::NBX, D::NBP>(w, tp + D::oM, PP, st + D::oFx, PP, DX4 / 4, [&](int i, int j, R v0, R v1) {
          if (i < DX && j < PP) {
            if (j == P) v0 += tp[D::oM + i * PP + P];
            if (j + 1 == P) v1 += tp[D::oM + i * PP + P];
            st2(tp + D::oR + i * PP + j, v0, v1);  // j even, PP even: 16-byte aligned
          }
        });
        w.team_sync();
        PHASE_MARK(0);
        // ---- B2: Q_ext = C_ext + gamma [F|f]^T W_ext, upper blocks; column P is q.  Only the xu block is mirrored
        //      (Qux feeds the gains and the value update); Qxx / Quu are consumed from their upper triangles. ----------
        {
          const R ie = tp[D::oScal + 0];
          const int b = team_b(tm);
          const R* Cg = a.C + ((size_t)b * (T + 1) + t) * szC;
          const R* Kg = a.C_kl ? a.C_kl + ((size_t)b * T + t) * szC : nullptr;
          const R* cg = a.c + ((size_t)b * (T + 1) + t) * P;
          const R* kg = a.c_kl ? a.c_kl + ((size_t)b * T + t) * P : nullptr;
          // C and C_kl entries (i, j), (i, j+1) of the fragment: raw loads issued before the product starts
          auto pre = [&](int i, int j, R* pv) {
            if (i >= P || i > j + 1) return;
            if (j + 1 < P) {
              ldg2(Cg + i * P + j, P % 2 == 0, pv[0], pv[1]);
              if (Kg) ldg2(Kg + i * P + j, P % 2 == 0, pv[2], pv[3]);
            } else {
              if (j < P) { pv[0] = ldg(Cg + i * P + j); if (Kg) pv[2] = ldg(Kg + i * P + j); }
              const int h = j == P ? 0 : 1;  // which of the two columns is the vector column P
              if (j == P || j + 1 == P) { pv[h] = ldg(cg + i); if (kg) pv[2 + h] = ldg(kg + i); }
            }
          };
          auto epi = [&](int i, int j, R v0, R v1, const R* pv) {
            if (i >= P || i > j + 1 || j > P) return;
            const R q0 = fma(gamma, v0, fma(pv[0], ie, pv[2])), q1 = fma(gamma, v1, fma(pv[1], ie, pv[3]));
            if (i <= j && j + 1 <= P) {
              st2(tp + D::oM + i * PP + j, q0, q1);
            } else {
              if (i <= j) tp[D::oM + i * PP + j] = q0;
              if (j + 1 <= P) tp[D::oM + i * PP + j + 1] = q1;
            }
            if (i < DX) {  // mirror Qxu -> Qux
              if (j >= DX && j < P) tp[D::oM + j * PP + i] = q0;
              if (j + 1 >= DX && j + 1 < P) tp[D::oM + (j + 1) * PP + i] = q1;
            }
          };
#define DONK_B2_ROW(RB_ROW)                                                                                      \
  if (RB_ROW < D::NBP)                                                                                           \
    warp_mma_pre<R, 1, (D::NBP - RB_ROW > 0 ? D::NBP - RB_ROW : 1), 4>(                                           \
        w, st + D::oFx + 8 * RB_ROW, PP, tp + D::oR + 8 * RB_ROW, PP, DX4 / 4,                                    \
        [&](int i, int j, R* pv) { pre(8 * RB_ROW + i, 8 * RB_ROW + j, pv); },                                    \
        [&](int i, int j, R v0, R v1, const R* pv) { epi(8 * RB_ROW + i, 8 * RB_ROW + j, v0, v1, pv); });
          DONK_B2_ROW(0) DONK_B2_ROW(1) DONK_B2_ROW(2) DONK_B2_ROW(3) DONK_B2_ROW(4)
#undef DONK_B2_ROW
        }
        w.team_sync();
        PHASE_MARK(1);
      }
      // all teams are done with the staged inputs of step t: fetch step t-1 while the factorisation, gains and
      // value update (team-private data only) run
      w.cta_sync();
      if (t > 0) stage_bwd(t - 1);
      PHASE_MARK(2);
      TEAMS_DO(tm) {
        if (!team_valid(tm)) continue;
        R* tp = TEAMP(tm);
        const size_t pt = PIDX(tm) * T + t;
        // ---- B34: Quu = L L^T in registers (every lane, redundantly); [K | k] = -Quu^-1 [Qux | q_u] --------
        LANE_FOR(ln, 32) {
          R Lm[DU][DU];  // lower triangle used
#pragma unroll
          for (int i = 0; i < DU; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) Lm[i][j] = tp[D::oM + (DX + j) * PP + DX + i];  // upper triangle of Quu
          R idg[DU];
          bool bad = false;
#pragma unroll
          for (int j = 0; j < DU; ++j) {
            R sj = Lm[j][j];
#pragma unroll
            for (int k = 0; k < j; ++k) sj = fma(-Lm[j][k], Lm[j][k], sj);
            bad = bad || !(sj > R(0));
            const R rs = rsqrt_r(sj);
            idg[j] = rs;         // 1 / L_jj
            Lm[j][j] = sj * rs;  // L_jj
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
#pragma unroll
          for (int c = 0; c < DU; ++c) {  // in-place inverse of the lower factor: Li = L^-1
            Lm[c][c] = idg[c];
#pragma unroll
            for (int i = c + 1; i < DU; ++i) {
              R acc = R(0);
#pragma unroll
              for (int k = c; k < i; ++k) acc = fma(Lm[i][k], Lm[k][c], acc);
              Lm[i][c] = -acc * idg[i];
            }
          }
          {  // pol_covar = Quu^-1 = Li^T Li (symmetric by construction, lqg.py:165-166); lane e stores entry e
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
                  const R quu = tp[D::oM + (DX + j) * PP + DX + i];
                  a.pinv[pt * szS + i * DU + j] = quu;
                  a.pinv[pt * szS + j * DU + i] = quu;
                }
              }
          }
          if (ln <= DX) {  // column ln of K (ln < DX) or k (ln == DX): z = -Li^T (Li x), x = Qux[:, ln] / q_u
            const int col = ln < DX ? ln : P;
            R x[DU];
#pragma unroll
            for (int b2 = 0; b2 < DU; ++b2) x[b2] = tp[D::oM + (DX + b2) * PP + col];
#pragma unroll
            for (int i = DU - 1; i >= 0; --i) {
              R acc = R(0);
#pragma unroll
              for (int k = 0; k <= i; ++k) acc = fma(Lm[i][k], x[k], acc);
              x[i] = acc;
            }
#pragma unroll
            for (int i = 0; i < DU; ++i) {
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
        PHASE_MARK(3);
        // ---- B5: [V | v] = [Qxx | q_x] + Qxu [K | k] (upper blocks, mirrored) (lqg.py:178-180) ---------------
        {
          auto epi = [&](int i, int j, R v0, R v1) {
            if (i >= DX || i > j + 1 || j > DX) return;
            if (i <= j && j + 1 < DX) {
              R m0, m1;
              ld2(tp + D::oM + i * PP + j, m0, m1);
              m0 += v0; m1 += v1;
              st2(tp + D::oM + i * PP + j, m0, m1);
              tp[D::oM + j * PP + i] = m0;
              tp[D::oM + (j + 1) * PP + i] = m1;
            } else {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int jj = j + h;
                const R acc = h ? v1 : v0;
                if (jj < DX) {
                  if (i <= jj) {
                    const R val = tp[D::oM + i * PP + jj] + acc;
                    tp[D::oM + i * PP + jj] = val;
                    tp[D::oM + jj * PP + i] = val;
                  }
                } else if (jj == DX) {
                  tp[D::oM + i * PP + P] += acc;
                }
              }
            }
          };
#define DONK_B5_ROW(RB_ROW)                                                                                        \
  if (RB_ROW < D::NBX)                                                                                             \
    warp_mma<R, 1, (D::NBX1 - RB_ROW > 0 ? D::NBX1 - RB_ROW : 1)>(                                                  \
        w, tp + D::oM + DX * PP + 8 * RB_ROW, PP, tp + D::oK + 8 * RB_ROW, KLD, DU4 / 4,                            \
        [&](int i, int j, R v0, R v1) { epi(8 * RB_ROW + i, 8 * RB_ROW + j, v0, v1); });
          DONK_B5_ROW(0) DONK_B5_ROW(1) DONK_B5_ROW(2) DONK_B5_ROW(3)
#undef DONK_B5_ROW
        }
        LANE_FOR(one, 1) {
          R ld = R(0);
#pragma unroll
          for (int j = 0; j < DU; ++j) ld += tp[D::oLds + j];
          tp[D::oScal + 3] += ld;
        }
        w.team_sync();
        PHASE_MARK(4);
      }
      cp_async_wait_all();
      w.cta_sync();
      PHASE_MARK(5);
    }
  }

  // =========================================================================================
  // Forward pass with KL and expected cost
  // =========================================================================================
  if (a.flags & ILQR_FORWARD) {
    const R reg = a.fwd_reg;
    const bool do_kl = (a.flags & ILQR_KL) != 0, do_cost = (a.flags & ILQR_COST) != 0;
    constexpr int oPK = D::oPK, oPL = D::oPL, oFv = D::oVec, oPk = oFv + DXP, oCc = oPk + 8;  // inside the stage area
    // the stage and K areas change layout: clear the paddings the forward products rely on
    CTA_FOR(idx, NC * D::STAGE) stages[idx] = R(0);
    TEAMS_DO(tm) {
      R* tp = TEAMP(tm);
      LANE_FOR(idx, D::KSZ) tp[D::oK + idx] = R(0);
    }
    w.cta_sync();
    auto stage_fwd = [&](int t) {
      for (int ci = 0; ci < NC; ++ci) {
        const int b = cond_b(ci);
        if (b < 0) continue;
        R* st = stages + (size_t)ci * D::STAGE;
        const size_t bt = (size_t)b * T + t, bt1 = (size_t)b * (T + 1) + t;
        CTA_FOR(idx, DX * P) {  // F^T: element (k, i) <- F[i][k]; rows P..P4-1 stay zero
          const int i = idx / P, k = idx % P;
          cp_async(st + D::oFx + k * DXP + i, a.F + bt * szF + idx);
        }
        CTA_FOR(i, DX) cp_async(st + oFv + i, a.f + bt * DX + i);
        if (do_cost) {
          CTA_FOR(one, 1) cp_async(st + oCc, a.cc + bt1);
        }
        if (do_kl) {
          CTA_FOR(idx, DU * DX) {  // K_prev^T: element (k, a) <- K_prev[a][k]
            const int a2 = idx / DX, k = idx % DX;
            cp_async(st + oPK + k * KTL + a2, a.prevK + bt * szK + idx);
          }
          CTA_FOR(idx, DU * DU) cp_async(st + oPL + idx, a.prevLam + bt * szS + idx);
          CTA_FOR(i, DU) cp_async(st + oPk + i, a.prevk + bt * DU + i);
          CTA_FOR(one, 1) cp_async(st + oCc + 1, a.prev_hld + bt);
        }
      }
    };
    auto stage_policy = [&](int tm, int t) {  // K_t^T, k_t, pol_covar_t of this team's problem (async)
      R* tp = TEAMP(tm);
      const size_t pt = PIDX(tm) * T + t;
      LANE_FOR(idx, DU * DX) {
        const int a2 = idx / DX, k = idx % DX;
        cp_async(tp + D::oK + k * KTL + a2, a.K + pt * szK + idx);
      }
      LANE_FOR(i, DU) cp_async(tp + D::oKv + i, a.k + pt * DU + i);
      LANE_FOR(idx, DU * DU) cp_async(tp + D::oPc + idx, a.pcov + pt * szS + idx);
    };
    stage_fwd(0);
    TEAMS_DO(tm) {
      if (!team_valid(tm)) continue;
      R* tp = TEAMP(tm);
      const int b = team_b(tm);
      LANE_FOR(idx, D::MSZ) tp[D::oM + idx] = R(0);
      w.team_sync();
      LANE_FOR(idx, DX * DX) tp[D::oM + (idx / DX) * PP + idx % DX] = ldg(a.x0S + (size_t)b * DX * DX + idx);
      LANE_FOR(i, DX) tp[D::oM + i * PP + P] = ldg(a.x0m + (size_t)b * DX + i);
      stage_policy(tm, 0);
    }
    cp_async_wait_all();
    w.cta_sync();
    PHASE_MARK(12);

    for (int t = 0; t < T; ++t) {
      TEAMS_DO(tm) {
        if (!team_valid(tm)) continue;
        R* tp = TEAMP(tm);
        const R* st = STAGEP(tm);
        const int b = team_b(tm);
        if (t + 1 < T) {  // pull next step's C / dyn_covar rows into L2
          constexpr int LINES = (P * P * (int)sizeof(R) + 127) / 128, LINESD = (DX * DX * (int)sizeof(R) + 127) / 128;
          LANE_FOR(ln, LINES + LINESD) {
            if (ln < LINES) { if (do_cost) prefetch_l2(a.C + ((size_t)b * (T + 1) + t + 1) * szC + (size_t)ln * (128 / sizeof(R))); }
            else prefetch_l2(a.dyn_covar + ((size_t)b * T + t + 1) * DX * DX + (size_t)(ln - LINES) * (128 / sizeof(R)));
          }
        }
        // ---- F1: [Sigma_ux | mu_u] = K [Sigma_xx | mu_x] (+ k); K_prev mu_x (lqg.py:227-230, 294) --------------
        warp_mma<R, 1, D::NBP>(w, tp + D::oK, KTL, tp + D::oM, PP, DX4 / 4, [&](int a2, int j, R v0, R v1) {
          if (a2 < DU) {
            if (j + 1 < DX) {
              st2(tp + D::oM + (DX + a2) * PP + j, v0, v1);
              tp[D::oM + j * PP + DX + a2] = v0;
              tp[D::oM + (j + 1) * PP + DX + a2] = v1;
            } else {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int jj = j + h;
                const R acc = h ? v1 : v0;
                if (jj < DX) {
                  tp[D::oM + (DX + a2) * PP + jj] = acc;
                  tp[D::oM + jj * PP + DX + a2] = acc;
                } else if (jj == P) {
                  tp[D::oM + (DX + a2) * PP + P] = acc + tp[D::oKv + a2];
                }
              }
            }
          }
        });
        if (do_kl) {
          warp_mma<R, 1, 1>(w, st + oPK, KTL, tp + D::oM + 8 * D::BP, PP, DX4 / 4, [&](int a2, int j, R v0, R v1) {
            if (a2 < DU) {
              if (8 * D::BP + j == P) tp[D::oDel + a2] = v0 + st[oPk + a2];
              if (8 * D::BP + j + 1 == P) tp[D::oDel + a2] = v1 + st[oPk + a2];
            }
          });
        }
        w.team_sync();
        PHASE_MARK(6);
        // ---- F2: Sigma_uu = Sigma_ux K^T + pol_covar (upper, mirrored) (lqg.py:231); delta_u ---------------------
        warp_mma<R, 1, 1>(w, tp + D::oM + DX, PP, tp + D::oK, KTL, DX4 / 4, [&](int a2, int j, R v0, R v1) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int bb = j + h;
            if (a2 <= bb && bb < DU) {
              const R val = (h ? v1 : v0) + tp[D::oPc + a2 * DU + bb];
              tp[D::oM + (DX + a2) * PP + DX + bb] = val;
              tp[D::oM + (DX + bb) * PP + DX + a2] = val;
            }
          }
        });
        if (do_kl) {
          LANE_FOR(a2, DU) tp[D::oDel + a2] -= tp[D::oM + (DX + a2) * PP + P];  // (K_prev mu_x + k_prev) - mu_u
        }
        w.team_sync();
        PHASE_MARK(7);
        // ---- F3: Rt_ext = ([F] [Sigma | mu])^T (PE x DX).  The expected cost of step t rides on the A fragments:
        //      every lane sees 1/32 of [Sigma | mu] anyway and adds C_kj (Sigma_kj + mu_k mu_j [+ reg]) for it --------
        {
          const R* Cg = a.C + ((size_t)b * (T + 1) + t) * szC;
          const R* cg = a.c + ((size_t)b * (T + 1) + t) * P;
#if defined(__CUDA_ARCH__)
          const int g = w.lane >> 2, tq = w.lane & 3;
          R c[D::NBP][D::NBX][2];
#pragma unroll
          for (int ra = 0; ra < D::NBP; ++ra)
#pragma unroll
            for (int rb = 0; rb < D::NBX; ++rb) { c[ra][rb][0] = R(0); c[ra][rb][1] = R(0); }
          R muj[D::NBP], cn[D::NBP], cn2[D::NBP];
          auto loadC = [&](int k, int ra) {
            const int j = 8 * ra + g;
            return (do_cost && k < P && j < P) ? ldg(Cg + k * P + j) : R(0);
          };
#pragma unroll
          for (int ra = 0; ra < D::NBP; ++ra) {
            muj[ra] = 8 * ra + g < P ? tp[D::oM + (8 * ra + g) * PP + P] : R(0);
            cn[ra] = loadC(tq, ra);
            cn2[ra] = loadC(tq + 4, ra);
          }
          R cost = R(0);
          const R* ap = tp + D::oM + tq * PP + g;
          const R* bp = st + D::oFx + tq * DXP + g;
          for (int k4 = 0; k4 < P4 / 4; ++k4) {
            const int k = 4 * k4 + tq;
            R cc_[D::NBP], af[D::NBP], bf[D::NBX];
#pragma unroll
            for (int ra = 0; ra < D::NBP; ++ra) { cc_[ra] = cn[ra]; cn[ra] = cn2[ra]; cn2[ra] = loadC(k + 8, ra); af[ra] = ap[8 * ra]; }
#pragma unroll
            for (int rb = 0; rb < D::NBX; ++rb) bf[rb] = bp[8 * rb];
            if (do_cost) {
              const R muk = k < P ? tp[D::oM + k * PP + P] : R(0);
#pragma unroll
              for (int ra = 0; ra < D::NBP; ++ra) {
                R sv = fma(muk, muj[ra], af[ra]);
                if (k == 8 * ra + g) sv += reg;
                cost = fma(cc_[ra], sv, cost);
              }
            }
#pragma unroll
            for (int ra = 0; ra < D::NBP; ++ra)
#pragma unroll
              for (int rb = 0; rb < D::NBX; ++rb) {
                if (sizeof(R) == 8) {
                  double d0 = (double)c[ra][rb][0], d1 = (double)c[ra][rb][1];
                  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                               : "+d"(d0), "+d"(d1)
                               : "d"((double)af[ra]), "d"((double)bf[rb]));
                  c[ra][rb][0] = (R)d0;
                  c[ra][rb][1] = (R)d1;
                }
              }
            ap += 4 * PP;
            bp += 4 * DXP;
          }
          if (sizeof(R) != 8) {  // float: no tensor path, each lane computes its two outputs per fragment
            for (int k = 0; k < P4; ++k)
#pragma unroll
              for (int ra = 0; ra < D::NBP; ++ra) {
                const R av = tp[D::oM + k * PP + 8 * ra + g];
#pragma unroll
                for (int rb = 0; rb < D::NBX; ++rb) {
                  c[ra][rb][0] = fma(av, st[D::oFx + k * DXP + 8 * rb + 2 * tq], c[ra][rb][0]);
                  c[ra][rb][1] = fma(av, st[D::oFx + k * DXP + 8 * rb + 2 * tq + 1], c[ra][rb][1]);
                }
              }
          }
#pragma unroll
          for (int ra = 0; ra < D::NBP; ++ra)
#pragma unroll
            for (int rb = 0; rb < D::NBX; ++rb) {
              const int j = 8 * ra + g, i = 8 * rb + 2 * tq;
              if (j < PE) {
                if (i + 1 < DX) st2(tp + D::oR + j * DXP + i, c[ra][rb][0], c[ra][rb][1]);
                else if (i < DX) tp[D::oR + j * DXP + i] = c[ra][rb][0];
              }
            }
          if (do_cost) {
            cost *= R(0.5);
            if (w.lane < P) cost = fma(tp[D::oM + w.lane * PP + P], ldg(cg + w.lane), cost);
            team_accumulate(w, tp + D::oScal + 2, cost);
          }
#else
          for (int j = 0; j < PE; ++j)
            for (int i = 0; i < DX; ++i) {
              R acc = R(0);
              for (int k = 0; k < P4; ++k) acc = fma(tp[D::oM + k * PP + j], st[D::oFx + k * DXP + i], acc);
              tp[D::oR + j * DXP + i] = acc;
            }
          if (do_cost) {
            R cost = R(0);
            for (int k = 0; k < P; ++k)
              for (int j = 0; j < P; ++j) {
                R sv = fma(tp[D::oM + k * PP + P], tp[D::oM + j * PP + P], tp[D::oM + k * PP + j]);
                if (k == j) sv += reg;
                cost = fma(Cg[k * P + j], sv, cost);
              }
            cost *= R(0.5);
            for (int i = 0; i < P; ++i) cost = fma(tp[D::oM + i * PP + P], cg[i], cost);
            tp[D::oScal + 2] += cost;
          }
#endif
        }
        PHASE_MARK(8);
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
        PHASE_MARK(9);
        // ---- F4: Sigma'_xx = (F Sigma) F^T + dyn_covar (upper blocks, mirrored); mu' = F mu + f; scalars -------------
        {
          const R* Dg = a.dyn_covar + ((size_t)b * T + t) * DX * DX;
          auto pre = [&](int i, int j, R* pv) {
            if (i >= DX || i > j + 1 || j >= DX) return;
            if (j + 1 < DX) ldg2(Dg + i * DX + j, DX % 2 == 0, pv[0], pv[1]);
            else pv[0] = ldg(Dg + i * DX + j);
          };
          auto epi = [&](int i, int j, R v0, R v1, const R* pv) {
            if (i >= DX || i > j + 1 || j >= DX) return;
            const R s0 = v0 + pv[0], s1 = v1 + pv[1];
            if (i <= j && j + 1 < DX) {
              st2(tp + D::oM + i * PP + j, s0, s1);
              tp[D::oM + j * PP + i] = s0;
              tp[D::oM + (j + 1) * PP + i] = s1;
            } else {
              if (i <= j) { tp[D::oM + i * PP + j] = s0; tp[D::oM + j * PP + i] = s0; }
              if (j + 1 < DX) { tp[D::oM + i * PP + j + 1] = s1; tp[D::oM + (j + 1) * PP + i] = s1; }
            }
          };
#define DONK_F4_ROW(RB_ROW)                                                                                        \
  if (RB_ROW < D::NBX)                                                                                             \
    warp_mma_pre<R, 1, (D::NBX - RB_ROW > 0 ? D::NBX - RB_ROW : 1), 2>(                                             \
        w, tp + D::oR + 8 * RB_ROW, DXP, st + D::oFx + 8 * RB_ROW, DXP, P4 / 4,                                     \
        [&](int i, int j, R* pv) { pre(8 * RB_ROW + i, 8 * RB_ROW + j, pv); },                                      \
        [&](int i, int j, R v0, R v1, const R* pv) { epi(8 * RB_ROW + i, 8 * RB_ROW + j, v0, v1, pv); });
          DONK_F4_ROW(0) DONK_F4_ROW(1) DONK_F4_ROW(2) DONK_F4_ROW(3)
#undef DONK_F4_ROW
        }
        LANE_FOR(i, DX) tp[D::oM + i * PP + P] = tp[D::oR + P * DXP + i] + st[oFv + i];
        if (do_kl) {
          LANE_FOR(ln, DU) {  // row ln of  tr(Lam_prev Sigma_new) + delta^T Lam_prev delta  (lqg.py:295-304)
            R tr = R(0), row = R(0);
#pragma unroll
            for (int j = 0; j < DU; ++j) {
              const R lij = st[oPL + ln * DU + j];
              tr = fma(lij, tp[D::oPc + j * DU + ln], tr);
              row = fma(lij, tp[D::oDel + j], row);
            }
            tp[D::oLds + ln] = fma(tp[D::oDel + ln], row, tr);
          }
        }
        w.team_sync();
        LANE_FOR(one, 1) {
          if (do_kl) {
            R s = R(0);
#pragma unroll
            for (int j = 0; j < DU; ++j) s += tp[D::oLds + j];
            tp[D::oScal + 1] += (s - R(DU)) * R(0.5) + st[oCc + 1];
          }
          if (do_cost) tp[D::oScal + 2] += st[oCc];
        }
        if (t + 1 < T) stage_policy(tm, t + 1);
        PHASE_MARK(10);
      }
      w.cta_sync();  // everyone is done with the staged inputs of step t
      if (t + 1 < T) {
        stage_fwd(t + 1);
        cp_async_wait_all();
        w.cta_sync();
      }
      PHASE_MARK(11);
    }

    // ---- final row T: state block only (lqg.py:221-222,242) ---------------------------------------------------
    TEAMS_DO(tm) {
      if (!team_valid(tm)) continue;
      R* tp = TEAMP(tm);
      const size_t bT = (size_t)team_b(tm) * (T + 1) + T;
      if (do_cost) {
        LANE_FOR(ln, 32) {
          R acc = R(0);
          for (int i = ln; i < DX; i += 32) {
            const R mi = tp[D::oM + i * PP + P];
            R row = R(0);
            for (int j = 0; j < DX; ++j)
              row = fma(ldg(a.C + bT * szC + (size_t)i * P + j), fma(mi, tp[D::oM + j * PP + P], tp[D::oM + j * PP + i]), row);
            acc += row * R(0.5) + mi * ldg(a.c + bT * P + i);
          }
          tp[D::oRow + ln] = acc;
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
        for (int i = 0; i < 32; ++i) acc += tp[D::oRow + i];
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
  // 8 warps per CTA by default: two CTAs share an SM and drift apart, so the tensor-heavy products of one overlap
  // the factorisation / epilogue / staging phases of the other (one 16-warp CTA runs its warps in lock-step)
  int g = force_G > 0 ? force_G : 8;
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
