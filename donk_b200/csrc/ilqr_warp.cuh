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
    if (w.lane == 0 && (blockIdx.x & 7) == 0)  /* every 8th CTA records: keeps the perturbation small */ \
      atomicAdd(&donk_phase_cycles[id], (unsigned long long)(ph_now - ph_t0));                    \
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
  // Forward pass: pol_covar_t (DU x DU) and k_t (DU).  The fragments of K^T (DX4 x KTL) only ever touch columns 0..7 of a
  // row, so with KTL = 12 the columns 8..11 are free: the two small arrays live there in strips of four (PC_IN_K).  The 50
  // reals this and the packed scalars below save are what lets a tenth (20,6) problem fit an SM at E = 1 (ILQR.optimize).
  static constexpr int PCN = (DU * DU + 3) / 4 * 4;
  static constexpr bool PC_IN_K = KTL >= 12 && DX4 * 4 >= PCN + DU4;
  static constexpr int oPc = oKv + (PC_IN_K ? 0 : 8);
  static constexpr int DUE = (DU + 1) / 2 * 2;
  static constexpr int oLds = oPc + (PC_IN_K ? 0 : PCN), oDel = oLds + DUE;
  static constexpr int oScal = oDel + DUE;   // scal: 0 1/eta 1 kl 2 cost 3 logdet 4 notpd
  // oRow: forward: compact copy of mu (P values, conflict-free reads in the cost phase); end: per-lane cost partials
  static constexpr int oRow = oScal + 6, TEAM = oRow + (P4 > 32 ? P4 : 32);
  static_assert(oRow % 2 == 0 && TEAM % 2 == 0 && oK % 2 == 0, "16-byte accesses: even offsets");
  DONK_HD static constexpr int pc_off(int e) { return PC_IN_K ? oK + (e >> 2) * KTL + 8 + (e & 3) : oPc + e; }
  DONK_HD static constexpr int kv_off(int i) { return PC_IN_K ? pc_off(PCN + i) : oKv + i; }
  // stage area per condition (reals): backward and forward views share the storage.  Everything the teams of a
  // condition share and that does not depend on eta is staged one step ahead with cp.async: F_t, C_kl_t (backward);
  // F_t^T, K_prev_t, Lam_prev_t, dyn_covar_t (forward).  C_t (scaled by 1/eta, and the cost weights) is the one input
  // that does not fit: it comes straight from global memory into registers, issued a whole product ahead of its use.
  static constexpr int oFx = 0;  // bwd: F_ext (DX4 x PP); fwd: F^T (P4 x DXP)
  static constexpr int FXS = (DX4 * PP > P4 * DXP ? DX4 * PP : P4 * DXP) + 8;
  static constexpr int PKL = PP, DXE = DXP;                // row strides % 16 in {4, 12}: conflict-free pair reads
  static constexpr int oCk = oFx + FXS;                    // bwd: [C_kl_t | c_kl_t] (P x PKL), vector in column P
  static constexpr int BWD_END = oCk + P * PKL;
  static constexpr int oPK = oFx + FXS;                    // fwd: K_prev^T (DX4 x KTL)
  static constexpr int oPL = oPK + DX4 * KTL + 8;          // fwd: Lam_prev (DU x DU)
  static constexpr int oVec = oPL + (DU * DU + 3) / 4 * 4; // fwd: f (DXP) | k_prev (8) | cc, phld (4)
  static constexpr int oD = oVec + DXP + 8 + 4;            // fwd: dyn_covar_t (DX x DXE)
  static constexpr int FWD_END = oD + DX * DXE;
  static constexpr int STAGE = ((BWD_END > FWD_END ? BWD_END : FWD_END) + 1) / 2 * 2;
};

// FULL: the call computes everything (backward + forward + KL + cost, eta given, no trajectory outputs) - the
// ILQR.step / ILQR.optimize case.  The option tests then fold at compile time instead of being re-read from the
// kernel parameters in every phase (the kernel is bound by issued instructions).
template <typename R, int DX, int DU>
DONK_HD bool ilqr_warp_is_full(const IlqrArgs<R>& a) {
  return a.flags == (ILQR_BACKWARD | ILQR_FORWARD | ILQR_KL | ILQR_COST) && a.C_kl && a.c_kl && a.eta && a.kl &&
         a.cost && !a.tcov && !a.tmean;
}

template <typename R, int DX, int DU, bool FULL>
DONK_D void ilqr_warp_cta(const WCtx& w, const IlqrArgs<R>& a, int cta, int ES, R* smem) {
  using D = WarpDims<DX, DU>;
  const bool opt_bwd = FULL || (a.flags & ILQR_BACKWARD) != 0, opt_fwd = FULL || (a.flags & ILQR_FORWARD) != 0;
  const bool has_ckl = FULL || a.C_kl != nullptr, has_cklv = FULL || a.c_kl != nullptr, has_eta = FULL || a.eta != nullptr;
  const bool want_tcov = !FULL && a.tcov != nullptr, want_tmean = !FULL && a.tmean != nullptr;
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
  auto team_valid_raw = [&](int tm) {
    const int b = team_b(tm), e = team_e(tm);
    return b < a.B && e < E && (a.active == nullptr || a.active[b] != 0);
  };
  auto cond_b_raw = [&](int ci) {  // condition ci of this CTA, or -1
    const int b = cg * NC + ci;
    return (b < a.B && ec * ES < E && (a.active == nullptr || a.active[b] != 0)) ? b : -1;
  };
#if defined(__CUDA_ARCH__)
  // looked up once: the flags live in global memory and these predicates guard every phase
  const bool tv_mine = team_valid_raw(w.team);
  const int cb0 = cond_b_raw(0);
  auto team_valid = [&](int) { return tv_mine; };
  auto cond_b = [&](int ci) { return ci == 0 ? cb0 : cond_b_raw(ci); };
#else
  auto team_valid = team_valid_raw;
  auto cond_b = cond_b_raw;
#endif
  {
    bool any = false;
    for (int ci = 0; ci < NC; ++ci) any = any || cond_b(ci) >= 0;
    if (!any) return;  // CTA-uniform
  }
  // L2 prefetch of the next C_t: the 128-byte lines of the matrix are dealt to the ES teams that share the condition
  constexpr int PF_LINES = (P * P * (int)sizeof(R) + 127) / 128;
#if defined(__CUDA_ARCH__)
  const int pf_lo = ((w.team % ES) * PF_LINES) / ES, pf_hi = ((w.team % ES + 1) * PF_LINES) / ES;
#else
  const int pf_lo = 0, pf_hi = 0;  // hint only
#endif
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
    LANE_FOR(one, 1) tp[D::oScal + 0] = has_eta ? R(1) / a.eta[PIDX(tm)] : R(1);
  }

  // =========================================================================================
  // Backward pass
  // =========================================================================================
  if (opt_bwd) {
    // async staging of step t: F_ext = [F_t | f_t | 0] (rows DX..DX4-1 stay zero), C_t, C_kl_t, c_t, c_kl_t
    auto stage_bwd = [&](int t) {
      for (int ci = 0; ci < NC; ++ci) {
        const int b = cond_b(ci);
        if (b < 0) continue;
        R* st = stages + (size_t)ci * D::STAGE;
        const size_t bt = (size_t)b * T + t;
        // rows of [F_t | f_t] and [C_kl_t | c_kl_t]: 16-byte copies where the row length is even
        if (P % 2 == 0 && a.aligned16) {
          constexpr int HP = (P + 1) / 2;
          CTA_FOR(idx, DX * HP) {
            const int k = idx / HP, j = 2 * (idx % HP);  // compile-time divisors
            cp_async_pair(st + D::oFx + k * PP + j, a.F + bt * szF + (size_t)k * P + j);
          }
          if (has_ckl) {
            CTA_FOR(idx, P * HP) {
              const int i = idx / HP, j = 2 * (idx % HP);
              cp_async_pair(st + D::oCk + i * D::PKL + j, a.C_kl + bt * szC + (size_t)i * P + j);
            }
          }
        } else {
          CTA_FOR(idx, DX * P) {
            const int k = idx / P, j = idx % P;
            cp_async(st + D::oFx + k * PP + j, a.F + bt * szF + idx);
          }
          if (has_ckl) {
            CTA_FOR(idx, P * P) {
              const int i = idx / P, j = idx % P;
              cp_async(st + D::oCk + i * D::PKL + j, a.C_kl + bt * szC + idx);
            }
          }
        }
        CTA_FOR(k, DX) cp_async(st + D::oFx + k * PP + P, a.f + bt * DX + k);
        if (has_cklv) {
          CTA_FOR(i, P) cp_async(st + D::oCk + i * D::PKL + P, a.c_kl + bt * P + i);
        }
      }
    };
    stage_bwd(T - 1);
    // terminal value V = C_ext[T,:DX,:DX], v = c_ext[T,:DX] (lqg.py:151-152)
    TEAMS_DO(tm) {
      if (!team_valid(tm)) continue;
      R* tp = TEAMP(tm);
      const R ie = has_eta ? R(1) / a.eta[PIDX(tm)] : R(1);
      const size_t bT = (size_t)team_b(tm) * (T + 1) + T;
      LANE_FOR(idx, DX * DX) {
        const int i = idx / DX, j = idx % DX;
        tp[D::oM + i * PP + j] = ldg(a.C + bT * szC + (size_t)i * P + j) * ie;
      }
      LANE_FOR(i, DX) tp[D::oM + i * PP + P] = ldg(a.c + bT * P + i) * ie;
      LANE_FOR(j, DU) tp[D::oLds + j] = R(1);
    }
    cp_async_wait_all();
    w.cta_sync();
    PHASE_MARK(12);

    for (int t = T - 1; t >= 0; --t) {
      TEAMS_DO(tm) {
        if (!team_valid(tm)) continue;
        R* tp = TEAMP(tm);
        const R* st = STAGEP(tm);
        const int b = team_b(tm);
        if (t > 0) {  // pull next step's C rows into L2 while this step computes: this team's share of the lines
          for (int ln = pf_lo + w.lane; ln < pf_hi; ln += 32)
            prefetch_l2(a.C + ((size_t)b * (T + 1) + t - 1) * szC + (size_t)ln * (128 / sizeof(R)));
        }
        // C_t entries (i, j), (i, j+1) of the Q fragments: raw loads issued right before the B2 product (50 DMMA at
        // (20,6): long enough to cover an L2 hit), consumed by its epilogue
        const R ie = tp[D::oScal + 0];
        const R* Cg = a.C + ((size_t)b * (T + 1) + t) * szC;
        const R* cg = a.c + ((size_t)b * (T + 1) + t) * P;
        const bool has_kl = has_ckl;
        auto pre = [&](int i, int j, R* pv) {
          if (i >= P || i > j + 1) return;
          if (j + 1 < P) {
            ldg2(Cg + i * P + j, P % 2 == 0, pv[0], pv[1]);
          } else {
            if (j < P) pv[0] = ldg(Cg + i * P + j);
            const int h = j == P ? 0 : 1;  // which of the two columns is the vector column P
            if (j == P || j + 1 == P) pv[h] = ldg(cg + i);
          }
        };
        // ---- B1: W_ext = V [F | f] (+ v in column P) ----------------------------------------------------
        warp_mma<R, D::NBX, D::NBP>(w, tp + D::oM, PP, st + D::oFx, PP, DX4 / 4, [&](int i, int j, R v0, R v1) {
          if (i < DX && j < PP) {
            if (j == P) v0 += tp[D::oM + i * PP + P];
            if (j + 1 == P) v1 += tp[D::oM + i * PP + P];
            st2s(tp + D::oR + i * PP + j, v0, v1, i);
          }
        });
        w.team_sync();
        PHASE_MARK(0);
        // ---- B2: Q_ext = C_ext / eta + C_kl + gamma [F|f]^T W_ext, upper blocks; column P is q.  Only the xu block is
        //      mirrored (Qux feeds the gains and the value update); Qxx / Quu are consumed from their upper triangles.
        //      C_kl comes from the staged copy of the condition, C from the registers loaded above. ------------------
        {
          auto epi = [&](int i, int j, R v0, R v1, const R* pv) {
            if (i >= P || i > j + 1 || j > P) return;
            R k0 = R(0), k1 = R(0);
            if (has_kl) ld2s(st + D::oCk + i * D::PKL + j, k0, k1, i);  // column P holds c_kl, beyond it zeros
            const R q0 = fma(gamma, v0, fma(pv[0], ie, k0)), q1 = fma(gamma, v1, fma(pv[1], ie, k1));
            if (i <= j && j + 1 <= P) {
              st2s(tp + D::oM + i * PP + j, q0, q1, i);
            } else {
              if (i <= j) tp[D::oM + i * PP + j] = q0;
              if (j + 1 <= P) tp[D::oM + i * PP + j + 1] = q1;
            }
            if (i < DX) {  // mirror Qxu -> Qux
              if (j >= DX && j < P) tp[D::oM + j * PP + i] = q0;
              if (j + 1 >= DX && j + 1 < P) tp[D::oM + (j + 1) * PP + i] = q1;
            }
          };
          warp_mma_tri<R, D::NBP, D::NBP, 2>(w, st + D::oFx, PP, tp + D::oR, PP, DX4 / 4, pre, epi);
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
            // log det: the diagonals are multiplied up over 8 steps and the logarithm taken once (a double-precision
            // log costs ~80 instructions); oDel holds the per-lane sums of the backward pass
            R pr = tp[D::oLds + ln] * dsel;
            if ((t & 7) == 0) {
              tp[D::oDel + ln] += log(pr);
              pr = R(1);
            }
            tp[D::oLds + ln] = pr;
          }
          if (ln == 31 && bad && tp[D::oScal + 4] == R(0)) tp[D::oScal + 4] = R(t + 1);
          // pol_inv_covar = Quu (lqg.py:168): one entry per lane
          for (int e = ln; e < DU * DU; e += 32) {
            const int i = e / DU, j = e % DU;
            a.pinv[pt * szS + e] = tp[D::oM + (DX + (i < j ? i : j)) * PP + DX + (i < j ? j : i)];
          }
          // One right-hand side per lane, two triangular substitutions each (scipy.linalg.solve(assume_a="pos"),
          // lqg.py:165-167):  column ln of Qux -> column ln of K;  q_u -> k;  column c of -I -> column c of
          // pol_covar = Quu^-1.  out = -Quu^-1 x in every case.
#pragma unroll
          for (int c0 = 0; c0 < DX + 1 + DU; c0 += 32) {
            const int col = c0 + ln;
            if (col < DX + 1 + DU) {
              R x[DU];
              const int mcol = col < DX ? col : P;
#pragma unroll
              for (int b2 = 0; b2 < DU; ++b2)
                x[b2] = col <= DX ? tp[D::oM + (DX + b2) * PP + mcol] : (b2 == col - DX - 1 ? R(-1) : R(0));
#pragma unroll
              for (int i = 0; i < DU; ++i) {  // L y = x
                R acc = x[i];
#pragma unroll
                for (int k = 0; k < i; ++k) acc = fma(-Lm[i][k], x[k], acc);
                x[i] = acc * idg[i];
              }
#pragma unroll
              for (int i = DU - 1; i >= 0; --i) {  // L^T z = y
                R acc = x[i];
#pragma unroll
                for (int k = i + 1; k < DU; ++k) acc = fma(-Lm[k][i], x[k], acc);
                x[i] = acc * idg[i];
              }
#pragma unroll
              for (int a2 = 0; a2 < DU; ++a2) {
                const R val = -x[a2];
                if (col <= DX) tp[D::oK + a2 * KLD + col] = val;
                if (col < DX) a.K[pt * szK + a2 * DX + col] = val;
                else if (col == DX) a.k[pt * DU + a2] = val;
                else a.pcov[pt * szS + a2 * DU + (col - DX - 1)] = val;
              }
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
              ld2s(tp + D::oM + i * PP + j, m0, m1, i);
              m0 += v0; m1 += v1;
              st2s(tp + D::oM + i * PP + j, m0, m1, i);
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
          warp_mma_tri<R, D::NBX, D::NBX1, 1>(w, tp + D::oM + DX * PP, PP, tp + D::oK, KLD, DU4 / 4, [](int, int, R*) {},
                                              [&](int i, int j, R v0, R v1, const R*) { epi(i, j, v0, v1); });
        }
        w.team_sync();
        PHASE_MARK(4);
      }
      cp_async_wait_all();
      w.cta_sync();
      PHASE_MARK(5);
    }
    TEAMS_DO(tm) {  // sum_t log det chol(Quu_t)  (t = 0 flushed the running products)
      if (!team_valid(tm)) continue;
      R* tp = TEAMP(tm);
      LANE_FOR(one, 1) {
        R ld = R(0);
#pragma unroll
        for (int j = 0; j < DU; ++j) ld += tp[D::oDel + j];
        tp[D::oScal + 3] = ld;
      }
      w.team_sync();
    }
  }

  // =========================================================================================
  // Forward pass with KL and expected cost
  // =========================================================================================
  if (opt_fwd) {
    const R reg = a.fwd_reg;
    const bool do_kl = FULL || (a.flags & ILQR_KL) != 0, do_cost = FULL || (a.flags & ILQR_COST) != 0;
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
        if (DX % 2 == 0 && a.aligned16) {
          CTA_FOR(idx, DX * (DX / 2)) {
            const int i = idx / (DX / 2), j = 2 * (idx % (DX / 2));
            cp_async_pair(st + D::oD + i * D::DXE + j, a.dyn_covar + bt * DX * DX + (size_t)i * DX + j);
          }
        } else {
          CTA_FOR(idx, DX * DX) {
            const int i = idx / DX, j = idx % DX;
            cp_async(st + D::oD + i * D::DXE + j, a.dyn_covar + bt * DX * DX + idx);
          }
        }
      }
    };
    auto stage_policy = [&](int tm, int t) {  // K_t^T, k_t, pol_covar_t of this team's problem (async)
      R* tp = TEAMP(tm);
      const size_t pt = PIDX(tm) * T + t;
#pragma unroll
      for (int a2 = 0; a2 < DU; ++a2) {
        LANE_FOR(k, DX) cp_async(tp + D::oK + k * KTL + a2, a.K + pt * szK + (size_t)a2 * DX + k);
      }
      LANE_FOR(i, DU) cp_async(tp + D::kv_off(i), a.k + pt * DU + i);
      LANE_FOR(idx, DU * DU) cp_async(tp + D::pc_off(idx), a.pcov + pt * szS + idx);
    };
    stage_fwd(0);
    TEAMS_DO(tm) {
      if (!team_valid(tm)) continue;
      R* tp = TEAMP(tm);
      const int b = team_b(tm);
      LANE_FOR(idx, D::MSZ) tp[D::oM + idx] = R(0);
      w.team_sync();
      LANE_FOR(idx, DX * DX) tp[D::oM + (idx / DX) * PP + idx % DX] = ldg(a.x0S + (size_t)b * DX * DX + idx);
      LANE_FOR(i, DX) {
        const R m0 = ldg(a.x0m + (size_t)b * DX + i);
        tp[D::oM + i * PP + P] = m0;
        tp[D::oRow + i] = m0;
      }
      stage_policy(tm, 0);
    }
    cp_async_wait_all();
    w.cta_sync();
    PHASE_MARK(12);

    for (int t = 0; t < T; ++t) {
      // ---- first half of step t: team-private data only (policy of step t was fetched during step t-1); the shared
      //      inputs of step t (F_t^T, dyn_covar_t, previous policy) are still in flight ---------------------------------
      TEAMS_DO(tm) {
        if (!team_valid(tm)) continue;
        R* tp = TEAMP(tm);
        const int b = team_b(tm);
        const R* Cg = a.C + ((size_t)b * (T + 1) + t) * szC;
        const R* cg = a.c + ((size_t)b * (T + 1) + t) * P;
        if (do_cost && t + 1 < T) {  // pull next step's C rows into L2: this team's share of the lines
          for (int ln = pf_lo + w.lane; ln < pf_hi; ln += 32)
            prefetch_l2(a.C + ((size_t)b * (T + 1) + t + 1) * szC + (size_t)ln * (128 / sizeof(R)));
        }
#if defined(__CUDA_ARCH__)
        // cost weights of this step: lane j holds column j of C_t (P values, row-wise coalesced loads), issued now,
        // consumed after F2
        static_assert(P <= 32, "cost phase: one column of C per lane");
        R creg[P], cvec = R(0), cdiag = R(0);
        if (do_cost) {
          const int jc = w.lane < P ? w.lane : 0;
#pragma unroll
          for (int k = 0; k < P; ++k) creg[k] = ldg(Cg + k * P + jc);
          cvec = ldg(cg + jc);
          cdiag = ldg(Cg + jc * P + jc);
        }
#endif
        cp_async_wait_group<1>();  // this lane's share of the policy of step t has landed
        w.team_sync();
        // ---- F1: [Sigma_ux | mu_u] = K [Sigma_xx | mu_x] (+ k) (lqg.py:227-230) -----------------------------------
        warp_mma<R, 1, D::NBP>(w, tp + D::oK, KTL, tp + D::oM, PP, DX4 / 4, [&](int a2, int j, R v0, R v1) {
          if (a2 < DU) {
            if (j + 1 < DX) {
              st2s(tp + D::oM + (DX + a2) * PP + j, v0, v1, a2);
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
                  const R mu_u = acc + tp[D::kv_off(a2)];
                  tp[D::oM + (DX + a2) * PP + P] = mu_u;
                  tp[D::oRow + DX + a2] = mu_u;
                }
              }
            }
          }
        });
        w.team_sync();
        PHASE_MARK(6);
        // ---- F2: Sigma_uu = Sigma_ux K^T + pol_covar (upper, mirrored) (lqg.py:231) ------------------------------
        warp_mma<R, 1, 1>(w, tp + D::oM + DX, PP, tp + D::oK, KTL, DX4 / 4, [&](int a2, int j, R v0, R v1) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int bb = j + h;
            if (a2 <= bb && bb < DU) {
              const R val = (h ? v1 : v0) + tp[D::pc_off(a2 * DU + bb)];
              tp[D::oM + (DX + a2) * PP + DX + bb] = val;
              tp[D::oM + (DX + bb) * PP + DX + a2] = val;
            }
          }
        });
        w.team_sync();
        PHASE_MARK(7);
        // ---- expected cost of step t: 1/2 sum_kj C_kj (Sigma_kj + mu_k mu_j [+ reg]) + c^T mu (quadratic_costs.py:106-119)
        if (do_cost) {
#if defined(__CUDA_ARCH__)
          // lane j: sum_k C_kj Sigma_kj and sum_k C_kj mu_k  (Sigma row k at consecutive addresses, mu_k broadcast)
          R acc1 = R(0), acc2 = R(0);
          const R* mrow = tp + D::oM + w.lane;
#pragma unroll
          for (int k = 0; k < P; k += 2) {
            R mk0, mk1;
            ld2(tp + D::oRow + k, mk0, mk1);
            acc1 = fma(creg[k], mrow[k * PP], acc1);
            acc2 = fma(creg[k], mk0, acc2);
            if (k + 1 < P) {
              acc1 = fma(creg[k + 1], mrow[(k + 1) * PP], acc1);
              acc2 = fma(creg[k + 1], mk1, acc2);
            }
          }
          R cost = R(0);
          if (w.lane < P) {
            const R mj = tp[D::oRow + w.lane];
            cost = fma(R(0.5), fma(mj, acc2, fma(reg, cdiag, acc1)), mj * cvec);
          }
          team_accumulate(w, tp + D::oScal + 2, cost);
#else
          R cost = R(0);
          for (int k = 0; k < P; ++k)
            for (int j = 0; j < P; ++j) {
              R sv = fma(tp[D::oRow + k], tp[D::oRow + j], tp[D::oM + k * PP + j]);
              if (k == j) sv += reg;
              cost = fma(Cg[k * P + j], sv, cost);
            }
          cost *= R(0.5);
          for (int i = 0; i < P; ++i) cost = fma(tp[D::oRow + i], cg[i], cost);
          tp[D::oScal + 2] += cost;
#endif
        }
        if (want_tcov) {
          LANE_FOR(idx, P * P) {
            const int i = idx / P, j = idx % P;
            a.tcov[(PIDX(tm) * (T + 1) + t) * szC + idx] = tp[D::oM + i * PP + j] + (i == j ? reg : R(0));
          }
        }
        if (want_tmean) {
          LANE_FOR(i, P) a.tmean[(PIDX(tm) * (T + 1) + t) * P + i] = tp[D::oM + i * PP + P];
        }
        PHASE_MARK(8);
      }
      cp_async_wait_all();
      w.cta_sync();  // the shared inputs of step t are in place
      PHASE_MARK(9);
      // ---- second half of step t --------------------------------------------------------------------------------
      TEAMS_DO(tm) {
        if (!team_valid(tm)) continue;
        R* tp = TEAMP(tm);
        const R* st = STAGEP(tm);
        if (do_kl) {  // delta_u = (K_prev mu_x + k_prev) - mu_u  (lqg.py:294)
          warp_mma<R, 1, 1>(w, st + oPK, KTL, tp + D::oM + 8 * D::BP, PP, DX4 / 4, [&](int a2, int j, R v0, R v1) {
            if (a2 < DU) {
              if (8 * D::BP + j == P) tp[D::oDel + a2] = v0 + st[oPk + a2] - tp[D::oM + (DX + a2) * PP + P];
              if (8 * D::BP + j + 1 == P) tp[D::oDel + a2] = v1 + st[oPk + a2] - tp[D::oM + (DX + a2) * PP + P];
            }
          });
        }
        // ---- F3: Rt_ext = ([F] [Sigma | mu])^T (PE x DX) (lqg.py:236-240) -----------------------------------------
        warp_mma<R, D::NBP, D::NBX>(w, tp + D::oM, PP, st + D::oFx, DXP, P4 / 4, [&](int j, int i, R v0, R v1) {
          if (j < PE) {
            if (i + 1 < DX) st2s(tp + D::oR + j * DXP + i, v0, v1, j);
            else if (i < DX) tp[D::oR + j * DXP + i] = v0;
          }
        });
        w.team_sync();
        PHASE_MARK(10);
        // ---- F4: Sigma'_xx = (F Sigma) F^T + dyn_covar (upper blocks, mirrored); mu' = F mu + f; scalars -------------
        {
          auto epi = [&](int i, int j, R v0, R v1) {
            if (i >= DX || i > j + 1 || j >= DX) return;
            if (i <= j && j + 1 < DX) {
              R d0, d1;
              ld2s(st + D::oD + i * D::DXE + j, d0, d1, i);
              const R s0 = v0 + d0, s1 = v1 + d1;
              st2s(tp + D::oM + i * PP + j, s0, s1, i);
              tp[D::oM + j * PP + i] = s0;
              tp[D::oM + (j + 1) * PP + i] = s1;
            } else {
              if (i <= j) {
                const R s0 = v0 + st[D::oD + i * D::DXE + j];
                tp[D::oM + i * PP + j] = s0;
                tp[D::oM + j * PP + i] = s0;
              }
              if (j + 1 < DX) {
                const R s1 = v1 + st[D::oD + i * D::DXE + j + 1];
                tp[D::oM + i * PP + j + 1] = s1;
                tp[D::oM + (j + 1) * PP + i] = s1;
              }
            }
          };
          warp_mma_tri<R, D::NBX, D::NBX, 1>(w, tp + D::oR, DXP, st + D::oFx, DXP, P4 / 4, [](int, int, R*) {},
                                             [&](int i, int j, R v0, R v1, const R*) { epi(i, j, v0, v1); });
        }
        LANE_FOR(i, DX) {
          const R mu_n = tp[D::oR + P * DXP + i] + st[oFv + i];
          tp[D::oM + i * PP + P] = mu_n;
          tp[D::oRow + i] = mu_n;
        }
        if (do_kl) {
          LANE_FOR(ln, DU) {  // row ln of  tr(Lam_prev Sigma_new) + delta^T Lam_prev delta  (lqg.py:295-304)
            R tr = R(0), row = R(0);
#pragma unroll
            for (int j = 0; j < DU; ++j) {
              const R lij = st[oPL + ln * DU + j];
              tr = fma(lij, tp[D::pc_off(j * DU + ln)], tr);
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
        cp_async_commit();
        PHASE_MARK(11);
      }
      w.cta_sync();  // everyone is done with the shared inputs of step t: fetch those of step t+1 behind F1 / F2 / cost
      if (t + 1 < T) stage_fwd(t + 1);
      cp_async_commit();
      PHASE_MARK(13);
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
      if (want_tcov) {
        LANE_FOR(idx, P * P) {
          const int i = idx / P, j = idx % P;
          a.tcov[(PIDX(tm) * (T + 1) + T) * szC + idx] = (i < DX && j < DX) ? tp[D::oM + i * PP + j] : R(0);
        }
      }
      if (want_tmean) {
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
      if (FULL || ((a.flags & ILQR_FORWARD) && (a.flags & ILQR_COST) && a.cost)) {
        R acc = R(0);
        for (int i = 0; i < 32; ++i) acc += tp[D::oRow + i];
        a.cost[pi] = tp[D::oScal + 2] + acc + ldg(a.cc + (size_t)team_b(tm) * (T + 1) + T);
      }
      if (FULL || ((a.flags & ILQR_FORWARD) && (a.flags & ILQR_KL) && a.kl)) a.kl[pi] = (tp[D::oScal + 1] + tp[D::oScal + 3]) / R(T);
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

// Teams (warps) per CTA.  A CTA holds ES eta candidates of NC = G / ES conditions and one stage area per condition.
// Pick the G that puts the most warps on an SM (shared memory and 128 registers per thread allow 16); ties go to the
// smaller CTA: two 8-warp CTAs drift apart, so the tensor-heavy products of one overlap the factorisation / epilogue /
// staging phases of the other (one 16-warp CTA runs its warps in lock-step).
template <typename R, int DX, int DU>
inline int ilqr_warp_plan(int B, int E, size_t smem_cap, int force_G, int& G, int& ES, int& grid, size_t& bytes) {
  const size_t sm_total = 228 * 1024, cta_reserved = 1024;
  int best_g = 0, best_warps = 0, best_es = 0;
  // E = 1 (the Brent steps of ILQR.optimize): every warp is a different condition with its own stage area, nothing is
  // shared, so one-warp CTAs that never meet at a CTA barrier win (config 3, 14 steps: 63.9 ms; five-warp CTAs with one
  // more warp per SM 75.8 ms; one ten-warp CTA 101 ms)
  for (int g = force_G > 0 ? force_G : (E == 1 ? 1 : 16); g >= 1; --g) {
    const int es = E < g ? E : g, gg = (g / es) * es;
    if (gg < 1) continue;
    const size_t by = ilqr_warp_bytes<R, DX, DU>(gg, gg / es);
    if (by > smem_cap) continue;
    int ctas = (int)(sm_total / (by + cta_reserved));
    if (ctas * gg > 16) ctas = 16 / gg;  // register file: 128 registers x 32 lanes x 16 warps
    if (gg > 8 && ctas > 1) ctas = 1;    // the large-CTA kernel is built for one CTA per SM
    const int warps = ctas * gg;
    // ties: CTAs of at most 8 warps (several per SM, out of phase; measured 56 vs 60 ms at config 3), then more eta
    // candidates per staged condition (fewer copies of the shared inputs), then the smaller CTA
    const bool small = gg <= 8, best_small = best_g <= 8;
    const bool better = warps != best_warps ? warps > best_warps
                        : small != best_small ? small
                        : es != best_es ? es > best_es
                                        : gg < best_g;
    if (best_g == 0 || better) { best_warps = warps; best_g = gg; best_es = es; }
    if (force_G > 0) break;
  }
  if (best_g < 1) return DONK_SMEM_LIMIT;
  G = best_g;
  ES = E < G ? E : G;
  G = (G / ES) * ES;
  bytes = ilqr_warp_bytes<R, DX, DU>(G, G / ES);
  const int NC = G / ES, n_ech = (E + ES - 1) / ES;
  grid = ((B + NC - 1) / NC) * n_ech;
  return DONK_OK;
}

}  // namespace donk
