// donk_b200 — iLQR step, CTA-per-problem variant for large state dimensions (BASELINE config 5: dX = 64, dU = 16).
//
// Same math and reference lines as ilqr.cuh / ilqr_warp.cuh (ILQR.step, donk/traj_opt/lqg.py:55-75; backward
// :126-182; forward :185-244; kl_divergence_action :276-306; expected_costs quadratic_costs.py:106-119).
//
// One (condition, eta) problem per CTA of G warps; the whole working set lives in shared memory for all T steps
// (p = 80: Q/V/Sigma 54 KB, W / (F Sigma)^T 44 KB, F_t double buffered 2 x 43 KB, gains 9 KB  ->  ~205 KB, one
// CTA per SM).  Every product of a step runs on the FP64 tensor path (mma.sync.m8n8k4.f64) as register tiles of
// RA x RB fragments; the tiles of a product are dealt to the warps by a longest-tile-first schedule computed once
// per CTA (compile-time dimensions, triangular products have ragged tiles).  Structure of a step:
//   backward  B1  W_ext = V [F | f]                                  all warps                 barrier
//             B2  Q_ext = C/eta + C_kl + gamma F_ext^T W_ext (upper) warps 1..: Qxx, Qxu, q_x
//                 warp 0: Quu, q_u, then chol(Quu), L^-1, Quu^-1     (the serial chain hides behind the product)
//                                                                                              barrier
//             B4  [K | k] = -Quu^-1 [Qux | q_u]                      all warps                 barrier
//             B5  [V | v] = [Qxx | q_x] + Qxu [K | k] (upper)        all warps                 barrier
//   forward   F1  [Sigma_ux | mu_u] = K [Sigma_xx | mu_x] (+ k)      all warps                 barrier
//             F2  Sigma_uu = Sigma_ux K^T + pol_covar; KL delta      few warps each            barrier
//             F3  Rt_ext = (F [Sigma | mu])^T; expected cost; KL     all warps                 barrier
//             F4  Sigma'_xx = (F Sigma) F^T + dyn_covar; mu'         all warps                 barrier
// F_t is staged with cp.async one step ahead into the other half of a double buffer (it is the operand of the
// two large products of either pass, so a single buffer could only be refilled behind the small ones);
// C_t, C_kl_t, dyn_covar_t enter through the epilogues straight from global memory / L2 (prefetched one step
// ahead).  As in the warp kernel the vectors ride in column P of the matrices, and all operands keep their natural
// orientation (both fragment orientations are bank-conflict free at leading dimensions with ld % 16 in {4, 12}).
// Dimensions: DX % 8 == 0 and DU % 8 == 0 (whole 8 x 8 fragments; the only ragged block is the vector column).
#pragma once
#include "common.cuh"
#include "dmma_tiles.cuh"
#include "ilqr.cuh"

namespace donk {

// Per-phase cycle counters (profiling builds only: profiles/prof_phases_coop.py compiles a separate library with
// -DDONK_PHASE_PROF; the shipped library has no trace of it).  Thread 0 of every CTA adds the time between marks.
#if defined(DONK_PHASE_PROF) && !defined(DONK_EMU)
__device__ unsigned long long donk_coop_phase_cycles[32];
#endif
#if defined(DONK_PHASE_PROF) && defined(__CUDA_ARCH__)
#define CPHASE_INIT long long cph_t0 = clock64();
#define CPHASE_MARK(id)                                                                          \
  do {                                                                                           \
    const long long cph_now = clock64();                                                         \
    if (w.tid == 0) atomicAdd(&donk_coop_phase_cycles[id], (unsigned long long)(cph_now - cph_t0)); \
    cph_t0 = cph_now;                                                                            \
  } while (0)
#else
#define CPHASE_INIT
#define CPHASE_MARK(id)
#endif

template <int DX, int DU>
struct CoopDims {
  static_assert(DX % 8 == 0 && DU % 8 == 0 && DU <= 32, "cooperative iLQR kernel: dX, dU multiples of 8, dU <= 32");
  static constexpr int P = DX + DU, PE = P + 1;
  static constexpr int PP = ld_pad(PE);       // leading dimension of M, W_ext, F_ext
  static constexpr int DXP = ld_pad(DX);      // ... of Rt_ext
  static constexpr int KLD = ld_pad(DX + 1);  // ... of Kn = [K | k] (DU x (DX + 1))
  static constexpr int SLD = ld_pad(DU);      // ... of Quu^-1, L, L^-1
  static constexpr int NBX = DX / 8, NBU = DU / 8, NBP = P / 8 + 1;  // 8-column blocks: state, action, all + vector
  static constexpr int BV = P / 8;            // the block whose first column is the vector column P
  // shared memory (reals); every region starts 16-byte aligned and has 8 spare reals behind it (fragment over-reads)
  static constexpr int MSZ = P * PP + 8;
  static constexpr int RSZ = (DX * PP > PE * DXP ? DX * PP : PE * DXP) + 8;
  static constexpr int FSZ = DX * PP + 8;
  static constexpr int KSZ = DU * KLD + 8;
  static constexpr int SSZ = DU * SLD + 8;
  static constexpr int oM = 0, oR = oM + MSZ, oF0 = oR + RSZ, oF1 = oF0 + FSZ, oK = oF1 + FSZ;
  static constexpr int oSinv = oK + KSZ, oL = oSinv + SSZ, oLi = oL + SSZ;
  static constexpr int oKp = oLi + SSZ;             // forward: [K_prev | k_prev] (DU x KLD), natural orientation
  static constexpr int oLp = oKp + KSZ;             // forward: Lam_prev (DU x SLD), [DU*SLD]: prev_hld
  static constexpr int oMu = oLp + SSZ;             // compact copy of mu (P)
  static constexpr int oDel = oMu + P + 8;          // KL delta (DU), idg (DU)
  static constexpr int oWp = oDel + 2 * DU + 8;     // per-warp partial sums: [32] cost, [32] kl
  static constexpr int oScal = oWp + 64;            // 0 1/eta 1 kl 2 cost 3 logdet 4 notpd
  static constexpr int REALS = oScal + 8;
  // tile shapes (fragments) of the products: rows always whole tiles, the last tile of a block row may be narrower
  static constexpr int B1_RA = 2, B1_RB = 6;   // W_ext: NBX x NBP
  static constexpr int B2_RA = 2, B2_RB = 3;   // Q_ext state rows: NBX x NBP, upper
  static constexpr int B4_RA = 1, B4_RB = 3;   // K: NBU x NBX (the k column: one extra fragment per block row)
  static constexpr int B5_RA = 2, B5_RB = 3;   // V_ext: NBX x (NBX + 1), upper
  static constexpr int F1_RA = 1, F1_RB = 3;   // Sigma_ux: NBU x NBX (mu_u: one extra fragment per block row)
  static constexpr int F3_RA = 2, F3_RB = 6;   // F [Sigma | mu]: NBX x NBP (stored transposed)
  static constexpr int F4_RA = 2, F4_RB = 3;   // Sigma'_xx: NBX x NBX, upper
  static_assert(NBX % 2 == 0, "cooperative iLQR kernel: dX must be a multiple of 16 (two block rows per tile)");
  enum { PH_B1, PH_B2, PH_B4, PH_B5, PH_F1, PH_F3, PH_F4, NPH };
};

template <int DX, int DU>
DONK_HD void coop_schedule_all(CoopSched& s, int G) {
  using D = CoopDims<DX, DU>;
  coop_schedule(s, D::PH_B1, G, D::B1_RA, D::B1_RB, D::NBX, D::NBP, false, false);
  coop_schedule(s, D::PH_B2, G, D::B2_RA, D::B2_RB, D::NBX, D::NBP, true, true);
  coop_schedule(s, D::PH_B4, G, D::B4_RA, D::B4_RB, D::NBU, D::NBX, false, false);
  coop_schedule(s, D::PH_B5, G, D::B5_RA, D::B5_RB, D::NBX, D::NBX + 1, true, false);
  coop_schedule(s, D::PH_F1, G, D::F1_RA, D::F1_RB, D::NBU, D::NBX, false, false);
  coop_schedule(s, D::PH_F3, G, D::F3_RA, D::F3_RB, D::NBX, D::NBP, false, false);
  coop_schedule(s, D::PH_F4, G, D::F4_RA, D::F4_RB, D::NBX, D::NBX, true, false);
}

// every tile must have found a slot (12 per warp and product); checked on the host by the launcher
template <int DX, int DU>
inline bool coop_schedule_fits(int G) {
  using D = CoopDims<DX, DU>;
  CoopSched s;
  coop_schedule_all<DX, DU>(s, G);
  struct { int ph, RA, RB, na, nb; bool tri; } prod[] = {
      {D::PH_B1, D::B1_RA, D::B1_RB, D::NBX, D::NBP, false}, {D::PH_B2, D::B2_RA, D::B2_RB, D::NBX, D::NBP, true},
      {D::PH_B4, D::B4_RA, D::B4_RB, D::NBU, D::NBX, false}, {D::PH_B5, D::B5_RA, D::B5_RB, D::NBX, D::NBX + 1, true},
      {D::PH_F1, D::F1_RA, D::F1_RB, D::NBU, D::NBX, false}, {D::PH_F3, D::F3_RA, D::F3_RB, D::NBX, D::NBP, false},
      {D::PH_F4, D::F4_RA, D::F4_RB, D::NBX, D::NBX, true}};
  for (auto& q : prod) {
    int tiles = 0;
    for (int ra0 = 0; ra0 < q.na; ra0 += q.RA)
      for (int rb0 = q.tri ? ra0 : 0; rb0 < q.nb; rb0 += q.RB)
        if (coop_tile_cost(ra0, rb0, q.RA, q.RB, q.na, q.nb, q.tri)) ++tiles;
    int placed = 0;
    for (int g = 0; g < G; ++g) placed += s.n[q.ph][g];
    if (tiles > 192 || placed != tiles) return false;
  }
  return true;
}

template <typename R, int DX, int DU>
DONK_HD bool ilqr_coop_is_full(const IlqrArgs<R>& a) {
  return a.flags == (ILQR_BACKWARD | ILQR_FORWARD | ILQR_KL | ILQR_COST) && a.C_kl && a.c_kl && a.eta && a.kl &&
         a.cost && !a.tcov && !a.tmean;
}

template <typename R, int DX, int DU>
inline size_t ilqr_coop_bytes() {
  return (size_t)CoopDims<DX, DU>::REALS * sizeof(R) + sizeof(CoopSched) + 16;
}

template <typename R, int DX, int DU, bool FULL>
DONK_D void ilqr_coop_cta(const WCtx& w, const IlqrArgs<R>& a, int cta, R* smem) {
  using D = CoopDims<DX, DU>;
  constexpr int P = D::P, PE = D::PE, PP = D::PP, DXP = D::DXP, KLD = D::KLD, SLD = D::SLD;
  constexpr int NBX = D::NBX, NBU = D::NBU, NBP = D::NBP;
  const bool opt_bwd = FULL || (a.flags & ILQR_BACKWARD) != 0, opt_fwd = FULL || (a.flags & ILQR_FORWARD) != 0;
  const bool has_ckl = FULL || a.C_kl != nullptr, has_cklv = FULL || a.c_kl != nullptr, has_eta = FULL || a.eta != nullptr;
  const bool want_tcov = !FULL && a.tcov != nullptr, want_tmean = !FULL && a.tmean != nullptr;
  const int G = w.G, T = a.T, E = a.E;
  const int b = cta / E, e = cta % E;
  if (b >= a.B || (a.active != nullptr && a.active[b] == 0)) return;  // CTA-uniform
  const size_t pi = (size_t)b * E + e;
  const R gamma = a.gamma;
  constexpr size_t szF = (size_t)DX * P, szC = (size_t)P * P, szK = (size_t)DU * DX, szS = (size_t)DU * DU;

  R* M = smem + D::oM;
  R* Rb = smem + D::oR;
  // the two halves of the F_ext double buffer: offsets from the shared-memory base (an array of pointers indexed at run
  // time made the compiler lose the address space: generic LD.E instead of LDS for every B fragment)
  auto Fbuf = [&](int buf) -> R* { return smem + (buf ? D::oF1 : D::oF0); };
  R* Kn = smem + D::oK;
  R* Sinv = smem + D::oSinv;
  R* Lm = smem + D::oL;
  R* Li = smem + D::oLi;
  R* Kp = smem + D::oKp;
  R* Lp = smem + D::oLp;
  R* mu = smem + D::oMu;
  R* del = smem + D::oDel;
  R* idg = del + DU;
  R* wp = smem + D::oWp;
  R* scal = smem + D::oScal;
  CoopSched& sched = *reinterpret_cast<CoopSched*>(smem + ((D::REALS + 1) / 2) * 2);

  // zero everything once: paddings must be exact zeros / finite
  CTA_FOR(idx, D::REALS) smem[idx] = R(0);
  w.cta_sync();
  CTA_FOR(one, 1) {
    coop_schedule_all<DX, DU>(sched, G);
    scal[0] = has_eta ? R(1) / a.eta[pi] : R(1);
  }
  w.cta_sync();
  const R ie = scal[0];

  // [F_t | f_t] -> Fs[buf] (natural orientation, f in column P), asynchronously
  auto stage_F = [&](int t, int buf) {
    const size_t bt = (size_t)b * T + t;
    R* st = Fbuf(buf);
    if (a.aligned16) {
      constexpr int HP = P / 2;
      CTA_FOR(idx, DX * HP) {
        const int k = idx / HP, j = 2 * (idx % HP);
        cp_async_pair(st + k * PP + j, a.F + bt * szF + (size_t)k * P + j);
      }
    } else {
      CTA_FOR(idx, DX * P) {
        const int k = idx / P, j = idx % P;
        cp_async(st + k * PP + j, a.F + bt * szF + idx);
      }
    }
    CTA_FOR(k, DX) cp_async(st + k * PP + P, a.f + bt * DX + k);
  };
  // 128-byte lines of a global array brought into L2 ahead of their use (hint)
  auto prefetch_lines = [&](const R* base, int reals) {
#if defined(__CUDA_ARCH__)
    const int lines = (reals * (int)sizeof(R) + 127) / 128;
    for (int ln = w.tid; ln < lines; ln += w.nthreads) prefetch_l2(base + (size_t)ln * (128 / sizeof(R)));
#else
    (void)base; (void)reals;
#endif
  };

  // =========================================================================================
  // Backward pass
  // =========================================================================================
  CPHASE_INIT
  if (opt_bwd) {
    stage_F(T - 1, (T - 1) & 1);
    {  // terminal value V = C_ext[T,:DX,:DX], v = c_ext[T,:DX] (lqg.py:151-152)
      const size_t bT = (size_t)b * (T + 1) + T;
      CTA_FOR(idx, DX * DX) {
        const int i = idx / DX, j = idx % DX;
        M[i * PP + j] = ldg(a.C + bT * szC + (size_t)i * P + j) * ie;
      }
      CTA_FOR(i, DX) M[i * PP + P] = ldg(a.c + bT * P + i) * ie;
    }
    cp_async_wait_all();
    w.cta_sync();
    CPHASE_MARK(12);

    for (int t = T - 1; t >= 0; --t) {
      const R* Fc = Fbuf(t & 1);
      const R* Cg = a.C + ((size_t)b * (T + 1) + t) * szC;
      const R* cg = a.c + ((size_t)b * (T + 1) + t) * P;
      const R* Kg = has_ckl ? a.C_kl + ((size_t)b * T + t) * szC : nullptr;
      const R* kg = has_cklv ? a.c_kl + ((size_t)b * T + t) * P : nullptr;
      const size_t pt = pi * T + t;
      if (t > 0) {  // next step's inputs: F into the other buffer, C / C_kl into L2
        stage_F(t - 1, (t - 1) & 1);
        prefetch_lines(a.C + ((size_t)b * (T + 1) + t - 1) * szC, P * P);
        if (has_ckl) prefetch_lines(a.C_kl + ((size_t)b * T + t - 1) * szC, P * P);
      }
      CPHASE_MARK(14);
      // ---- B1: W_ext = V [F | f] (+ v in column P) (lqg.py:160-161) ---------------------------------------------
      TEAMS_DO(tm) {
        for (int s = 0; s < sched.n[D::PH_B1][tm]; ++s) {
          const int ra0 = sched.t[D::PH_B1][tm][s][0], rb0 = sched.t[D::PH_B1][tm][s][1];
          tile_mma_n<R, D::B1_RA, D::B1_RB, 0, false, false, false>(
              w, NBP - rb0, false, M + 8 * ra0, PP, Fc + 8 * rb0, PP, DX / 4, [](int, int, R*) {},
              [&](int il, int jl, R v0, R v1, const R*) {
                const int i = 8 * ra0 + il, j = 8 * rb0 + jl;
                if (j > P) return;
                if (j == P) { Rb[i * PP + P] = v0 + M[i * PP + P]; return; }
                st2s(Rb + i * PP + j, v0, v1, i);
              });
        }
      }
      CPHASE_MARK(15);
      w.cta_sync();
      CPHASE_MARK(0);
      // ---- B2: Q_ext = C_ext / eta + C_kl + gamma [F|f]^T W_ext, upper blocks; column P is q; Qxu mirrored into Qux.
      //      Warp 0 computes the action rows (Quu, q_u) and then factors Quu while the others run the state rows.
      {
        auto pre = [&](int i, int j, R* pv) {  // C (pv[0..1]) and C_kl (pv[2..3]) entries of (i, j), (i, j+1)
          if (i > j + 1) return;
          if (j < P) {
            ldg2(Cg + (size_t)i * P + j, true, pv[0], pv[1]);
            if (has_ckl) ldg2(Kg + (size_t)i * P + j, true, pv[2], pv[3]);
          } else if (j == P) {
            pv[0] = ldg(cg + i);
            if (has_cklv) pv[2] = ldg(kg + i);
          }
        };
        auto epi = [&](int i, int j, R v0, R v1, const R* pv) {
          if (i > j + 1 || j > P) return;
          const R q0 = fma(gamma, v0, fma(pv[0], ie, pv[2])), q1 = fma(gamma, v1, fma(pv[1], ie, pv[3]));
          if (j == P) { M[i * PP + P] = q0; return; }
          if (i <= j) st2s(M + i * PP + j, q0, q1, i);
          else M[i * PP + j + 1] = q1;
          if (i < DX && j >= DX) {  // mirror Qxu -> Qux (operand of the gains and of the value update)
            M[j * PP + i] = q0;
            M[(j + 1) * PP + i] = q1;
          }
        };
        TEAMS_DO(tm) {
          if (tm == 0) {
            // action rows: blocks (NBX + ra, NBX + rb), ra <= rb, plus the vector block
            tile_mma<R, NBU, NBU + 1, 4, false, false, true>(
                w, Fc + DX, PP, Rb + DX, PP, DX / 4,
                [&](int il, int jl, R* pv) { pre(DX + il, DX + jl, pv); },
                [&](int il, int jl, R v0, R v1, const R* pv) { epi(DX + il, DX + jl, v0, v1, pv); });
            w.team_sync();
            CPHASE_MARK(1);
            // ---- Quu = L L^T, L^-1, Quu^-1 = L^-T L^-1 (lqg.py:165-166): the one serial chain of a step ---------------
#if defined(__CUDA_ARCH__)
            {
              // Right-looking factorisation with row i of the working matrix in lane i's registers: the pivot and the
              // finished column travel by shuffles, the trailing update of a column is DU - j - 1 independent FMAs.  Per
              // column the dependent path is shuffle -> rsqrt -> multiply -> shuffle -> FMA (~180 cycles); the
              // shared-memory version below took ~500 (two dependent load-FMA chains of length j and a store / warp
              // barrier / load round trip per column).
              const int li = w.lane < DU ? w.lane : DU - 1;  // lanes beyond DU shadow the last row (never stored)
              R arow[DU];
#pragma unroll
              for (int k = 0; k < DU; ++k)
                arow[k] = M[(DX + (k < li ? k : li)) * PP + DX + (k < li ? li : k)];  // Quu is held in its upper triangle
              bool bad = false;
#pragma unroll
              for (int j = 0; j < DU; ++j) {
                const R piv = __shfl_sync(0xffffffffu, arow[j], j);
                bad = bad || !(piv > R(0));
                const R rs = rsqrt(piv);
                const R lij = arow[j] * rs;  // L[i][j] for i >= j (lane j: sqrt(piv))
                arow[j] = lij;
                if (w.lane == j) idg[j] = rs;
#pragma unroll
                for (int k = j + 1; k < DU; ++k) arow[k] = fma(-lij, __shfl_sync(0xffffffffu, lij, k), arow[k]);
              }
              if (w.lane < DU) {
#pragma unroll
                for (int k = 0; k < DU; ++k) Lm[li * SLD + k] = k <= li ? arow[k] : R(0);
              }
              if (w.lane == 0 && bad && scal[4] == R(0)) scal[4] = R(t + 1);
            }
            w.team_sync();
#else
#pragma unroll
            for (int j = 0; j < DU; ++j) {
              LANE_FOR(r, DU - j) {
                const int i = j + r;
                R sj = M[(DX + j) * PP + DX + j], si = M[(DX + j) * PP + DX + i];  // upper triangle of Quu
#pragma unroll
                for (int k = 0; k < j; ++k) {
                  const R ljk = Lm[j * SLD + k];
                  sj = fma(-ljk, ljk, sj);
                  si = fma(-Lm[i * SLD + k], ljk, si);
                }
                const R rs = rsqrt_r(sj);
                if (i == j) {
                  if (!(sj > R(0)) && scal[4] == R(0)) scal[4] = R(t + 1);
                  idg[j] = rs;
                  Lm[j * SLD + j] = sj * rs;
                } else {
                  Lm[i * SLD + j] = si * rs;
                }
              }
              w.team_sync();
            }
#endif
            LANE_FOR(c, DU) {  // column c of L^-1 by forward substitution (lane-private)
              R x[DU];
#pragma unroll
              for (int i = 0; i < DU; ++i) {
                R acc = i == c ? R(1) : R(0);
#pragma unroll
                for (int k = 0; k < i; ++k) acc = fma(-Lm[i * SLD + k], x[k], acc);
                x[i] = i < c ? R(0) : acc * idg[i];
                Li[i * SLD + c] = x[i];
              }
            }
            w.team_sync();
            // Quu^-1 = L^-T L^-1 on the tensor path (upper fragments, mirrored: symmetric by construction); the
            // outputs that are not on the path to the gains (pol_covar, pol_inv_covar, log det) are written in B4
            tile_mma<R, NBU, NBU, 0, false, false, true>(
                w, Li, SLD, Li, SLD, DU / 4, [](int, int, R*) {},
                [&](int i, int j, R v0, R v1, const R*) {
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                    if (i <= j + h) {
                      const R val = h ? v1 : v0;
                      Sinv[i * SLD + j + h] = val;
                      Sinv[(j + h) * SLD + i] = val;
                    }
                  }
                });
            CPHASE_MARK(2);
          }
          if (tm != 0 || G == 1) {
            for (int s = 0; s < sched.n[D::PH_B2][tm]; ++s) {
              const int ra0 = sched.t[D::PH_B2][tm][s][0], rb0 = sched.t[D::PH_B2][tm][s][1];
              tile_mma_n<R, D::B2_RA, D::B2_RB, 4, false, false, true>(
                  w, NBP - rb0, rb0 == ra0, Fc + 8 * ra0, PP, Rb + 8 * rb0, PP, DX / 4,
                  [&](int il, int jl, R* pv) { pre(8 * ra0 + il, 8 * rb0 + jl, pv); },
                  [&](int il, int jl, R v0, R v1, const R* pv) { epi(8 * ra0 + il, 8 * rb0 + jl, v0, v1, pv); });
            }
          }
        }
      }
      w.cta_sync();
      CPHASE_MARK(3);
      // ---- B4: [K | k] = -Quu^-1 [Qux | q_u] (lqg.py:170-172) --------------------------------------------------------
      TEAMS_DO(tm) {
        for (int s = 0; s < sched.n[D::PH_B4][tm]; ++s) {
          const int ra0 = sched.t[D::PH_B4][tm][s][0], rb0 = sched.t[D::PH_B4][tm][s][1];
          tile_mma_n<R, D::B4_RA, D::B4_RB, 0, false, false, false>(
              w, NBX - rb0, false, Sinv + 8 * ra0, SLD, M + DX * PP + 8 * rb0, PP, DU / 4, [](int, int, R*) {},
              [&](int il, int jl, R v0, R v1, const R*) {
                const int a2 = 8 * ra0 + il, j = 8 * rb0 + jl;
                st2s(Kn + a2 * KLD + j, -v0, -v1, a2);
                a.K[pt * szK + (size_t)a2 * DX + j] = -v0;
                a.K[pt * szK + (size_t)a2 * DX + j + 1] = -v1;
              });
        }
        // outputs of the factorisation that nothing on the critical path waits for (lqg.py:167-168; the log-determinant
        // enters the KL divergence): written here, by warps the gain tiles leave idle
        if (tm == (G > 3 ? G - 3 : 0)) {
          LANE_FOR(idx, DU * DU) {
            const int i = idx / DU, j = idx % DU;
            a.pcov[pt * szS + idx] = Sinv[i * SLD + j];
            a.pinv[pt * szS + idx] = M[(DX + (i < j ? i : j)) * PP + DX + (i < j ? j : i)];  // pol_inv_covar = Quu
          }
        }
        if (tm == (G > 4 ? G - 4 : 0)) {
          team_reduce_add(w, scal + 3, DU, [&](int j) { return log(Lm[j * SLD + j]); });  // sum log diag chol(Quu)
        }
        // k = -Quu^-1 q_u: the right-hand side is column P of the action rows of Q; one fragment per block row, dealt to
        // the warps from the back (the schedule fills them from the front)
        for (int ra = 0; ra < NBU; ++ra) {
          if ((G - 1 - ra % G) != tm) continue;
          tile_mma<R, 1, 1, 0, false, false, false>(
              w, Sinv + 8 * ra, SLD, M + DX * PP + P, PP, DU / 4, [](int, int, R*) {},
              [&](int il, int jl, R v0, R, const R*) {
                if (jl == 0) {
                  Kn[(8 * ra + il) * KLD + DX] = -v0;
                  a.k[pt * DU + 8 * ra + il] = -v0;
                }
              });
        }
      }
      w.cta_sync();
      CPHASE_MARK(4);
      // ---- B5: [V | v] = [Qxx | q_x] + Qxu [K | k] (upper blocks, mirrored) (lqg.py:178-180) ---------------------
      TEAMS_DO(tm) {
        for (int s = 0; s < sched.n[D::PH_B5][tm]; ++s) {
          const int ra0 = sched.t[D::PH_B5][tm][s][0], rb0 = sched.t[D::PH_B5][tm][s][1];
          tile_mma_n<R, D::B5_RA, D::B5_RB, 0, false, false, true>(
              w, NBX + 1 - rb0, rb0 == ra0, M + DX * PP + 8 * ra0, PP, Kn + 8 * rb0, KLD, DU / 4, [](int, int, R*) {},
              [&](int il, int jl, R v0, R v1, const R*) {
                const int i = 8 * ra0 + il, j = 8 * rb0 + jl;
                if (j == DX) { M[i * PP + P] += v0; return; }
                if (j > DX || i > j + 1) return;
                if (i <= j) {
                  R m0, m1;
                  ld2s(M + i * PP + j, m0, m1, i);
                  m0 += v0; m1 += v1;
                  st2s(M + i * PP + j, m0, m1, i);
                  M[j * PP + i] = m0;
                  M[(j + 1) * PP + i] = m1;
                } else {  // i == j + 1: the diagonal entry of an odd row
                  M[i * PP + i] += v1;
                }
              });
        }
      }
      cp_async_wait_all();
      w.cta_sync();
      CPHASE_MARK(5);
    }
  }

  // =========================================================================================
  // Forward pass with KL and expected cost
  // =========================================================================================
  if (opt_fwd) {
    const R reg = a.fwd_reg;
    const bool do_kl = FULL || (a.flags & ILQR_KL) != 0, do_cost = FULL || (a.flags & ILQR_COST) != 0;
    auto stage_policy = [&](int t) {  // [K_t | k_t] -> Kn, pol_covar_t -> Sinv (asynchronously)
      const size_t pt = pi * T + t;
      if (a.aligned16) {
        CTA_FOR(idx, DU * (DX / 2)) {
          const int a2 = idx / (DX / 2), j = 2 * (idx % (DX / 2));
          cp_async_pair(Kn + a2 * KLD + j, a.K + pt * szK + (size_t)a2 * DX + j);
        }
      } else {
        CTA_FOR(idx, DU * DX) cp_async(Kn + (idx / DX) * KLD + idx % DX, a.K + pt * szK + idx);
      }
      CTA_FOR(a2, DU) cp_async(Kn + a2 * KLD + DX, a.k + pt * DU + a2);
      CTA_FOR(idx, DU * DU) cp_async(Sinv + (idx / DU) * SLD + idx % DU, a.pcov + pt * szS + idx);
    };
    auto stage_prev = [&](int t) {  // [K_prev_t | k_prev_t] -> Kp, Lam_prev_t -> Lp, prev_hld_t (asynchronously)
      const size_t bt = (size_t)b * T + t;
      if (a.aligned16) {
        CTA_FOR(idx, DU * (DX / 2)) {
          const int a2 = idx / (DX / 2), j = 2 * (idx % (DX / 2));
          cp_async_pair(Kp + a2 * KLD + j, a.prevK + bt * szK + (size_t)a2 * DX + j);
        }
      } else {
        CTA_FOR(idx, DU * DX) cp_async(Kp + (idx / DX) * KLD + idx % DX, a.prevK + bt * szK + idx);
      }
      CTA_FOR(a2, DU) cp_async(Kp + a2 * KLD + DX, a.prevk + bt * DU + a2);
      CTA_FOR(idx, DU * DU) cp_async(Lp + (idx / DU) * SLD + idx % DU, a.prevLam + bt * szS + idx);
      CTA_FOR(one, 1) cp_async(Lp + DU * SLD, a.prev_hld + bt);
    };
    // cost weights C_t: every thread keeps its share (thread-linear over the P x P entries) in registers, loaded at the
    // start of the step and consumed in F3 - one full-latency trip to L2 per step instead of one per row
    constexpr int CQ = (P * P + 255) / 256;
    const bool c_early = do_cost && w.nthreads * CQ >= P * P;
    w.cta_sync();
    stage_F(0, 0);
    stage_policy(0);
    if (do_kl) stage_prev(0);
    CTA_FOR(idx, D::MSZ) M[idx] = R(0);
    w.cta_sync();
    CTA_FOR(idx, DX * DX) M[(idx / DX) * PP + idx % DX] = ldg(a.x0S + (size_t)b * DX * DX + idx);
    CTA_FOR(i, DX) {
      const R m0 = ldg(a.x0m + (size_t)b * DX + i);
      M[i * PP + P] = m0;
      mu[i] = m0;
    }
    CTA_FOR(i, 64) wp[i] = R(0);
    cp_async_wait_all();
    w.cta_sync();
    CPHASE_MARK(13);

    for (int t = 0; t < T; ++t) {
      const R* Fc = Fbuf(t & 1);
      const size_t bt = (size_t)b * T + t, bt1 = (size_t)b * (T + 1) + t;
      const R* Cg = a.C + bt1 * szC;
      const R* cg = a.c + bt1 * P;
      if (t + 1 < T) {
        stage_F(t + 1, (t + 1) & 1);
        if (do_cost) prefetch_lines(a.C + (bt1 + 1) * szC, P * P);
        prefetch_lines(a.dyn_covar + (bt + 1) * DX * DX, DX * DX);
        if (do_kl) prefetch_lines(a.prevK + (bt + 1) * szK, DU * DX);
      }
      R creg[CQ];
      if (c_early) {
#pragma unroll
        for (int q = 0; q < CQ; ++q) {
          const int idx = w.tid + q * w.nthreads;
          creg[q] = idx < P * P ? ldg(Cg + idx) : R(0);
        }
      }
      CPHASE_MARK(16);
      // ---- F1: [Sigma_ux | mu_u] = K [Sigma_xx | mu_x] (+ k) (lqg.py:227-230) --------------------------------------
      TEAMS_DO(tm) {
        for (int s = 0; s < sched.n[D::PH_F1][tm]; ++s) {
          const int ra0 = sched.t[D::PH_F1][tm][s][0], rb0 = sched.t[D::PH_F1][tm][s][1];
          tile_mma_n<R, D::F1_RA, D::F1_RB, 0, true, false, false>(
              w, NBX - rb0, false, Kn + 8 * ra0 * KLD, KLD, M + 8 * rb0, PP, DX / 4, [](int, int, R*) {},
              [&](int il, int jl, R v0, R v1, const R*) {
                const int a2 = 8 * ra0 + il, j = 8 * rb0 + jl;
                st2s(M + (DX + a2) * PP + j, v0, v1, a2);
                M[j * PP + DX + a2] = v0;
                M[(j + 1) * PP + DX + a2] = v1;
              });
        }
        // mu_u = K mu_x + k: one fragment per block row against the vector column, dealt to the warps from the back
        for (int ra = 0; ra < NBU; ++ra) {
          if ((G - 1 - ra % G) != tm) continue;
          tile_mma<R, 1, 1, 0, true, false, false>(
              w, Kn + 8 * ra * KLD, KLD, M + P, PP, DX / 4, [](int, int, R*) {},
              [&](int il, int jl, R v0, R, const R*) {
                if (jl == 0) {
                  const int a2 = 8 * ra + il;
                  const R mu_u = v0 + Kn[a2 * KLD + DX];
                  M[(DX + a2) * PP + P] = mu_u;
                  mu[DX + a2] = mu_u;
                }
              });
        }
      }
      CPHASE_MARK(17);
      w.cta_sync();
      CPHASE_MARK(6);
      // ---- F2: Sigma_uu = Sigma_ux K^T + pol_covar (upper, mirrored) (lqg.py:231); KL: delta_u, trace -----------------
      TEAMS_DO(tm) {
        // fragment (ra, rb), ra <= rb, of the action block: one per warp, round robin
        int q = 0;
        for (int ra = 0; ra < NBU; ++ra)
          for (int rb = ra; rb < NBU; ++rb, ++q) {
            if (q % G != tm) continue;
            tile_mma<R, 1, 1, 0, false, true, false>(
                w, M + DX + 8 * ra, PP, Kn + 8 * rb * KLD, KLD, DX / 4, [](int, int, R*) {},
                [&](int il, int jl, R v0, R v1, const R*) {
                  const int a2 = 8 * ra + il, bb = 8 * rb + jl;
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                    if (a2 <= bb + h) {
                      const R val = (h ? v1 : v0) + Sinv[a2 * SLD + bb + h];
                      M[(DX + a2) * PP + DX + bb + h] = val;
                      M[(DX + bb + h) * PP + DX + a2] = val;
                    }
                  }
                });
          }
        if (do_kl && tm == (NBU * (NBU + 1) / 2) % G) {
          // delta_u = (K_prev mu_x + k_prev) - mu_u  (lqg.py:294; mu_u = K mu_x + k comes from F1)
          tile_mma<R, NBU, 1, 0, true, false, false>(
              w, Kp, KLD, M + P, PP, DX / 4, [](int, int, R*) {},
              [&](int a2, int jl, R v0, R, const R*) {
                if (jl == 0) del[a2] = v0 + Kp[a2 * KLD + DX] - M[(DX + a2) * PP + P];
              });
        }
        if (do_kl && tm == G - 1) {  // tr(Lam_prev Sigma_new) / 2 + per-step constants (lqg.py:295-304)
          R part = R(0);
          LANE_FOR(idx, DU * DU) part = fma(Lp[(idx / DU) * SLD + idx % DU], Sinv[(idx % DU) * SLD + idx / DU], part);
          team_accumulate(w, wp + 32 + tm, part * R(0.5));
          LANE_FOR(one, 1) wp[32 + tm] += Lp[DU * SLD] - R(DU) * R(0.5);
        }
      }
      w.cta_sync();
      CPHASE_MARK(7);
      if (t + 1 < T) stage_policy(t + 1);  // Kn / Sinv are dead until the next step
      // ---- F3: Rt_ext = (F [Sigma | mu])^T (lqg.py:236-240); expected cost; KL quadratic term; trajectory outputs ----
      {
        R cost_part = R(0);
        if (do_cost) {  // 1/2 sum_kj C_kj (Sigma_kj + mu_k mu_j) + 1/2 reg tr C + c^T mu (quadratic_costs.py:106-119)
          if (c_early) {
#pragma unroll
            for (int q = 0; q < CQ; ++q) {
              const int idx = w.tid + q * w.nthreads;
              if (idx < P * P) {
                const int k = idx / P, j = idx % P;
                R sv = fma(mu[k], mu[j], M[k * PP + j]);
                if (k == j) sv += reg;
                cost_part = fma(R(0.5) * creg[q], sv, cost_part);
              }
            }
          } else {
            CTA_FOR(idx, P * P) {
              const int k = idx / P, j = idx % P;
              R sv = fma(mu[k], mu[j], M[k * PP + j]);
              if (k == j) sv += reg;
              cost_part = fma(R(0.5) * ldg(Cg + idx), sv, cost_part);
            }
          }
          CTA_FOR(j, P) cost_part = fma(ldg(cg + j), mu[j], cost_part);
        }
        CPHASE_MARK(18);
        TEAMS_DO(tm) {
          for (int s = 0; s < sched.n[D::PH_F3][tm]; ++s) {
            const int ra0 = sched.t[D::PH_F3][tm][s][0], rb0 = sched.t[D::PH_F3][tm][s][1];
            // rows i of F, columns j of [Sigma | mu]; stored transposed (Rt_ext is the k-major A operand of F4)
            tile_mma_n<R, D::F3_RA, D::F3_RB, 0, true, false, false>(
                w, NBP - rb0, false, Fc + 8 * ra0 * PP, PP, M + 8 * rb0, PP, P / 4, [](int, int, R*) {},
                [&](int il, int jl, R v0, R v1, const R*) {
                  const int i = 8 * ra0 + il, j = 8 * rb0 + jl;
                  if (j < PE) Rb[j * DXP + i] = v0;
                  if (j + 1 < PE) Rb[(j + 1) * DXP + i] = v1;
                });
          }
          if (do_kl && tm == 0) {  // delta^T Lam_prev delta / 2
            R part = R(0);
            LANE_FOR(idx, DU * DU) part = fma(del[idx / DU] * Lp[(idx / DU) * SLD + idx % DU], del[idx % DU], part);
            team_accumulate(w, wp + 32 + tm, part * R(0.5));
          }
        }
        CPHASE_MARK(19);
        if (do_cost) {
#if defined(DONK_EMU)
          wp[0] += cost_part;
#else
          team_accumulate(w, wp + w.team, cost_part);
#endif
        }
        if (want_tcov) {
          CTA_FOR(idx, P * P) {
            const int i = idx / P, j = idx % P;
            a.tcov[(pi * (T + 1) + t) * szC + idx] = M[i * PP + j] + (i == j ? reg : R(0));
          }
        }
        if (want_tmean) {
          CTA_FOR(i, P) a.tmean[(pi * (T + 1) + t) * P + i] = mu[i];
        }
      }
      w.cta_sync();
      CPHASE_MARK(8);
      if (do_kl && t + 1 < T) stage_prev(t + 1);  // Kp / Lp are dead until F2 of the next step
      // ---- F4: Sigma'_xx = (F Sigma) F^T + dyn_covar (upper blocks, mirrored); mu' = F mu + f --------------------------
      {
        const R* Dg = a.dyn_covar + bt * DX * DX;
        TEAMS_DO(tm) {
          for (int s = 0; s < sched.n[D::PH_F4][tm]; ++s) {
            const int ra0 = sched.t[D::PH_F4][tm][s][0], rb0 = sched.t[D::PH_F4][tm][s][1];
            tile_mma_n<R, D::F4_RA, D::F4_RB, 2, false, true, true>(
                w, NBX - rb0, rb0 == ra0, Rb + 8 * ra0, DXP, Fc + 8 * rb0 * PP, PP, P / 4,
                [&](int il, int jl, R* pv) {
                  const int i = 8 * ra0 + il, j = 8 * rb0 + jl;
                  if (i <= j + 1) ldg2(Dg + (size_t)i * DX + j, a.aligned16 != 0, pv[0], pv[1]);
                },
                [&](int il, int jl, R v0, R v1, const R* pv) {
                  const int i = 8 * ra0 + il, j = 8 * rb0 + jl;
                  if (i > j + 1) return;
                  const R s0 = v0 + pv[0], s1 = v1 + pv[1];
                  if (i <= j) {
                    st2s(M + i * PP + j, s0, s1, i);
                    M[j * PP + i] = s0;
                    M[(j + 1) * PP + i] = s1;
                  } else {
                    M[i * PP + i] = s1;
                  }
                });
          }
        }
        CTA_FOR(i, DX) {
          const R mu_n = Rb[P * DXP + i] + Fc[i * PP + P];
          M[i * PP + P] = mu_n;
          mu[i] = mu_n;
        }
        if (do_cost) {
          CTA_FOR(one, 1) wp[0] += ldg(a.cc + bt1);
        }
      }
      cp_async_wait_all();
      w.cta_sync();
      CPHASE_MARK(9);
    }

    // ---- final row T: state block only (lqg.py:221-222,242) ---------------------------------------------------
    {
      const size_t bT = (size_t)b * (T + 1) + T;
      if (do_cost) {
        R cost_part = R(0);
        WARP_ROWS_FOR(k, DX) {
          const R mk = mu[k];
          LANE_FOR(j, DX) cost_part = fma(R(0.5) * ldg(a.C + bT * szC + (size_t)k * P + j), fma(mk, mu[j], M[k * PP + j]), cost_part);
        }
        WARP_ROWS_FOR(one, 1) {
          LANE_FOR(j, DX) cost_part = fma(ldg(a.c + bT * P + j), mu[j], cost_part);
        }
#if defined(DONK_EMU)
        wp[0] += cost_part;
#else
        team_accumulate(w, wp + w.team, cost_part);
#endif
      }
      if (want_tcov) {
        CTA_FOR(idx, P * P) {
          const int i = idx / P, j = idx % P;
          a.tcov[(pi * (T + 1) + T) * szC + idx] = (i < DX && j < DX) ? M[i * PP + j] : R(0);
        }
      }
      if (want_tmean) {
        CTA_FOR(i, P) a.tmean[(pi * (T + 1) + T) * P + i] = i < DX ? mu[i] : R(0);
      }
      w.cta_sync();
    }
  }

  // ---- results ----------------------------------------------------------------------------------------------------
  CTA_FOR(one, 1) {
    if (FULL || ((a.flags & ILQR_FORWARD) && (a.flags & ILQR_COST) && a.cost)) {
      R acc = R(0);
      for (int i = 0; i < 32; ++i) acc += wp[i];
      a.cost[pi] = acc + ldg(a.cc + (size_t)b * (T + 1) + T);
    }
    if (FULL || ((a.flags & ILQR_FORWARD) && (a.flags & ILQR_KL) && a.kl)) {
      R acc = R(0);
      for (int i = 0; i < 32; ++i) acc += wp[32 + i];
      a.kl[pi] = (acc + scal[3]) / R(T);
    }
    if (a.info) a.info[pi] = (int)scal[4];
  }
}

}  // namespace donk
