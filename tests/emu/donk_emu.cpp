// Host emulation of the donk_b200 kernel bodies — TEST INFRASTRUCTURE ONLY.
//
// Compiles donk_b200/csrc/*.cuh with -DDONK_EMU (see common.cuh): one host thread plays a
// whole CTA, phases run sequentially (forward or reversed order, donk_emu_set_reverse) and
// shared memory is a heap buffer.  It exists so kernel logic can be checked against the
// oracle in the build container, which has no GPU.  The product (donk_b200/_lib.py) never
// loads this library; entry points carry the donk_emu_ prefix and take HOST pointers.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../donk_b200/csrc/fitlr.cuh"
#include "../../donk_b200/csrc/fitlr_warp.cuh"
#include "../../donk_b200/csrc/fitlr_coop.cuh"
#include "../../donk_b200/csrc/gmm.cuh"
#include "../../donk_b200/csrc/gmm_mma.cuh"
#include "../../donk_b200/csrc/ilqr.cuh"
#include "../../donk_b200/csrc/kmeans.cuh"
#include "../../donk_b200/csrc/ilqr_warp.cuh"
#include "../../donk_b200/csrc/optimize.cuh"
#include "../../donk_b200/csrc/costs.cuh"
#include "../../donk_b200/csrc/ilqr_coop.cuh"

namespace donk {
int g_emu_reverse = 0;
}
using namespace donk;

static const size_t kSmemCap = 227 * 1024;

static int env_int(const char* name, int unset = 0) {
  const char* v = getenv(name);
  return v ? atoi(v) : unset;
}

template <typename R, int DX, int DU>
static int emu_ilqr_warp(const IlqrArgs<R>& a) {
  int G, ES, grid;
  size_t bytes;
  int rc = ilqr_warp_plan<R, DX, DU>(a.B, a.E, kSmemCap, env_int("DONK_ILQR_WARPS"), G, ES, grid, bytes);
  if (rc != DONK_OK) return rc;
  std::vector<R> smem(bytes / sizeof(R) + 4);
  WCtx w{0, 1, 0, 0, G};
  for (int cta = 0; cta < grid; ++cta) {
    memset(smem.data(), 0xFF, smem.size() * sizeof(R));
    if (ilqr_warp_is_full<R, DX, DU>(a)) ilqr_warp_cta<R, DX, DU, true>(w, a, cta, ES, smem.data());
    else ilqr_warp_cta<R, DX, DU, false>(w, a, cta, ES, smem.data());
  }
  return DONK_OK;
}

template <int DX, int DU>
static int emu_ilqr_coop(const IlqrArgs<double>& a) {
  int G = env_int("DONK_COOP_WARPS");
  if (G <= 0) G = 8;
  if (G > 16) G = 16;
  const size_t bytes = ilqr_coop_bytes<double, DX, DU>();
  if (bytes > kSmemCap || !coop_schedule_fits<DX, DU>(G)) return DONK_SMEM_LIMIT;
  std::vector<double> smem(bytes / sizeof(double) + 4);
  WCtx w{0, 1, 0, 0, G};
  for (int cta = 0; cta < a.B * a.E; ++cta) {
    memset(smem.data(), 0xFF, smem.size() * sizeof(double));
    if (ilqr_coop_is_full<double, DX, DU>(a)) ilqr_coop_cta<double, DX, DU, true>(w, a, cta, smem.data());
    else ilqr_coop_cta<double, DX, DU, false>(w, a, cta, smem.data());
  }
  return DONK_OK;
}

template <typename R>
static int emu_ilqr_step(IlqrArgs<R> a) {
  if (a.B < 0 || a.E < 1 || a.T < 1 || a.dX < 1 || a.dU < 1) return DONK_BAD_SHAPE;
  if ((a.flags & ILQR_KL) && (a.flags & (ILQR_BACKWARD | ILQR_FORWARD)) != (ILQR_BACKWARD | ILQR_FORWARD))
    return DONK_BAD_SHAPE;
  if (a.B == 0) return DONK_OK;
  if (!env_int("DONK_ILQR_GENERIC")) {
    if (a.dX == 20 && a.dU == 6) return emu_ilqr_warp<R, 20, 6>(a);
    if (a.dX == 14 && a.dU == 7) return emu_ilqr_warp<R, 14, 7>(a);
    if (a.dX == 6 && a.dU == 2) return emu_ilqr_warp<R, 6, 2>(a);
    if constexpr (sizeof(R) == 8) {
      int rc_c = -1;
      if (a.dX == 64 && a.dU == 16) rc_c = emu_ilqr_coop<64, 16>(a);
      else if (a.dX == 16 && a.dU == 8) rc_c = emu_ilqr_coop<16, 8>(a);
      if (rc_c >= 0 && rc_c != DONK_SMEM_LIMIT) return rc_c;
    }
  }
  IlqrPlan pl;
  int rc;
  rc = ilqr_plan(a.B, a.E, a.dX, a.dU, sizeof(R), kSmemCap, env_int("DONK_ILQR_SLOTS"), 0, pl);
  if (rc != DONK_OK) return rc;
  a.S = pl.S;
  a.ES = pl.ES;
  std::vector<R> smem(pl.smem / sizeof(R) + 4);
  Ctx ctx{0, 1};
  for (int cta = 0; cta < pl.grid; ++cta) {
    memset(smem.data(), 0xFF, smem.size() * sizeof(R));  // poison: kernels must not depend on initial contents
    ilqr_cta<R>(ctx, a, cta, smem.data());
  }
  return DONK_OK;
}

static int emu_fit_prior_wts(FitArgs<double>& a, std::vector<double>& wts);

template <int DX, int DU>
static int emu_fit_warp(FitArgs<double> a) {
  int G;
  size_t bytes;
  bool pre_wts = false;  // as the CUDA launcher (abi_fit.cu: fit_warp_launch)
  if (a.K > 0) {
    const int force_in = env_int("DONK_FIT_PRIOR_INKERNEL", -1);
    pre_wts = a.N > 64 || (force_in < 0 ? (a.K >= 8 && (long long)a.B * a.N * a.T >= 32768) : force_in == 0);
  }
  int rc = fit_warp_plan<double, DX, DU>(kSmemCap, 0, a.N, a.K, G, bytes, pre_wts);
  if (rc != DONK_OK) return rc;
  std::vector<double> consts, wts;
  if (pre_wts) {
    rc = emu_fit_prior_wts(a, wts);
    if (rc != DONK_OK) return rc;
  }
  if (a.K > 0 && !pre_wts) {
    const int d = 2 * DX + DU, len = a.K * d + a.K;
    consts.resize(len);
    for (int i = 0; i < len; ++i) gmm_prior_consts_item<double>(d, a.K, a.gw, a.gmeans, a.gpc, consts.data(), i);
    a.gconst = consts.data();
  }
  std::vector<double> smem(bytes / sizeof(double) + 2);
  WCtx w{0, 1, 0, 0, G};
  const long long nfits = (long long)a.B * a.T;
  for (long long cta = 0; cta < (nfits + G - 1) / G; ++cta) {
    memset(smem.data(), 0xFF, smem.size() * sizeof(double));
    fit_warp_cta<double, DX, DU>(w, a, (int)cta, smem.data());
  }
  return DONK_OK;
}

extern "C" {

void donk_emu_set_reverse(int r) { g_emu_reverse = r; }

#define EMU_ILQR_STEP(SUFFIX, R)                                                                                  \
  int donk_emu_ilqr_step_##SUFFIX(int B, int E, int T, int dX, int dU, int flags, const R* F, const R* f,        \
                                  const R* dyn_covar, const R* C, const R* c, const R* cc, const R* C_kl,        \
                                  const R* c_kl, const R* prevK, const R* prevk, const R* prevLam,               \
                                  const R* prev_hld, const R* x0_mean, const R* x0_covar, const R* eta, R gamma, \
                                  R fwd_reg, const int* active, R* K, R* k, R* pol_covar, R* pol_inv_covar,      \
                                  R* kl, R* cost, R* traj_mean, R* traj_covar, int* info, void*) {               \
    IlqrArgs<R> a;                                                                                               \
    memset(&a, 0, sizeof(a));                                                                                    \
    a.B = B; a.E = E; a.T = T; a.dX = dX; a.dU = dU; a.flags = flags;                                            \
    a.F = F; a.f = f; a.dyn_covar = dyn_covar; a.C = C; a.c = c; a.cc = cc; a.C_kl = C_kl; a.c_kl = c_kl;        \
    a.prevK = prevK; a.prevk = prevk; a.prevLam = prevLam; a.prev_hld = prev_hld;                                \
    a.x0m = x0_mean; a.x0S = x0_covar; a.eta = eta; a.gamma = gamma; a.fwd_reg = fwd_reg;                        \
    a.K = K; a.k = k; a.pcov = pol_covar; a.pinv = pol_inv_covar; a.kl = kl; a.cost = cost;                      \
    a.tmean = traj_mean; a.tcov = traj_covar; a.info = info; a.active = active; a.aligned16 = 1;                                  \
    return emu_ilqr_step<R>(a);                                                                                  \
  }
EMU_ILQR_STEP(f64, double)
EMU_ILQR_STEP(f32, float)

int donk_emu_extended_costs_kl_f64(int n, int dX, int dU, const double* K, const double* k, const double* Lam,
                                   double* C_kl, double* c_kl, void*) {
  const int npairs = 4;
  std::vector<double> smem(extended_costs_smem(dX, dU) * npairs + 2);
  Ctx ctx{0, 1};
  for (int cta = 0; cta < (n + npairs - 1) / npairs; ++cta) {
    memset(smem.data(), 0xFF, smem.size() * sizeof(double));
    extended_costs_cta<double>(ctx, n, npairs, cta, dX, dU, K, k, Lam, C_kl, c_kl, smem.data());
  }
  return DONK_OK;
}

int donk_emu_spd_factor_f64(int n, int d, const double* A, double* Ainv, double* chol, double* half_logdet, int* info,
                            void*) {
  std::vector<double> ws(2 * d * d);
  for (int i = 0; i < n; ++i) {
    const size_t o = (size_t)i * d * d;
    int bad = spd_inverse_item<double>(d, A + o, Ainv ? Ainv + o : nullptr, chol ? chol + o : nullptr,
                                       half_logdet ? half_logdet + i : nullptr, ws.data());
    if (info) info[i] = bad;
  }
  return DONK_OK;
}

int donk_emu_kl_divergence_action_f64(int B, int T, int dX, int dU, const double* X, const double* K, const double* k,
                                      const double* covar, const double* hld, const double* pK, const double* pk,
                                      const double* pLam, const double* phld, double* kl, void*) {
  std::vector<double> delta(dU);
  for (int b = 0; b < B; ++b) {
    double acc = 0;
    for (int t = 0; t < T; ++t) {
      const size_t bt = (size_t)b * T + t;
      acc += kl_term_item<double>(dX, dU, X + bt * dX, K + bt * dU * dX, k + bt * dU, covar + bt * dU * dU, hld[bt],
                                  pK + bt * dU * dX, pk + bt * dU, pLam + bt * dU * dU, phld[bt], delta.data());
    }
    kl[b] = acc / T;
  }
  return DONK_OK;
}

int donk_emu_quadratic_cost_l2_f64(int n, int p, const double* t, const double* w, double* C, double* c, double* cc, void*) {
  if (n < 0 || p < 1) return DONK_BAD_SHAPE;
  const size_t total = (size_t)n * p * p;
  for (size_t it = 0; it < total; ++it) cost_l2_item<double>(p, t, w, C, c, cc, g_emu_reverse ? total - 1 - it : it);
  return DONK_OK;
}
int donk_emu_quadratic_cost_l1_f64(int n, int p, const double* xu, const double* t, const double* w, double alpha,
                                   double* C, double* c, double* cc, void*) {
  if (n < 0 || p < 1) return DONK_BAD_SHAPE;
  const size_t total = (size_t)n * p * p;
  for (size_t it = 0; it < total; ++it) cost_l1_item<double>(p, xu, t, w, alpha, C, c, cc, g_emu_reverse ? total - 1 - it : it);
  return DONK_OK;
}
int donk_emu_dynamics_log_prob_f64(int B, int N, int T, int dX, int dU, const double* X, const double* U, const double* F,
                                   const double* f, const double* covar_inv, const double* half_logdet, double* log_prob,
                                   void*) {
  if (B < 0 || N < 0 || T < 1 || dX < 1 || dU < 1) return DONK_BAD_SHAPE;
  std::vector<double> diff(dX);
  const size_t total = (size_t)B * N * T;
  for (size_t idx = 0; idx < total; ++idx)
    log_prob[idx] = dyn_log_prob_item<double>(N, T, dX, dU, X, U, F, f, covar_inv, half_logdet, idx, diff.data());
  return DONK_OK;
}

size_t donk_emu_ilqr_optimize_workspace_bytes(int B) { return optimize_workspace_bytes(B); }

// the graph of abi_optimize.cu as a host loop: the same rounds, the same controller items
int donk_emu_ilqr_optimize_f64(int B, int T, int dX, int dU, const double* F, const double* f, const double* dyn_covar,
                               const double* C, const double* c, const double* cc, const double* C_kl,
                               const double* c_kl, const double* prevK, const double* prevk, const double* prevLam,
                               const double* prev_hld, const double* x0_mean, const double* x0_covar,
                               const double* kl_step, double min_eta, double max_eta, double eta_a, double eta_b,
                               double xa, double xb, double rtol, int maxiter, double* eta, int* active, double* K,
                               double* k, double* pol_covar, double* pol_inv_covar, double* kl, double* cost,
                               double* traj_mean, double* traj_covar, int* info, int* status, int* counters, void* work,
                               size_t work_bytes, void*) {
  if (B < 0 || T < 1 || dX < 1 || dU < 1 || maxiter < 1 || work_bytes < optimize_workspace_bytes(B)) return DONK_BAD_SHAPE;
  if (B == 0) return DONK_OK;
  IlqrArgs<double> s;
  memset(&s, 0, sizeof(s));
  s.B = B; s.E = 1; s.T = T; s.dX = dX; s.dU = dU;
  s.flags = ILQR_BACKWARD | ILQR_FORWARD | ILQR_KL | ILQR_COST;
  s.F = F; s.f = f; s.dyn_covar = dyn_covar; s.C = C; s.c = c; s.cc = cc; s.C_kl = C_kl; s.c_kl = c_kl;
  s.prevK = prevK; s.prevk = prevk; s.prevLam = prevLam; s.prev_hld = prev_hld;
  s.x0m = x0_mean; s.x0S = x0_covar; s.eta = eta; s.gamma = 1.0; s.fwd_reg = 1e-6;
  s.K = K; s.k = k; s.pcov = pol_covar; s.pinv = pol_inv_covar; s.kl = kl; s.cost = cost;
  s.info = info; s.active = active; s.aligned16 = 1;
  IlqrArgs<double> fin = s;
  fin.tmean = traj_mean; fin.tcov = traj_covar;
  uintptr_t wp = ((uintptr_t)work + 15) & ~(uintptr_t)15;
  OptArgs<double> o;
  memset(&o, 0, sizeof(o));
  o.B = B; o.kl_step = kl_step; o.min_eta = min_eta; o.max_eta = max_eta; o.eta_a = eta_a; o.eta_b = eta_b;
  o.xa = xa; o.xb = xb; o.xtol = 2e-12; o.rtol = rtol; o.maxiter = maxiter;
  o.kl = kl; o.info = info; o.eta = eta; o.active = active; o.status = status;
  o.st = reinterpret_cast<OptState<double>*>(wp);
  o.ctl = reinterpret_cast<OptCtl*>(wp + (size_t)B * sizeof(OptState<double>));
  memset(o.ctl, 0, sizeof(OptCtl));
  auto round = [&](int phase) {
    int rc = DONK_OK;
    if (phase != OPT_FINAL && phase != OPT_INIT) rc = emu_ilqr_step<double>(s);
    if (rc != DONK_OK) return rc;
    int n = 0;
    for (int it = 0; it < B; ++it) n += optimize_ctl_item<double>(o, phase, emu_idx(it, B), &o.ctl->err);
    o.ctl->n_active = n;
    if (phase == OPT_MIN) o.ctl->n_after_min = n;
    if (phase == OPT_LOOP) o.ctl->loop_iters += 1;
    return DONK_OK;
  };
  int rc = DONK_OK;
  for (int ph = OPT_INIT; ph <= OPT_FB && rc == DONK_OK; ++ph) rc = round(ph);
  while (rc == DONK_OK && o.ctl->n_active > 0 && o.ctl->loop_iters <= maxiter + 2) rc = round(OPT_LOOP);
  if (rc == DONK_OK) rc = round(OPT_FINAL);
  if (rc == DONK_OK) rc = emu_ilqr_step<double>(fin);
  counters[0] = o.ctl->loop_iters; counters[1] = o.ctl->err; counters[2] = o.ctl->n_after_min;
  return rc;
}

int donk_emu_brent_update_f64(int B, int phase, double xa, double xb, const double* fa, const double* fb,
                              const double* fnew, double kl_step, double xtol, double rtol, int maxiter,
                              double* state, double* eta, int* active, int* status, int* n_active, void*) {
  *n_active = 0;
  for (int b = 0; b < B; ++b) {
    BrentState<double>* st = reinterpret_cast<BrentState<double>*>(state) + b;
    if (active[b] == 0) continue;
    int act = 0;
    double e = eta[b];
    brent_item<double>(*st, phase, xa, xb, phase == 0 ? fa[b] - kl_step : 0.0, phase == 0 ? fb[b] - kl_step : 0.0,
                       phase == 0 ? 0.0 : fnew[b] - kl_step, xtol, rtol, maxiter, &e, &act);
    eta[b] = e;
    active[b] = act;
    status[b] = (int)st->status;
    if (act) *n_active += 1;
  }
  return DONK_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// fit_lr / transitions / state fit / GMM EM under emulation
// ---------------------------------------------------------------------------------------------
// forward declaration: responsibilities through the emulated E pass (defined with the GMM entry points below)
extern "C" int donk_emu_gmm_predict_proba_f64(int n, int d, int K, const double* X, const double* weights, const double* means,
                                              const double* prec_chol, double* resp, void*);

// as the CUDA launcher (abi_fit_coop.cu: fit_prior_wts_launch): E pass over the transition view of the roll-outs, then
// the average over the samples of every fit
extern "C" {
static void emu_gmm_e(GmmArgs<double>& a);
}
static int emu_fit_prior_wts(FitArgs<double>& a, std::vector<double>& wts) {
  const int d = 2 * a.dX + a.dU;
  const size_t nrows = (size_t)a.B * a.N * a.T;
  std::vector<double> resp(nrows * a.K);
  wts.resize((size_t)a.B * a.T * a.K);
  GmmArgs<double> g;
  memset(&g, 0, sizeof(g));
  g.n = (int)nrows; g.d = d; g.K = a.K; g.w = a.gw; g.means = a.gmeans; g.pc = a.gpc;
  g.tX = a.X; g.tU = a.U; g.tT = a.T; g.tdX = a.dX; g.tdU = a.dU;
  g.nblocks = (g.n + GMM_TS - 1) / GMM_TS;
  g.resp = resp.data();
  emu_gmm_e(g);
  for (size_t idx = 0; idx < wts.size(); ++idx) {
    const int k = (int)(idx % a.K);
    const size_t bt = idx / a.K, t = bt % a.T, b = bt / a.T;
    double acc = 0.0;
    for (int n = 0; n < a.N; ++n) acc += resp[((b * a.N + n) * a.T + t) * a.K + k];
    wts[idx] = acc / a.N;
  }
  a.wts_in = wts.data();
  return DONK_OK;
}

template <int DX, int DU>
static int emu_fit_coop(FitArgs<double>& a) {
  const size_t bytes = fit_coop_bytes<double, DX, DU>();
  if (bytes > kSmemCap) return DONK_SMEM_LIMIT;
  std::vector<double> wts;
  if (a.K > 0) {
    int rc = emu_fit_prior_wts(a, wts);
    if (rc != DONK_OK) return rc;
  }
  std::vector<double> smem(bytes / sizeof(double) + 4);
  int G = env_int("DONK_COOP_WARPS");
  if (G <= 0) G = 8;
  if (env_int("DONK_EMU_TRACE")) fprintf(stderr, "[emu] fit_coop_cta<%d,%d> K=%d N=%d G=%d\n", DX, DU, a.K, a.N, G);
  WCtx w{0, 1, 0, 0, G};
  const int ncta = 3;  // a few persistent "CTAs", each striding over the fits
  for (int cta = 0; cta < ncta; ++cta) {
    memset(smem.data(), 0xFF, smem.size() * sizeof(double));
    fit_coop_cta<double, DX, DU>(w, a, cta, ncta, smem.data());
  }
  a.wts_in = nullptr;
  return DONK_OK;
}

extern "C" {

static int emu_fit_lr_launch(FitArgs<double>& a) {
  const int B = a.B, N = a.N, T = a.T, dX = a.dX, dU = a.dU, gmm_K = a.K;
  if (!a.niw_Phi && !env_int("DONK_FIT_GENERIC")) {
    int rc = DONK_SMEM_LIMIT;
    if (dX == 20 && dU == 6) rc = emu_fit_warp<20, 6>(a);
    else if (dX == 14 && dU == 7) rc = emu_fit_warp<14, 7>(a);
    else if (dX == 6 && dU == 2) rc = emu_fit_warp<6, 2>(a);
    if (rc != DONK_SMEM_LIMIT) return rc;  // as the CUDA launcher: a plan that does not fit falls through
  }
  if (!env_int("DONK_FIT_GENERIC") && a.info != nullptr) {  // cooperative kernel + generic redo of the flagged fits
    int rc_c = -1;
    if (dX == 64 && dU == 16) rc_c = emu_fit_coop<64, 16>(a);
    else if (dX == 16 && dU == 8) rc_c = emu_fit_coop<16, 8>(a);
    if (rc_c == DONK_OK) a.redo = 1;
    else if (rc_c >= 0 && rc_c != DONK_SMEM_LIMIT) return rc_c;
  }
  size_t bytes = 0;
  const int force = env_int("DONK_FIT_CHUNK");
  a.NS = fit_plan(dX, dU, gmm_K, force > 0 ? force : N, sizeof(double), kSmemCap, bytes);
  if (a.NS == 0) return DONK_SMEM_LIMIT;
  FitLayout L(dX, dU, gmm_K, N, a.NS);
  std::vector<double> smem(L.total + 2);
  Ctx ctx{0, 1};
  if (env_int("DONK_EMU_TRACE")) {
    int redo = 0;
    for (int i = 0; i < B * T; ++i) redo += a.redo && a.info[i] == 2;
    fprintf(stderr, "[emu] fit_lr_cta generic: %d of %d fits\n", a.redo ? redo : B * T, B * T);
  }
  for (int cta = 0; cta < B * T; ++cta) {
    memset(smem.data(), 0xFF, smem.size() * sizeof(double));
    fit_lr_cta<double>(ctx, a, cta, smem.data());
  }
  return DONK_OK;
}

int donk_emu_fit_lr_f64(int B, int N, int T, int dX, int dU, const double* X, const double* U, int gmm_K,
                        const double* gmm_weights, const double* gmm_means, const double* gmm_covariances,
                        const double* gmm_prec_chol, double n_pool, double prior_weight, double regularization,
                        double* F, double* f, double* covar, int* info, void*) {
  if (B < 0 || N <= 1 || T < 1 || dX < 1 || dU < 1 || gmm_K < 0) return DONK_BAD_SHAPE;
  FitArgs<double> a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.N = N; a.T = T; a.dX = dX; a.dU = dU; a.X = X; a.U = U; a.K = gmm_K;
  a.gw = gmm_weights; a.gmeans = gmm_means; a.gcov = gmm_covariances; a.gpc = gmm_prec_chol;
  a.n_pool = n_pool; a.prior_weight = prior_weight; a.reg = regularization;
  a.F = F; a.f = f; a.covar = covar; a.info = info; a.aligned16 = 1;
  return emu_fit_lr_launch(a);
}

// single precision: the generic CTA kernel instantiated for float (the CUDA launcher also tries the warp kernels)
int donk_emu_fit_lr_f32(int B, int N, int T, int dX, int dU, const float* X, const float* U, int gmm_K,
                        const float* gmm_weights, const float* gmm_means, const float* gmm_covariances,
                        const float* gmm_prec_chol, float n_pool, float prior_weight, float regularization,
                        float* F, float* f, float* covar, int* info, void*) {
  if (B < 0 || N <= 1 || T < 1 || dX < 1 || dU < 1 || gmm_K < 0) return DONK_BAD_SHAPE;
  FitArgs<float> a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.N = N; a.T = T; a.dX = dX; a.dU = dU; a.X = X; a.U = U; a.K = gmm_K;
  a.gw = gmm_weights; a.gmeans = gmm_means; a.gcov = gmm_covariances; a.gpc = gmm_prec_chol;
  a.n_pool = n_pool; a.prior_weight = prior_weight; a.reg = regularization;
  a.F = F; a.f = f; a.covar = covar; a.info = info; a.aligned16 = 1;
  size_t bytes = 0;
  a.NS = fit_plan(dX, dU, gmm_K, N, sizeof(float), kSmemCap, bytes);
  if (a.NS == 0) return DONK_SMEM_LIMIT;
  FitLayout L(dX, dU, gmm_K, N, a.NS);
  std::vector<float> smem(L.total + 2);
  Ctx ctx{0, 1};
  for (int cta = 0; cta < B * T; ++cta) {
    memset(smem.data(), 0xFF, smem.size() * sizeof(float));
    fit_lr_cta<float>(ctx, a, cta, smem.data());
  }
  return DONK_OK;
}

int donk_emu_fit_lr_niw_f64(int B, int N, int T, int dX, int dU, const double* X, const double* U, const double* mu0,
                            const double* Phi, const double* Nm, const double* Nc, double prior_weight,
                            double regularization, double* F, double* f, double* covar, int* info, void*) {
  if (B < 0 || N <= 1 || T < 1 || dX < 1 || dU < 1 || !mu0 || !Phi || !Nm || !Nc) return DONK_BAD_SHAPE;
  FitArgs<double> a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.N = N; a.T = T; a.dX = dX; a.dU = dU; a.X = X; a.U = U; a.K = 0;
  a.niw_mu0 = mu0; a.niw_Phi = Phi; a.niw_Nmean = Nm; a.niw_Ncovar = Nc;
  a.prior_weight = prior_weight; a.reg = regularization;
  a.F = F; a.f = f; a.covar = covar; a.info = info; a.aligned16 = 1;
  return emu_fit_lr_launch(a);
}

int donk_emu_to_transitions_f64(int N, int T, int dX, int dU, const double* X, const double* U, double* XUX, void*) {
  const size_t total = (size_t)N * T * (2 * dX + dU);
  for (size_t i = 0; i < total; ++i) XUX[i] = transitions_item<double>(T, dX, dU, X, U, i);
  return DONK_OK;
}

int donk_emu_state_fit_f64(int B, int N, int dX, const double* X0, size_t sample_stride, size_t cond_stride,
                           double* mean, double* covar, void*) {
  if (B < 0 || N < 2 || dX < 1) return DONK_BAD_SHAPE;
  for (size_t i = 0; i < (size_t)B * dX * dX; ++i) state_fit_item<double>(N, dX, X0, sample_stride, cond_stride, i, mean, covar);
  return DONK_OK;
}

size_t donk_emu_gmm_stats_size(int d, int K) { return gmm_stats_len(d, K); }
size_t donk_emu_gmm_workspace_bytes(int n_local, int d, int K) { return gmm_ws_reals(n_local, d, K) * sizeof(double); }

static void emu_gmm_e(GmmArgs<double>& a) {
  std::vector<double> smem(gmm_e_smem(a.d, a.K) + 2);
  Ctx ctx{0, 1};
  const int ncta = a.nblocks < 7 ? a.nblocks : 7;  // a few "CTAs", each striding over blocks
  for (int cta = 0; cta < ncta; ++cta) {
    memset(smem.data(), 0xFF, smem.size() * sizeof(double));
    gmm_e_cta<double>(ctx, a, cta, ncta, smem.data());
  }
}

}  // extern "C"

template <int NB, int RA>
static void emu_gmm_e_mma(GmmArgs<double>& a, const GmmMmaPlan& pl) {
  const int TS = 8 * RA * pl.Ge;
  a.nblocks = (a.n + TS - 1) / TS;
  std::vector<double> smem(gmm_e_mma_smem<NB, RA>(a.K, pl.Ge) + 2);
  WCtx w{0, 1, 0, 0, pl.Ge};
  const int ncta = a.nblocks < 5 ? a.nblocks : 5;
  for (int cta = 0; cta < ncta; ++cta) {
    memset(smem.data(), 0xFF, smem.size() * sizeof(double));
    gmm_e_mma_cta<double, NB, RA>(w, a, cta, ncta, smem.data());
  }
}

extern "C" {

// ---- stand-in for the tensor-core E pass (abi_gmm_tc.cu is tcgen05 code with no host form): the same interface and the
// same arithmetic in plain float (operands rounded to fp32, fp32 accumulation), so that the host logic around it can be
// exercised on CPU.  The planes buffer holds (x - c) as float in its first plane; the second stays zero.
int donk_emu_gmm_tc_supported(int d, int K) { return d >= 1 && d <= 80 && K >= 1 ? 1 : 0; }
size_t donk_emu_gmm_tc_planes_bytes(int n, int d) {
  const size_t one = ((size_t)n * round_up(d, 8) * sizeof(float) + 255) & ~(size_t)255;
  return 2 * one;
}
size_t donk_emu_gmm_tc_workspace_bytes(int n, int d, int K) { return 256 + (size_t)n * K * sizeof(float) + (size_t)K * d * 8; }
int donk_emu_gmm_tc_prepare_f32(int n, int d, const double* X, const double* center, void* planes, size_t planes_bytes, void*) {
  if (n < 0 || d < 1 || planes_bytes < donk_emu_gmm_tc_planes_bytes(n, d)) return DONK_BAD_SHAPE;
  const int dp = round_up(d, 8);
  float* hi = static_cast<float*>(planes);
  memset(planes, 0, donk_emu_gmm_tc_planes_bytes(n, d));
  for (size_t r = 0; r < (size_t)n; ++r)
    for (int j = 0; j < d; ++j) hi[r * dp + j] = (float)(X[r * d + j] - (center ? center[j] : 0.0));
  return DONK_OK;
}
int donk_emu_gmm_tc_estep_f32(int n, int d, int K, const void* planes, const double* center, const double* weights,
                              const double* means, const double* prec_chol, double* resp, double* lse_part, void* work,
                              size_t work_bytes, void*) {
  if (n < 0 || d < 1 || K < 1 || work_bytes < donk_emu_gmm_tc_workspace_bytes(n, d, K)) return DONK_BAD_SHAPE;
  const int dp = round_up(d, 8);
  const float* hi = static_cast<const float*>(planes);
  const size_t dd = (size_t)d * d;
  std::vector<float> bias((size_t)K * d), cst(K), lp(K);
  for (int k = 0; k < K; ++k) {
    double ld = 0.0;
    for (int j = 0; j < d; ++j) {
      double acc = 0.0;
      for (int q = 0; q <= j; ++q) acc = fma(means[(size_t)k * d + q] - (center ? center[q] : 0.0), prec_chol[k * dd + (size_t)q * d + j], acc);
      bias[(size_t)k * d + j] = (float)acc;
      ld += log(prec_chol[k * dd + (size_t)j * d + j]);
    }
    cst[k] = (float)(ld + log(weights[k]) - 0.5 * d * 1.8378770664093454835606594728112);
  }
  const size_t nblocks = ((size_t)n + GMM_TS - 1) / GMM_TS;
  for (size_t b = 0; b < nblocks; ++b) lse_part[b] = 0.0;
  for (size_t r = 0; r < (size_t)n; ++r) {
    for (int k = 0; k < K; ++k) {
      float ss = 0.f;
      for (int j = 0; j < d; ++j) {
        float acc = 0.f;
        for (int q = 0; q <= j; ++q) acc = fmaf(hi[r * dp + q], (float)prec_chol[k * dd + (size_t)q * d + j], acc);
        const float y = acc - bias[(size_t)k * d + j];
        ss = fmaf(y, y, ss);
      }
      lp[k] = cst[k] - 0.5f * ss;
    }
    double m = lp[0];
    for (int k = 1; k < K; ++k) m = lp[k] > m ? (double)lp[k] : m;
    double s = 0.0;
    for (int k = 0; k < K; ++k) s += exp((double)lp[k] - m);
    for (int k = 0; k < K; ++k) resp[r * K + k] = exp((double)lp[k] - m) / s;
    lse_part[r / GMM_TS] += m + log(s);
  }
  (void)work;
  return DONK_OK;
}

// ---- stand-in for the tensor-core M pass: same interface; float operands, float accumulation over runs of 4096 rows
// that are then added up in double (the kernel drains its fp32 accumulators every 64 tiles of 64 rows).  The mplanes
// buffer holds (x - c) as float, row-major (n, d).
int donk_emu_gmm_tc_m_supported(int d, int K) { return d >= 1 && d <= 80 && K >= 1 ? 1 : 0; }
size_t donk_emu_gmm_tc_mplanes_bytes(int n, int d) {
  const size_t one = ((size_t)round_up(d + 1, 16) * round_up(n > 0 ? n : 1, 64) * sizeof(float) + 255) & ~(size_t)255;
  return 2 * one;
}
size_t donk_emu_gmm_tc_mwork_bytes(int n, int d, int K) { return 256 + (size_t)round_up(n > 0 ? n : 1, 64) * K * 4; }
int donk_emu_gmm_tc_mprepare_f32(int n, int d, const double* X, const double* center, void* mplanes, size_t bytes, void*) {
  if (n < 0 || d < 1 || bytes < donk_emu_gmm_tc_mplanes_bytes(n, d)) return DONK_BAD_SHAPE;
  float* x = static_cast<float*>(mplanes);
  for (size_t r = 0; r < (size_t)n; ++r)
    for (int j = 0; j < d; ++j) x[r * d + j] = (float)(X[r * d + j] - (center ? center[j] : 0.0));
  return DONK_OK;
}
// stand-in for the E pass that hands over to the M pass: the responsibilities are kept for it by address (the emulation's
// M pass reads fp64 responsibilities; the CUDA library keeps transposed fp32 copies in `mwork` instead)
int donk_emu_gmm_tc_estep_m_f32(int n, int d, int K, const void* planes, const double* center, const double* weights,
                                const double* means, const double* prec_chol, double* resp, double* lse_part, void* work,
                                size_t work_bytes, void* mwork, size_t mwork_bytes, void*) {
  if (!mwork || !resp || mwork_bytes < donk_emu_gmm_tc_mwork_bytes(n, d, K)) return DONK_BAD_SHAPE;
  int rc = donk_emu_gmm_tc_estep_f32(n, d, K, planes, center, weights, means, prec_chol, resp, lse_part, work, work_bytes, nullptr);
  if (rc != DONK_OK) return rc;
  *reinterpret_cast<const double**>((reinterpret_cast<uintptr_t>(mwork) + 15) & ~(uintptr_t)15) = resp;
  return DONK_OK;
}

int donk_emu_gmm_tc_mstep_f32(int n, int d, int K, const void* mplanes, const double* center, const double* resp,
                              const double* lse_part, double* stats, double* shift, void* work, size_t work_bytes, void*) {
  if (n < 0 || d < 1 || K < 1 || work_bytes < donk_emu_gmm_tc_mwork_bytes(n, d, K)) return DONK_BAD_SHAPE;
  if (!resp) resp = *reinterpret_cast<const double**>((reinterpret_cast<uintptr_t>(work) + 15) & ~(uintptr_t)15);
  const size_t total = gmm_stats_len(d, K);
  for (size_t i = 0; i < total; ++i) stats[i] = 0.0;
  if (n == 0) return DONK_OK;
  const float* x = static_cast<const float*>(mplanes);
  double* nk = stats;
  double* s1 = stats + K;
  double* S2 = stats + (size_t)K * (1 + d);
  std::vector<float> acc((size_t)d * (d + 1));
  for (int k = 0; k < K; ++k) {
    for (size_t r0 = 0; r0 < (size_t)n; r0 += 4096) {
      std::fill(acc.begin(), acc.end(), 0.f);
      const size_t r1 = r0 + 4096 < (size_t)n ? r0 + 4096 : (size_t)n;
      for (size_t r = r0; r < r1; ++r) {
        const float w = (float)resp[r * K + k];
        nk[k] += resp[r * K + k];
        for (int i = 0; i < d; ++i) {
          const float a = w * x[r * d + i];
          acc[(size_t)i * (d + 1) + d] += a;
          for (int j = i; j < d; ++j) acc[(size_t)i * (d + 1) + j] = fmaf(a, x[r * d + j], acc[(size_t)i * (d + 1) + j]);
        }
      }
      for (int i = 0; i < d; ++i) {
        s1[(size_t)k * d + i] += acc[(size_t)i * (d + 1) + d];
        for (int j = i; j < d; ++j) S2[((size_t)k * d + i) * d + j] += acc[(size_t)i * (d + 1) + j];
      }
    }
    for (int i = 0; i < d; ++i)
      for (int j = 0; j < i; ++j) S2[((size_t)k * d + i) * d + j] = S2[((size_t)k * d + j) * d + i];
    for (int j = 0; j < d; ++j) shift[(size_t)k * d + j] = center ? center[j] : 0.0;
  }
  double lse = 0.0;
  for (size_t b = 0; b < ((size_t)n + GMM_TS - 1) / GMM_TS; ++b) lse += lse_part[b];
  stats[total - 1] = lse;
  (void)work;
  return DONK_OK;
}

int donk_emu_gmm_em_step_mode_f64(int n_local, int d, int K, const double* X, const double* weights, const double* means,
                                  const double* prec_chol, const double* resp_in, double* stats, double* shift,
                                  void* workspace, size_t workspace_bytes, int mode, void*) {
  if (n_local < 0 || d < 1 || K < 1 || mode < 0 || mode > 3) return DONK_BAD_SHAPE;
  if (workspace_bytes < donk_emu_gmm_workspace_bytes(n_local, d, K)) return DONK_BAD_SHAPE;
  const size_t total = gmm_stats_len(d, K);
  if (n_local == 0) { memset(stats, 0, total * sizeof(double)); return DONK_OK; }
  GmmArgs<double> a;
  memset(&a, 0, sizeof(a));
  a.n = n_local; a.d = d; a.K = K; a.X = X; a.w = weights; a.means = means; a.pc = prec_chol;
  a.nblocks = (n_local + GMM_TS - 1) / GMM_TS;
  a.nch = gmm_nch(n_local, d, K);
  double* ws = static_cast<double*>(workspace);
  a.resp = ws; a.lse_part = ws + (size_t)n_local * K; a.part = a.lse_part + a.nblocks;
  const GmmMmaPlan pl = gmm_mma_plan(n_local, d, K);
  const bool want_mma = mode == 2 || ((mode == 0 || mode == 3) && (n_local >= 65536 || env_int("DONK_GMM_MMA")));
  const bool mma = pl.ok && !env_int("DONK_GMM_GENERIC") && weights != nullptr && want_mma;  // as the CUDA launcher
  if (mode == 3) {
    if (resp_in) return DONK_BAD_SHAPE;  // the workspace already holds resp and lse_part (donk_emu_gmm_tc_estep_f32)
  } else if (resp_in) {
    a.resp = const_cast<double*>(resp_in);
    memset(a.lse_part, 0, (size_t)a.nblocks * sizeof(double));
  } else if (mma) {
    switch (pl.NB) {
      case 2: emu_gmm_e_mma<2, 2>(a, pl); break;
      case 5: emu_gmm_e_mma<5, 2>(a, pl); break;
      case 6: emu_gmm_e_mma<6, 2>(a, pl); break;
      case 9: emu_gmm_e_mma<9, 2>(a, pl); break;
      case 12: emu_gmm_e_mma<12, 2>(a, pl); break;
      case 15: emu_gmm_e_mma<15, 2>(a, pl); break;
      default: emu_gmm_e_mma<18, 2>(a, pl); break;
    }
  } else {
    emu_gmm_e(a);
  }
  if (mma) {
    a.nch = pl.nch;
    a.shift_out = shift;
    std::vector<double> smem(gmm_m_mma_smem(d, pl.KC) + 2);
    WCtx w{0, 1, 0, 0, pl.Gm};
    const int grid = ((K + pl.KC - 1) / pl.KC) * a.nch;
    for (int cta = 0; cta < grid; ++cta) {
      memset(smem.data(), 0xFF, smem.size() * sizeof(double));
      if (pl.NBW == 3) gmm_m_mma_cta<double, 3, 4>(w, a, cta, smem.data());
      else if (pl.NBW <= 6) gmm_m_mma_cta<double, 6, 2>(w, a, cta, smem.data());
      else gmm_m_mma_cta<double, 11, 2>(w, a, cta, smem.data());
    }
    for (size_t i = 0; i < total; ++i) gmm_reduce_item<double>(a, i, stats);
    return DONK_OK;
  }
  if (shift) memcpy(shift, means, (size_t)K * d * sizeof(double));
  if (gmm_m_threads(d) == 0) return DONK_SMEM_LIMIT;
  std::vector<double> smem(gmm_m_smem(d) + 2);
  Ctx ctx{0, 1};
  for (int cta = 0; cta < K * a.nch; ++cta) {
    memset(smem.data(), 0xFF, smem.size() * sizeof(double));
    gmm_m_cta<double>(ctx, a, cta, smem.data());
  }
  for (size_t i = 0; i < total; ++i) gmm_reduce_item<double>(a, i, stats);
  return DONK_OK;
}

int donk_emu_gmm_em_step_f64(int n_local, int d, int K, const double* X, const double* weights, const double* means,
                             const double* prec_chol, const double* resp_in, double* stats, double* shift,
                             void* workspace, size_t workspace_bytes, void* stream) {
  return donk_emu_gmm_em_step_mode_f64(n_local, d, K, X, weights, means, prec_chol, resp_in, stats, shift, workspace,
                                       workspace_bytes, 0, stream);
}

int donk_emu_gmm_m_finalize_f64(double n_total, int d, int K, const double* stats, const double* shift,
                                double reg_covar, double* weights, double* means, double* covariances,
                                double* prec_chol, double* lower_bound, int* info, void*) {
  GmmFinArgs<double> a;
  a.d = d; a.K = K; a.n_total = n_total; a.reg_covar = reg_covar; a.stats = stats; a.shift = shift;
  a.w = weights; a.means = means; a.cov = covariances; a.pc = prec_chol; a.lower_bound = lower_bound; a.info = info;
  std::vector<double> smem(gmm_fin_smem(d) + 2);
  Ctx ctx{0, 1};
  for (int k = 0; k < K; ++k) {
    memset(smem.data(), 0xFF, smem.size() * sizeof(double));
    gmm_finalize_cta<double>(ctx, a, k, smem.data());
  }
  return DONK_OK;
}

int donk_emu_gmm_predict_proba_f64(int n, int d, int K, const double* X, const double* weights, const double* means,
                                   const double* prec_chol, double* resp, void*) {
  if (n == 0) return DONK_OK;
  GmmArgs<double> a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.d = d; a.K = K; a.X = X; a.w = weights; a.means = means; a.pc = prec_chol;
  a.nblocks = (n + GMM_TS - 1) / GMM_TS;
  a.resp = resp;
  emu_gmm_e(a);
  return DONK_OK;
}

int donk_emu_gmm_prec_chol_from_precisions_f64(int d, int K, const double* precisions, double* prec_chol, int* info,
                                               void*) {
  std::vector<double> smem(gmm_fin_smem(d) + 2);
  Ctx ctx{0, 1};
  for (int k = 0; k < K; ++k) {
    memset(smem.data(), 0xFF, smem.size() * sizeof(double));
    gmm_prec_chol_cta<double>(ctx, d, precisions, prec_chol, info, k, smem.data());
  }
  return DONK_OK;
}

}  // extern "C"

extern "C" {

size_t donk_emu_kmeans_workspace_bytes(int d, int K) { return (size_t)444 * ((size_t)K * d + K) * sizeof(double); }

int donk_emu_kmeans_pass_f64(int n, int d, int K, const double* X, const double* centers, int* labels, double* mind2,
                             double* dist, double* sums, void* workspace, size_t workspace_bytes, void*) {
  if (sums && workspace_bytes < donk_emu_kmeans_workspace_bytes(d, K)) return DONK_BAD_SHAPE;
  const int len = K * d + K, ncta = 3;
  std::vector<double> smem(kmeans_pass_smem(d, K) + 2);
  Ctx ctx{0, 1};
  double* part = sums ? static_cast<double*>(workspace) : nullptr;
  for (int cta = 0; cta < ncta; ++cta) {
    memset(smem.data(), 0xFF, smem.size() * sizeof(double));
    kmeans_pass_cta<double>(ctx, n, d, K, X, centers, labels, mind2, dist, part, cta, ncta, smem.data());
  }
  if (sums)
    for (int i = 0; i < len; ++i) kmeans_reduce_item<double>(ncta, len, part, sums, i);
  return DONK_OK;
}


int donk_emu_kmeanspp_round_f64(int n, int d, int Kc, const double* X, const double* cand, const double* closest,
                                double* dmin, double* pot, void* workspace, size_t workspace_bytes, void*) {
  if (n < 1 || d < 1 || Kc < 1 || Kc > 16) return DONK_BAD_SHAPE;
  if (workspace_bytes < donk_emu_kmeans_workspace_bytes(d, Kc)) return DONK_BAD_SHAPE;
  (void)workspace;
  int rc = donk_emu_kmeans_pass_f64(n, d, Kc, X, cand, nullptr, nullptr, dmin, nullptr, nullptr, 0, nullptr);
  if (rc != DONK_OK) return rc;
  for (int k = 0; k < Kc; ++k) pot[k] = 0.0;
  for (size_t r = 0; r < (size_t)n; ++r) {
    kmeanspp_min_item<double>(Kc, closest, dmin, r);
    for (int k = 0; k < Kc; ++k) pot[k] += dmin[r * Kc + k];
  }
  return DONK_OK;
}

}  // extern "C"
