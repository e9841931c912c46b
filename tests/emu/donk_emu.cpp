// Host emulation of the donk_b200 kernel bodies — TEST INFRASTRUCTURE ONLY.
//
// Compiles donk_b200/csrc/*.cuh with -DDONK_EMU (see common.cuh): one host thread plays a
// whole CTA, phases run sequentially (forward or reversed order, donk_emu_set_reverse) and
// shared memory is a heap buffer.  It exists so kernel logic can be checked against the
// oracle in the build container, which has no GPU.  The product (donk_b200/_lib.py) never
// loads this library; entry points carry the donk_emu_ prefix and take HOST pointers.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../donk_b200/csrc/fitlr.cuh"
#include "../../donk_b200/csrc/gmm.cuh"
#include "../../donk_b200/csrc/ilqr.cuh"

namespace donk {
int g_emu_reverse = 0;
}
using namespace donk;

static const size_t kSmemCap = 227 * 1024;

static int env_int(const char* name) {
  const char* v = getenv(name);
  return v ? atoi(v) : 0;
}

extern "C" {

void donk_emu_set_reverse(int r) { g_emu_reverse = r; }

int donk_emu_ilqr_step_f64(int B, int E, int T, int dX, int dU, int flags, const double* F, const double* f,
                           const double* dyn_covar, const double* C, const double* c, const double* cc,
                           const double* C_kl, const double* c_kl, const double* prevK, const double* prevk,
                           const double* prevLam, const double* prev_hld, const double* x0_mean,
                           const double* x0_covar, const double* eta, double gamma, double fwd_reg,
                           const int* active, double* K, double* k, double* pol_covar, double* pol_inv_covar,
                           double* kl, double* cost, double* traj_mean, double* traj_covar, int* info, void*) {
  IlqrArgs<double> a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.E = E; a.T = T; a.dX = dX; a.dU = dU; a.flags = flags;
  a.F = F; a.f = f; a.dyn_covar = dyn_covar; a.C = C; a.c = c; a.cc = cc; a.C_kl = C_kl; a.c_kl = c_kl;
  a.prevK = prevK; a.prevk = prevk; a.prevLam = prevLam; a.prev_hld = prev_hld;
  a.x0m = x0_mean; a.x0S = x0_covar; a.eta = eta; a.gamma = gamma; a.fwd_reg = fwd_reg;
  a.K = K; a.k = k; a.pcov = pol_covar; a.pinv = pol_inv_covar; a.kl = kl; a.cost = cost;
  a.tmean = traj_mean; a.tcov = traj_covar; a.info = info; a.active = active;
  if (B < 0 || E < 1 || T < 1 || dX < 1 || dU < 1) return DONK_BAD_SHAPE;
  if ((flags & ILQR_KL) && (flags & (ILQR_BACKWARD | ILQR_FORWARD)) != (ILQR_BACKWARD | ILQR_FORWARD))
    return DONK_BAD_SHAPE;
  if (B == 0) return DONK_OK;
  IlqrPlan pl;
  int rc = ilqr_plan(B, E, dX, dU, sizeof(double), kSmemCap, env_int("DONK_ILQR_SLOTS"), 0, pl);
  if (rc != DONK_OK) return rc;
  a.S = pl.S;
  a.ES = pl.ES;
  std::vector<double> smem(pl.smem / sizeof(double) + 2);
  Ctx ctx{0, 1};
  for (int cta = 0; cta < pl.grid; ++cta) {
    // poison shared memory: kernels must not depend on its initial contents
    memset(smem.data(), 0xFF, smem.size() * sizeof(double));
    ilqr_cta<double>(ctx, a, cta, smem.data());
  }
  return DONK_OK;
}

int donk_emu_extended_costs_kl_f64(int n, int dX, int dU, const double* K, const double* k, const double* Lam,
                                   double* C_kl, double* c_kl, void*) {
  const int p = dX + dU;
  for (size_t idx = 0; idx < (size_t)n * p; ++idx) extended_costs_item<double>(idx / p, (int)(idx % p), dX, dU, K, k, Lam, C_kl, c_kl);
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
