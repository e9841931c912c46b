// donk_b200 — the steps either side of the hot path: quadratic cost constructors
// (donk/costs/quadratic_costs.py:121-166) and transition log-probabilities under fitted dynamics
// (donk/dynamics/linear_dynamics.py:54-64), batched over conditions.  Element-wise, HBM-bound.
#pragma once
#include "common.cuh"

namespace donk {

// L(xu) = 0.5 sum_i w_i (t_i - xu_i)^2 as (C, c, cc); exact.  One item per element of C: idx in [0, n p p).
template <typename R>
DONK_HD void cost_l2_item(int p, const R* t, const R* w, R* C, R* c, R* cc, size_t idx) {
  const int col = (int)(idx % p);
  const size_t ir = idx / p;
  const int r = (int)(ir % p);
  const size_t i = ir / p;
  const R wr = w[i * p + r];
  C[idx] = r == col ? wr : R(0);
  if (col == 0) c[i * p + r] = -(wr * t[i * p + r]);
  if (col == 0 && r == 0) {
    R acc = R(0);
    for (int j = 0; j < p; ++j) acc += t[i * p + j] * w[i * p + j] * t[i * p + j];
    cc[i] = acc / 2;
  }
}

// L(xu) = sum_i w_i sqrt((t_i - xu_i)^2 + alpha): second-order Taylor expansion at xu.
template <typename R>
DONK_HD void cost_l1_item(int p, const R* xu, const R* t, const R* w, R alpha, R* C, R* c, R* cc, size_t idx) {
  const int col = (int)(idx % p);
  const size_t ir = idx / p;
  const int r = (int)(ir % p);
  const size_t i = ir / p;
  const size_t e = i * p + r;
  const R sq = (t[e] - xu[e]) * (t[e] - xu[e]) + alpha;
  const R ad = sqrt(sq);
  const R cu = sq * ad;
  C[idx] = r == col ? alpha * w[e] / cu : R(0);
  if (col == 0) c[e] = w[e] * ((xu[e] - t[e]) / ad - alpha * xu[e] / cu);
  if (col == 0 && r == 0) {
    R acc = R(0);
    for (int j = 0; j < p; ++j) {
      const size_t q = i * p + j;
      const R s2 = (t[q] - xu[q]) * (t[q] - xu[q]) + alpha;
      const R a2 = sqrt(s2);
      const R c2 = s2 * a2;
      acc += w[q] * (a2 + xu[q] * (t[q] - xu[q]) / a2 + alpha * xu[q] * xu[q] / c2 / 2);
    }
    cc[i] = acc;
  }
}

// log N(x_{t+1}; F_t [x_t; u_t] + f_t, covar_t) for sample n of condition b at time t.
// X (B,N,T+1,dX), U (B,N,T,dU), F (B,T,dX,p), f (B,T,dX), Sinv (B,T,dX,dX), hld (B,T) = 0.5 log det covar.
// diff: dX reals of scratch owned by the caller.
template <typename R>
DONK_HD R dyn_log_prob_item(int N, int T, int dX, int dU, const R* X, const R* U, const R* F, const R* f,
                            const R* Sinv, const R* hld, size_t idx, R* diff) {
  const int t = (int)(idx % T);
  const size_t bn = idx / T;
  const size_t b = bn / N;
  const int p = dX + dU;
  const R* x = X + (bn * (T + 1) + t) * dX;
  const R* xn = x + dX;
  const R* u = U + (bn * T + t) * dU;
  const size_t bt = b * T + t;
  const R* Ft = F + bt * dX * p;
  for (int i = 0; i < dX; ++i) {
    R m = f[bt * dX + i];
    for (int j = 0; j < dX; ++j) m = fma(Ft[i * p + j], x[j], m);
    for (int j = 0; j < dU; ++j) m = fma(Ft[i * p + dX + j], u[j], m);
    diff[i] = xn[i] - m;
  }
  const R* S = Sinv + bt * dX * dX;
  R maha = R(0);
  for (int i = 0; i < dX; ++i) {
    R row = R(0);
    for (int j = 0; j < dX; ++j) row = fma(S[i * dX + j], diff[j], row);
    maha = fma(diff[i], row, maha);
  }
  return R(-0.5) * (R(dX) * R(1.8378770664093454835606594728112) + 2 * hld[bt] + maha);
}

}  // namespace donk
