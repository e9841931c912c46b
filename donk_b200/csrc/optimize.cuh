// donk_b200 — ILQR.optimize as a device-side state machine (donk/traj_opt/lqg.py:90-123).
//
// The reference runs, per condition: step(min_eta) -> "constraint inactive?" -> step(max_eta) -> "max_eta too low?" ->
// scipy.optimize.brentq on g(x) = kl(step(exp x)) - kl_step (which itself evaluates g(log min_eta), g(log max_eta)
// first) -> step(exp root).  Here all B conditions advance in lock-step; after every batched step kernel one small
// controller kernel consumes the fresh KL values, advances each condition's record and writes the next multipliers.
// No decision needs the host: the rounds of the Brent loop are the body of a CUDA-graph WHILE node whose condition the
// controller sets from the count of still-searching conditions (abi_optimize.cu).
#pragma once
#include "ilqr.cuh"

namespace donk {

enum OptPhase { OPT_INIT = -1, OPT_MIN = 0, OPT_MAX = 1, OPT_FA = 2, OPT_FB = 3, OPT_LOOP = 4, OPT_FINAL = 5 };

// error bits collected over all conditions (host raises what the reference would raise)
enum OptErr { OPT_ERR_NOT_PD = 1, OPT_ERR_MAX_ETA = 2, OPT_ERR_SIGN = 4, OPT_ERR_MAXITER = 8 };

// per-condition status: 0 root found, 1 constraint inactive at min_eta (lqg.py:108), 2 max_eta too low (lqg.py:112),
// 3 Brent failed (same sign / maxiter), 4 Quu not positive definite in one of the evaluations
template <typename R>
struct OptState {
  BrentState<R> br;  // 12 reals
  R fa;              // g(log min_eta), kept until g(log max_eta) arrives
  R pad[3];
};

// control block of one optimisation (ints in device memory)
struct OptCtl {
  int n_active;      // conditions still searching after the last controller round
  int acc;           // accumulator of the round in flight
  int ticket;        // CTAs of the round that have finished
  int loop_iters;    // rounds of the Brent loop executed
  int err;           // OptErr bits
  int n_after_min;   // conditions whose constraint is active at min_eta
  int pad[2];
};

template <typename R>
struct OptArgs {
  int B;
  const R* kl_step;  // (B)
  R min_eta, max_eta, eta_a, eta_b, xa, xb, xtol, rtol;
  int maxiter;
  const R* kl;       // (B) written by the step kernel (E = 1)
  const int* info;   // (B) written by the step kernel
  R* eta;            // (B) multipliers of the next step
  int* active;       // (B) 1 = evaluate this condition in the next step
  int* status;       // (B)
  OptState<R>* st;   // (B)
  OptCtl* ctl;
};

// One condition, one controller round.  Returns 1 while the condition keeps searching.
template <typename R>
DONK_HD int optimize_ctl_item(const OptArgs<R>& a, int phase, int b, int* err) {
  OptState<R>& st = a.st[b];
  if (phase == OPT_INIT) {  // before the first evaluation: every condition at min_eta
    a.eta[b] = a.min_eta;
    a.active[b] = 1;
    return 0;
  }
  if (phase == OPT_FINAL) {
    // the multiplier of the returned policy: the root, or min_eta where the search never ran / failed
    if (!(a.status[b] == 0 && st.br.status == R(1))) a.eta[b] = a.min_eta;
    a.active[b] = 1;
    return 0;
  }
  int act = phase == OPT_MIN ? 1 : a.active[b];
  if (phase == OPT_MIN) { a.status[b] = 0; st.br.status = R(0); }
  if (act && a.info[b] != 0) {  // numpy.linalg.LinAlgError in the reference
    a.status[b] = 4; act = 0; *err |= OPT_ERR_NOT_PD;
  }
  const R g = act ? a.kl[b] - a.kl_step[b] : R(0);
  R next = a.eta[b];
  if (phase == OPT_MIN) {
    if (act && a.kl[b] <= a.kl_step[b]) { a.status[b] = 1; act = 0; }  // lqg.py:108
    next = a.max_eta;
  } else if (phase == OPT_MAX) {
    if (act && a.kl[b] > a.kl_step[b]) { a.status[b] = 2; act = 0; *err |= OPT_ERR_MAX_ETA; }  // lqg.py:112
    next = a.eta_a;
  } else if (phase == OPT_FA) {
    st.fa = g;
    next = a.eta_b;
  } else if (act) {
    int go = 0;
    brent_item<R>(st.br, phase == OPT_FB ? 0 : 1, a.xa, a.xb, st.fa, g, g, a.xtol, a.rtol, a.maxiter, &next, &go);
    if (!go) {
      if (st.br.status == R(2)) { a.status[b] = 3; *err |= OPT_ERR_SIGN; }
      else if (st.br.status == R(3)) { a.status[b] = 3; *err |= OPT_ERR_MAXITER; }
    }
    act = go;
  }
  a.eta[b] = next;
  a.active[b] = act;
  return act;
}

DONK_HD size_t optimize_workspace_bytes(int B) {
  return (size_t)(B > 0 ? B : 1) * sizeof(OptState<double>) + sizeof(OptCtl) + 64;
}

}  // namespace donk
