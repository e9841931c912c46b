// donk_b200 — C ABI, part: ILQR.optimize (donk/traj_opt/lqg.py:90-123) as ONE CUDA-graph launch.
//
// graph:  reset, ctl(init) -> [step(min_eta) ctl] [step(max_eta) ctl] [step(e^xa) ctl] [step(e^xb) ctl]
//               -> WHILE(any condition still searching) { step(eta) ctl }
//               -> ctl(final) -> step(eta_final)
// The controller kernel (optimize.cuh) sets the WHILE condition on the device, so the host neither reads a KL value nor
// counts active conditions while the search runs.  Instantiated graphs are cached by their full argument set.
#include <mutex>
#include <vector>

#include "abi_common.cuh"
#include "optimize.cuh"

using namespace donk;
using namespace donk_abi;

namespace donk_abi {
int ilqr_step_launch_f64(const IlqrArgs<double>& a, void* stream);  // abi_ilqr.cu
}

namespace {

__global__ void optimize_ctl_kernel(const OptArgs<double> a, int phase, cudaGraphConditionalHandle handle, int use_handle) {
  __shared__ int s_act, s_err, s_last;
  if (threadIdx.x == 0) { s_act = 0; s_err = 0; }
  __syncthreads();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < a.B) {
    int err = 0;
    const int act = optimize_ctl_item<double>(a, phase, b, &err);
    if (act) atomicAdd(&s_act, 1);
    if (err) atomicOr(&s_err, err);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    OptCtl* c = a.ctl;
    if (s_act) atomicAdd(&c->acc, s_act);
    if (s_err) atomicOr(&c->err, s_err);
    __threadfence();
    s_last = atomicAdd(&c->ticket, 1) == (int)gridDim.x - 1;
    if (s_last) {  // the round is complete: publish the count and decide whether the loop goes on
      __threadfence();
      const int n = atomicExch(&c->acc, 0);
      c->ticket = 0;
      c->n_active = n;
      if (phase == OPT_MIN) c->n_after_min = n;
      if (phase == OPT_LOOP) c->loop_iters += 1;
      // maxiter + 2 rounds always suffice (brent_item gives up after maxiter): a guard against a corrupted record
      if (use_handle && (phase == OPT_FB || phase == OPT_LOOP))
        cudaGraphSetConditional(handle, (n > 0 && c->loop_iters <= a.maxiter + 2) ? 1u : 0u);
    }
  }
}

struct PlanKey {
  int dev, B, T, dX, dU, maxiter;
  const void* ptr[32];
  double sc[8];
};

struct Plan {
  PlanKey key;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  unsigned long long stamp = 0;
  int launches_fixed = 0;  // kernels outside the loop body
};

std::mutex g_mu;
std::vector<Plan> g_plans;
unsigned long long g_stamp = 0;
cudaStream_t g_capture_stream[64] = {nullptr};

int ctl_launch(const OptArgs<double>& o, int phase, cudaGraphConditionalHandle h, int use_handle, cudaStream_t st) {
  optimize_ctl_kernel<<<(o.B + 255) / 256, 256, 0, st>>>(o, phase, h, use_handle);
  LAUNCHED("optimize_ctl_kernel");
  return DONK_OK;
}

// one evaluation round: the batched step on the current multipliers, then the controller
int round_launch(const IlqrArgs<double>& s, const OptArgs<double>& o, int phase, cudaGraphConditionalHandle h,
                 cudaStream_t st) {
  int rc = ilqr_step_launch_f64(s, st);
  if (rc != DONK_OK) return rc;
  return ctl_launch(o, phase, h, 1, st);
}

int build_plan(Plan& pl, const IlqrArgs<double>& search, const IlqrArgs<double>& final_step, const OptArgs<double>& o) {
  cudaStream_t cs = g_capture_stream[pl.key.dev & 63];
  if (!cs) {
    CU(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    g_capture_stream[pl.key.dev & 63] = cs;
  }
  CU(cudaGraphCreate(&pl.graph, 0));
  cudaGraphConditionalHandle h;
  CU(cudaGraphConditionalHandleCreate(&h, pl.graph, 0, cudaGraphCondAssignDefault));
  int rc = DONK_OK;
  // -- prefix: the four fixed evaluations ----------------------------------------------------------------
  CU(cudaStreamBeginCaptureToGraph(cs, pl.graph, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
  cudaError_t e = cudaMemsetAsync(o.ctl, 0, sizeof(OptCtl), cs);
  if (e == cudaSuccess) rc = ctl_launch(o, OPT_INIT, h, 0, cs);
  for (int ph = OPT_MIN; ph <= OPT_FB && rc == DONK_OK && e == cudaSuccess; ++ph) rc = round_launch(search, o, ph, h, cs);
  cudaGraph_t same = nullptr;
  cudaError_t e2 = cudaStreamEndCapture(cs, &same);
  if (rc != DONK_OK) return rc;
  if (e != cudaSuccess) return cuda_fail(e, "optimize: capture prefix");
  if (e2 != cudaSuccess) return cuda_fail(e2, "optimize: end capture (prefix)");
  // the prefix is a chain: its single leaf is what the loop depends on
  size_t n_nodes = 0;
  CU(cudaGraphGetNodes(pl.graph, nullptr, &n_nodes));
  std::vector<cudaGraphNode_t> nodes(n_nodes);
  CU(cudaGraphGetNodes(pl.graph, nodes.data(), &n_nodes));
  std::vector<cudaGraphNode_t> leaves;
  for (cudaGraphNode_t nd : nodes) {
    size_t n_dep = 0;
    CU(cudaGraphNodeGetDependentNodes(nd, nullptr, &n_dep));
    if (n_dep == 0) leaves.push_back(nd);
  }
  // -- the Brent loop ---------------------------------------------------------------------------------------
  cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
  cp.conditional.handle = h;
  cp.conditional.type = cudaGraphCondTypeWhile;
  cp.conditional.size = 1;
  cudaGraphNode_t loop;
  CU(cudaGraphAddNode(&loop, pl.graph, leaves.data(), leaves.size(), &cp));
  cudaGraph_t body = cp.conditional.phGraph_out[0];
  CU(cudaStreamBeginCaptureToGraph(cs, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
  rc = round_launch(search, o, OPT_LOOP, h, cs);
  e2 = cudaStreamEndCapture(cs, &same);
  if (rc != DONK_OK) return rc;
  if (e2 != cudaSuccess) return cuda_fail(e2, "optimize: end capture (loop body)");
  // -- suffix: final multipliers, final step ----------------------------------------------------------------
  CU(cudaStreamBeginCaptureToGraph(cs, pl.graph, &loop, nullptr, 1, cudaStreamCaptureModeRelaxed));
  rc = ctl_launch(o, OPT_FINAL, h, 0, cs);
  if (rc == DONK_OK) rc = ilqr_step_launch_f64(final_step, cs);
  e2 = cudaStreamEndCapture(cs, &same);
  if (rc != DONK_OK) return rc;
  if (e2 != cudaSuccess) return cuda_fail(e2, "optimize: end capture (suffix)");
  CU(cudaGraphInstantiate(&pl.exec, pl.graph, 0));
  return DONK_OK;
}

void destroy_plan(Plan& pl) {
  if (pl.exec) cudaGraphExecDestroy(pl.exec);
  if (pl.graph) cudaGraphDestroy(pl.graph);
  pl.exec = nullptr;
  pl.graph = nullptr;
}

}  // namespace

extern "C" {

size_t donk_ilqr_optimize_workspace_bytes(int B) { return optimize_workspace_bytes(B); }

int donk_ilqr_optimize_f64(int B, int T, int dX, int dU, const double* F, const double* f, const double* dyn_covar,
                           const double* C, const double* c, const double* cc, const double* C_kl, const double* c_kl,
                           const double* prevK, const double* prevk, const double* prevLam, const double* prev_hld,
                           const double* x0_mean, const double* x0_covar, const double* kl_step, double min_eta,
                           double max_eta, double eta_a, double eta_b, double xa, double xb, double rtol, int maxiter,
                           double* eta, int* active, double* K, double* k, double* pol_covar, double* pol_inv_covar,
                           double* kl, double* cost, double* traj_mean, double* traj_covar, int* info, int* status,
                           int* counters, void* work, size_t work_bytes, void* stream) {
  if (B < 0 || T < 1 || dX < 1 || dU < 1 || maxiter < 1) return DONK_BAD_SHAPE;
  if (work_bytes < optimize_workspace_bytes(B)) {
    snprintf(g_err, sizeof(g_err), "ilqr_optimize: workspace of %zu bytes, %zu needed", work_bytes, optimize_workspace_bytes(B));
    return DONK_BAD_SHAPE;
  }
  if (B == 0) return DONK_OK;
  IlqrArgs<double> s;
  memset(&s, 0, sizeof(s));
  s.B = B; s.E = 1; s.T = T; s.dX = dX; s.dU = dU;
  s.flags = ILQR_BACKWARD | ILQR_FORWARD | ILQR_KL | ILQR_COST;
  s.F = F; s.f = f; s.dyn_covar = dyn_covar; s.C = C; s.c = c; s.cc = cc; s.C_kl = C_kl; s.c_kl = c_kl;
  s.prevK = prevK; s.prevk = prevk; s.prevLam = prevLam; s.prev_hld = prev_hld;
  s.x0m = x0_mean; s.x0S = x0_covar; s.eta = eta; s.gamma = 1.0; s.fwd_reg = 1e-6;
  s.K = K; s.k = k; s.pcov = pol_covar; s.pinv = pol_inv_covar; s.kl = kl; s.cost = cost;
  s.info = info; s.active = active;
  s.aligned16 = (((uintptr_t)F | (uintptr_t)C_kl | (uintptr_t)dyn_covar | (uintptr_t)C | (uintptr_t)K) & 15) == 0;
  IlqrArgs<double> fin = s;  // the returned step: every condition, trajectory outputs if asked for
  fin.tmean = traj_mean; fin.tcov = traj_covar;
  uintptr_t wp = ((uintptr_t)work + 15) & ~(uintptr_t)15;
  OptArgs<double> o;
  memset(&o, 0, sizeof(o));
  o.B = B; o.kl_step = kl_step; o.min_eta = min_eta; o.max_eta = max_eta; o.eta_a = eta_a; o.eta_b = eta_b;
  o.xa = xa; o.xb = xb; o.xtol = 2e-12; o.rtol = rtol; o.maxiter = maxiter;
  o.kl = kl; o.info = info; o.eta = eta; o.active = active; o.status = status;
  o.st = reinterpret_cast<OptState<double>*>(wp);
  o.ctl = reinterpret_cast<OptCtl*>(wp + (size_t)B * sizeof(OptState<double>));

  PlanKey key;
  memset(&key, 0, sizeof(key));
  cudaGetDevice(&key.dev);
  key.B = B; key.T = T; key.dX = dX; key.dU = dU; key.maxiter = maxiter;
  const void* ptrs[] = {F, f, dyn_covar, C, c, cc, C_kl, c_kl, prevK, prevk, prevLam, prev_hld, x0_mean, x0_covar,
                        kl_step, eta, active, K, k, pol_covar, pol_inv_covar, kl, cost, traj_mean, traj_covar, info,
                        status, work};
  static_assert(sizeof(ptrs) / sizeof(ptrs[0]) <= 32, "PlanKey::ptr too small");
  for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); ++i) key.ptr[i] = ptrs[i];
  const double scs[] = {min_eta, max_eta, eta_a, eta_b, xa, xb, rtol};
  for (size_t i = 0; i < sizeof(scs) / sizeof(scs[0]); ++i) key.sc[i] = scs[i];

  cudaGraphExec_t exec = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_mu);
    Plan* hit = nullptr;
    if (!env_int("DONK_OPT_NOCACHE"))
      for (Plan& p : g_plans)
        if (memcmp(&p.key, &key, sizeof(key)) == 0) hit = &p;
    if (!hit) {
      const size_t cap = 16;
      if (g_plans.size() >= cap) {  // evict the least recently used graph
        size_t old = 0;
        for (size_t i = 1; i < g_plans.size(); ++i)
          if (g_plans[i].stamp < g_plans[old].stamp) old = i;
        destroy_plan(g_plans[old]);
        g_plans.erase(g_plans.begin() + old);
      }
      Plan pl;
      pl.key = key;
      const int rc = build_plan(pl, s, fin, o);
      if (rc != DONK_OK) { destroy_plan(pl); return rc; }
      g_plans.push_back(pl);
      hit = &g_plans.back();
    }
    hit->stamp = ++g_stamp;
    exec = hit->exec;
  }
  CU(cudaGraphLaunch(exec, (cudaStream_t)stream));
  // counters (device): loop rounds, error bits, conditions searched — read by the caller together with status
  CU(cudaMemcpyAsync(counters, &o.ctl->loop_iters, 3 * sizeof(int), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return DONK_OK;
}

}  // extern "C"
