/* donk_b200 — C ABI of the B200-native trajectory-optimisation hot path.
 *
 * The reference (DiddiZ/donk.ai) is pure Python and has no FFI layer; the boundary it
 * exposes for this path is its Python API.  Each entry point below replaces the body of
 * one reference function (cited file:line, relative to the reference root) and is what
 * a binding of that function would call (see INTEGRATION.md for the ctypes stubs).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to C-contiguous memory owned by the caller;
 *     the library never allocates or frees caller-visible memory;
 *   - `_f64` entry points take double, `_f32` float (same layouts);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls
 *     are asynchronous with respect to the host unless stated;
 *   - return value: DONK_OK, or an error code; numerical failures that depend on the
 *     data (a matrix that is not positive definite) are reported per problem through
 *     the `info` device arrays, not through the return value;
 *   - no global state; re-entrant; thread-safe per stream.
 */
#ifndef DONK_B200_H
#define DONK_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DONK_OK 0
#define DONK_BAD_SHAPE 1
#define DONK_NOT_PD 2
#define DONK_CUDA_ERROR 3
#define DONK_SMEM_LIMIT 4

/* flags of donk_ilqr_step_* */
#define DONK_ILQR_BACKWARD 1
#define DONK_ILQR_FORWARD 2
#define DONK_ILQR_KL 4
#define DONK_ILQR_COST 8

/* Library / build identification: "donk_b200 <version> sm_100a". */
const char* donk_version(void);
/* Last CUDA error string seen by this library on the calling thread ("" if none). */
const char* donk_last_error(void);
/* Number of kernels this library has launched since load (for bench.py's gpu_launches). */
unsigned long long donk_launch_count(void);

/* ---- iLQR ------------------------------------------------------------------------------ */

/* ILQR.step batched over B conditions x E eta candidates:  donk/traj_opt/lqg.py:55-75
 * (backward lqg.py:126-182, forward lqg.py:185-244, kl_divergence_action lqg.py:276-306,
 * QuadraticCosts.expected_costs donk/costs/quadratic_costs.py:106-119).
 *
 *   p = dX+dU.  Inputs per condition b:
 *     F (B,T,dX,p)  f (B,T,dX)  dyn_covar (B,T,dX,dX)           LinearDynamics
 *     C (B,T+1,p,p) c (B,T+1,p) cc (B,T+1)                      QuadraticCosts
 *     C_kl (B,T,p,p) c_kl (B,T,p)     extended costs of prev_pol (may be NULL: plain LQR)
 *     prevK (B,T,dU,dX) prevk (B,T,dU) prevLam (B,T,dU,dU) = prev_pol.inv_covar,
 *     prev_hld (B,T) = sum log diag chol(prev_pol.covar)        (needed with DONK_ILQR_KL)
 *     x0_mean (B,dX) x0_covar (B,dX,dX)                         StateDistribution
 *     eta (B,E)   Lagrange multipliers (NULL: eta = 1, costs used as given)
 *     gamma       discount (backward's `gamma`), fwd_reg = forward's `regularization`
 *     active (B) int32, optional: conditions with 0 are skipped (outputs untouched)
 *   Outputs per (b,e):
 *     K (B,E,T,dU,dX) k (B,E,T,dU) pol_covar, pol_inv_covar (B,E,T,dU,dU)   always written
 *       (with DONK_ILQR_FORWARD but not DONK_ILQR_BACKWARD they are inputs: the policy)
 *     kl (B,E), cost (B,E)          with DONK_ILQR_KL / DONK_ILQR_COST
 *     traj_mean (B,E,T+1,p), traj_covar (B,E,T+1,p,p)   optional (NULL to skip)
 *     info (B,E) int32: 0, or 1+t of the first step whose Quu was not positive definite
 *       (the reference raises numpy.linalg.LinAlgError there, lqg.py:165)
 *   DONK_ILQR_KL requires DONK_ILQR_BACKWARD|DONK_ILQR_FORWARD (the log-determinants of the
 *   new policy come from the backward factorisation).
 */
int donk_ilqr_step_f64(int B, int E, int T, int dX, int dU, int flags,
                       const double* F, const double* f, const double* dyn_covar,
                       const double* C, const double* c, const double* cc,
                       const double* C_kl, const double* c_kl,
                       const double* prevK, const double* prevk, const double* prevLam, const double* prev_hld,
                       const double* x0_mean, const double* x0_covar,
                       const double* eta, double gamma, double fwd_reg, const int* active,
                       double* K, double* k, double* pol_covar, double* pol_inv_covar,
                       double* kl, double* cost, double* traj_mean, double* traj_covar,
                       int* info, void* stream);
int donk_ilqr_step_f32(int B, int E, int T, int dX, int dU, int flags,
                       const float* F, const float* f, const float* dyn_covar,
                       const float* C, const float* c, const float* cc,
                       const float* C_kl, const float* c_kl,
                       const float* prevK, const float* prevk, const float* prevLam, const float* prev_hld,
                       const float* x0_mean, const float* x0_covar,
                       const float* eta, float gamma, float fwd_reg, const int* active,
                       float* K, float* k, float* pol_covar, float* pol_inv_covar,
                       float* kl, float* cost, float* traj_mean, float* traj_covar,
                       int* info, void* stream);

/* extended_costs_kl: donk/traj_opt/lqg.py:247-273.  n = number of (b,t) pairs (B*T).
 *   K (n,dU,dX) k (n,dU) Lam (n,dU,dU) -> C_kl (n,p,p) c_kl (n,p). */
int donk_extended_costs_kl_f64(int n, int dX, int dU, const double* K, const double* k, const double* Lam,
                               double* C_kl, double* c_kl, void* stream);

/* Lazy chol_covar / inv_covar of TimeVaryingLinearGaussian: donk/models/tvlg.py:62-99,
 * donk/utils/batched.py:6-49.  A (n,d,d) SPD -> Ainv (n,d,d), chol (n,d,d) lower, half_logdet (n) =
 * sum log diag chol; any output may be NULL.  info (n) int32: 1 where A is not positive definite
 * (numpy.linalg.cholesky raises LinAlgError there). */
int donk_spd_factor_f64(int n, int d, const double* A, double* Ainv, double* chol, double* half_logdet,
                        int* info, void* stream);

/* kl_divergence_action for given states: donk/traj_opt/lqg.py:276-306.
 *   X (B,T,dX); new policy K,k,covar + hld (B,T); previous policy pK,pk,pLam + phld (B,T) -> kl (B). */
int donk_kl_divergence_action_f64(int B, int T, int dX, int dU, const double* X,
                                  const double* K, const double* k, const double* covar, const double* hld,
                                  const double* pK, const double* pk, const double* pLam, const double* phld,
                                  double* kl, void* stream);

/* Brent root search of ILQR.optimize, one state machine per condition, advanced in lock-step:
 * donk/traj_opt/lqg.py:116-123 + scipy.optimize.brentq (xtol=2e-12, maxiter=100 defaults).
 *   phase 0: fa (B) = kl(step(exp xa)) - kl_step, fb (B) likewise at xb; initialises state.
 *   phase 1: fnew (B) = kl(step(eta)) of the eta written by the previous call.
 *   state (B,12) doubles (opaque), eta (B) next multiplier to evaluate (or the root once
 *   converged), active (B) int32 1 while the search of that condition continues,
 *   status (B) int32: 0 running, 1 converged, 2 f(a),f(b) same sign, 3 maxiter,
 *   n_active (1) int32 device counter of still-active conditions.
 *   kl_step is subtracted from fa/fb/fnew inside. */
int donk_brent_update_f64(int B, int phase, double xa, double xb, const double* fa, const double* fb,
                          const double* fnew, double kl_step, double xtol, double rtol, int maxiter,
                          double* state, double* eta, int* active, int* status, int* n_active, void* stream);

/* ILQR.optimize for B conditions in lock-step, entirely on the device: donk/traj_opt/lqg.py:90-123 (the constraint
 * check at min_eta, the feasibility check at max_eta, scipy.optimize.brentq in log space, the step at the root).
 * One CUDA-graph launch: four fixed evaluation rounds, a WHILE node around the Brent rounds whose condition a small
 * controller kernel sets from the number of still-searching conditions, the final step.  Inputs as donk_ilqr_step_f64
 * (E = 1, gamma = 1, forward regularisation 1e-6); kl_step (B) per-condition KL bound.
 *   eta_a = exp(xa), eta_b = exp(xb) with xa = log(min_eta), xb = log(max_eta): the multipliers brentq evaluates first
 *   (passed in so that they are the host libm's values, as in the reference).
 *   eta (B) out: the multiplier of the returned policy; active (B) int32 scratch;
 *   K,k,pol_covar,pol_inv_covar,kl,cost,traj_mean,traj_covar,info: results of the final step (E = 1);
 *   status (B) int32: 0 root found, 1 constraint inactive at min_eta, 2 max_eta too low (ValueError in the reference),
 *                     3 brentq failed (same sign / maxiter), 4 Quu not positive definite (LinAlgError);
 *   counters (3) int32 device: Brent rounds executed, OR of error bits (1 not PD, 2 max_eta, 4 sign, 8 maxiter),
 *                     number of conditions whose constraint is active at min_eta;
 *   work: donk_ilqr_optimize_workspace_bytes(B) bytes of device scratch. */
size_t donk_ilqr_optimize_workspace_bytes(int B);
int donk_ilqr_optimize_f64(int B, int T, int dX, int dU, const double* F, const double* f, const double* dyn_covar,
                           const double* C, const double* c, const double* cc, const double* C_kl, const double* c_kl,
                           const double* prevK, const double* prevk, const double* prevLam, const double* prev_hld,
                           const double* x0_mean, const double* x0_covar, const double* kl_step, double min_eta,
                           double max_eta, double eta_a, double eta_b, double xa, double xb, double rtol, int maxiter,
                           double* eta, int* active, double* K, double* k, double* pol_covar, double* pol_inv_covar,
                           double* kl, double* cost, double* traj_mean, double* traj_covar, int* info, int* status,
                           int* counters, void* work, size_t work_bytes, void* stream);

/* Quadratic cost constructors, batched: donk/costs/quadratic_costs.py:121-139 (L2, exact) and :141-166 (L1,
 * second-order expansion at xu).  t, w[, xu] (n,p) -> C (n,p,p), c (n,p), cc (n). */
int donk_quadratic_cost_l2_f64(int n, int p, const double* t, const double* w, double* C, double* c, double* cc,
                               void* stream);
int donk_quadratic_cost_l1_f64(int n, int p, const double* xu, const double* t, const double* w, double alpha,
                               double* C, double* c, double* cc, void* stream);

/* LinearDynamics.log_prob, batched: donk/dynamics/linear_dynamics.py:54-64.
 *   X (B,N,T+1,dX), U (B,N,T,dU), F (B,T,dX,p), f (B,T,dX), covar_inv (B,T,dX,dX), half_logdet (B,T) (both from
 *   donk_spd_factor_f64 on covar) -> log_prob (B,N,T). */
int donk_dynamics_log_prob_f64(int B, int N, int T, int dX, int dU, const double* X, const double* U, const double* F,
                               const double* f, const double* covar_inv, const double* half_logdet, double* log_prob,
                               void* stream);

/* ---- dynamics fit ---------------------------------------------------------------------- */

/* fit_lr batched over B conditions: donk/dynamics/linear_dynamics.py:137-189, with the GMM /
 * normal-inverse-Wishart prior of donk/dynamics/prior/prior_gmm.py:26-41 and prior.py:18-45 fused in.
 *   X (B,N,T+1,dX), U (B,N,T,dU) -> F (B,T,dX,p), f (B,T,dX), covar (B,T,dX,dX).
 *   Prior (gmm_K = 0: none): weights (K), means (K,d), covariances (K,d,d), prec_chol (K,d,d)
 *   with d = 2dX+dU, n_pool = GMMPrior.N, prior_weight as in fit_lr.
 *   info (B,T) int32: 0 ok, 1 = Sigma[:p,:p] not positive definite after regularisation.
 *   N <= 1 returns DONK_BAD_SHAPE (the reference raises ValueError, linear_dynamics.py:154). */
int donk_fit_lr_f64(int B, int N, int T, int dX, int dU, const double* X, const double* U,
                    int gmm_K, const double* gmm_weights, const double* gmm_means, const double* gmm_covariances,
                    const double* gmm_prec_chol, double n_pool, double prior_weight, double regularization,
                    double* F, double* f, double* covar, int* info, void* stream);

/* fit_lr in single precision (north-star "stated fp32 tolerance"): the same kernels instantiated for float storage and
 * arithmetic (the tensor-path products accumulate in FP64).  Tolerance against the fp64 path for well-conditioned
 * data (N > 2(dX+dU), regularisation 1e-6): F within 2e-3, f within 2e-3 of max|f|, covar within 5e-3, max-norm
 * relative (tests/test_fit_gmm_kernels.py::test_fit_lr_f32_tolerance).  Same arguments as donk_fit_lr_f64. */
int donk_fit_lr_f32(int B, int N, int T, int dX, int dU, const float* X, const float* U, int gmm_K,
                    const float* gmm_weights, const float* gmm_means, const float* gmm_covariances,
                    const float* gmm_prec_chol, float n_pool, float prior_weight, float regularization,
                    float* F, float* f, float* covar, int* info, void* stream);

/* fit_lr under an explicit normal-inverse-Wishart prior per (condition, t): what linear_dynamics.py:168-175 does
 * with ANY DynamicsPrior (donk/dynamics/prior/prior.py:62-82) after the caller has run prior.eval(xux_t):
 *   sigma_t = NIW(mu0, Phi, N_mean, N_covar).posterior(mean_t, cov_t, N / prior_weight).map_covar()  (prior.py:18-45).
 *   niw_mu0 (B,T,d), niw_Phi (B,T,d,d), niw_N_mean (B,T), niw_N_covar (B,T); everything else as donk_fit_lr_f64. */
int donk_fit_lr_niw_f64(int B, int N, int T, int dX, int dU, const double* X, const double* U, const double* niw_mu0,
                        const double* niw_Phi, const double* niw_N_mean, const double* niw_N_covar,
                        double prior_weight, double regularization, double* F, double* f, double* covar, int* info,
                        void* stream);

/* to_transitions: donk/dynamics/utils.py:4-17.  X (N,T+1,dX), U (N,T,dU) -> XUX (N*T, 2dX+dU). */
int donk_to_transitions_f64(int N, int T, int dX, int dU, const double* X, const double* U, double* XUX,
                            void* stream);

/* StateDistribution.fit batched: donk/samples/state_dist.py:27-37 (unbiased covariance).
 *   X0 (B,N,dX) with row stride `ld` doubles between samples -> mean (B,dX), covar (B,dX,dX). */
int donk_state_fit_f64(int B, int N, int dX, const double* X0, size_t sample_stride, size_t cond_stride,
                       double* mean, double* covar, void* stream);

/* ---- GMM prior: EM ----------------------------------------------------------------------- */

/* Number of doubles in the packed sufficient statistics of donk_gmm_em_step_f64:
 *   [ nk (K) | sum r (x-mu_k) (K,d) | sum r (x-mu_k)(x-mu_k)^T (K,d,d) | sum lse (1) ]. */
size_t donk_gmm_stats_size(int d, int K);
/* Workspace bytes donk_gmm_em_step_f64 needs for n_local rows. */
size_t donk_gmm_workspace_bytes(int n_local, int d, int K);

/* One E-step + sufficient-statistics pass over this rank's rows of the transition pool:
 * sklearn.mixture.GaussianMixture E-step and the sums of the M-step, as called from
 * GMMPrior.update (donk/dynamics/prior/prior_gmm.py:14-24).  X (n_local,d).
 * stats receives this rank's partial sums (the caller all-reduces them across ranks).
 * resp_in (n_local,K), optional: use these responsibilities instead of running the E-step (the
 * one-hot responsibilities of a hard initial labelling, sklearn/mixture/_base.py:119-128); the
 * lse entry of stats is then 0 and weights / prec_chol are not read.
 * shift (K,d), output: the point every cluster's statistics are centred on (first- and second-order sums are
 * taken of x - shift_k to avoid cancellation); pass it on to donk_gmm_m_finalize_f64. */
int donk_gmm_em_step_f64(int n_local, int d, int K, const double* X,
                         const double* weights, const double* means, const double* prec_chol,
                         const double* resp_in,
                         double* stats, double* shift, void* workspace, size_t workspace_bytes, void* stream);
/* Same, with the kernel family chosen by the caller: mode 0 = by pool size (scalar kernels below 65 536 rows), 1 = scalar
 * kernels (statistics centred on every cluster's own mean: agreement with scikit-learn at rounding level whatever the
 * spread of the clusters), 2 = DMMA kernels (centred on the mixture mean: fastest; error ~ eps (|mu_k - x_bar| / sigma)^2),
 * 3 = M pass only: responsibilities and per-block log-sum-exp totals were put into the workspace by
 * donk_gmm_tc_estep_f32 (kernel family of the M pass as in mode 0). */
int donk_gmm_em_step_mode_f64(int n_local, int d, int K, const double* X,
                              const double* weights, const double* means, const double* prec_chol,
                              const double* resp_in,
                              double* stats, double* shift, void* workspace, size_t workspace_bytes, int mode,
                              void* stream);

/* ---- E pass of GMM EM at fp32 tolerance on the tcgen05 tensor cores (abi_gmm_tc.cu) ------------------------------
 * GaussianMixture._e_step (sklearn/mixture/_base.py:552-582) behind GMMPrior.update, prior_gmm.py:14-24, for callers that
 * accept the north star's "stated fp32 tolerance": y = (x - c)(P_k) as one GEMM on tcgen05.mma kind::tf32 (operands split
 * into two TF32 planes, three products, fp32 accumulators in tensor memory), TMA tile loads, log-probabilities reduced
 * out of TMEM.  Tolerance against the fp64 E pass: responsibilities within 2e-4 absolute, the mean log-likelihood within
 * 1e-5 relative (tests/test_fit_gmm_kernels.py::test_gmm_tc_estep_tolerance).  d <= 80 (operands resident in shared
 * memory); other shapes keep the fp64 kernels.
 *   donk_gmm_tc_supported(d, K)          1 when the kernel handles the shape
 *   donk_gmm_tc_planes_bytes(n, d)       bytes of the split copy of the pool (two fp32 planes, row pitch round_up(d, 8))
 *   donk_gmm_tc_workspace_bytes(n, d, K) bytes of per-iteration scratch
 *   donk_gmm_tc_prepare_f32              X (n,d) fp64 minus center (d, may be NULL) -> planes; once per fit
 *   donk_gmm_tc_estep_f32                planes + current parameters -> resp (n,K) fp64 and lse_part (ceil(n/64)) fp64,
 *                                        i.e. the first two regions of the EM workspace; continue with
 *                                        donk_gmm_em_step_mode_f64(mode = 3), which then runs only the M pass. */
int donk_gmm_tc_supported(int d, int K);
size_t donk_gmm_tc_planes_bytes(int n, int d);
size_t donk_gmm_tc_workspace_bytes(int n, int d, int K);
int donk_gmm_tc_prepare_f32(int n, int d, const double* X, const double* center, void* planes, size_t planes_bytes,
                            void* stream);
int donk_gmm_tc_estep_f32(int n, int d, int K, const void* planes, const double* center, const double* weights,
                          const double* means, const double* prec_chol, double* resp, double* lse_part, void* work,
                          size_t work_bytes, void* stream);

/* ---- M pass of GMM EM at fp32 tolerance on the tcgen05 tensor cores (abi_gmm_tc.cu) ------------------------------
 * GaussianMixture._m_step (sklearn/mixture/_gaussian_mixture.py: _estimate_gaussian_parameters) behind GMMPrior.update,
 * prior_gmm.py:14-24: the weighted scatter S_k = sum_n r_nk (x_n - c)(x_n - c)^T and first moments of all clusters as one
 * GEMM over the sample index on tcgen05.mma kind::tf32 (3xTF32 split products, fp32 accumulators in tensor memory drained
 * into fp64 partial sums every 64 row tiles, TMA loads of a transposed split copy of the pool, the r_nk x_n operand built in
 * shared memory by producer warps).  Stated tolerance of a full tf32x3 fit against the fp64 fit: means within 1e-4 of the
 * data scale, covariances within 1e-3 relative, lower bound within 1e-5 relative
 * (tests/test_fit_gmm_kernels.py::test_gmm_tf32x3_fit_matches_fp64_fit).  Same shapes as the E pass (d <= 80).
 *   donk_gmm_tc_m_supported(d, K)        1 when the kernel handles the shape
 *   donk_gmm_tc_mplanes_bytes(n, d)      bytes of the transposed split copy (two fp32 planes round_up(d+1,16) x round_up(n,64))
 *   donk_gmm_tc_mwork_bytes(n, d, K)     bytes of per-iteration scratch
 *   donk_gmm_tc_mprepare_f32             X (n,d) fp64 minus center -> mplanes; once per fit
 *   donk_gmm_tc_mstep_f32                mplanes + resp (n,K) fp64 + lse_part (ceil(n/64)) (as donk_gmm_tc_estep_f32 left
 *                                        them in the EM workspace) -> stats (donk_gmm_stats_size) of this rank's rows and
 *                                        shift (K,d) = center; continue with the all-reduce and donk_gmm_m_finalize_f64. */
int donk_gmm_tc_m_supported(int d, int K);
size_t donk_gmm_tc_mplanes_bytes(int n, int d);
size_t donk_gmm_tc_mwork_bytes(int n, int d, int K);
int donk_gmm_tc_mprepare_f32(int n, int d, const double* X, const double* center, void* mplanes, size_t mplanes_bytes,
                             void* stream);
int donk_gmm_tc_mstep_f32(int n, int d, int K, const void* mplanes, const double* center, const double* resp,
                          const double* lse_part, double* stats, double* shift, void* work, size_t work_bytes,
                          void* stream);
/* E pass that hands over to the tensor-core M pass directly: as donk_gmm_tc_estep_f32, and in addition the transposed
 * single-precision responsibilities and their per-tile column sums are written into the M pass's scratch `mwork`
 * (donk_gmm_tc_mwork_bytes) in the layout donk_gmm_tc_mstep_f32 reads; call that with resp = NULL afterwards (it then
 * skips its own transpose of the fp64 responsibilities).  resp (n,K) is still written (callers may read it). */
int donk_gmm_tc_estep_m_f32(int n, int d, int K, const void* planes, const double* center, const double* weights,
                            const double* means, const double* prec_chol, double* resp, double* lse_part, void* work,
                            size_t work_bytes, void* mwork, size_t mwork_bytes, void* stream);

/* M-step finalisation from (all-reduced) statistics of n_total rows: new weights, means,
 * covariances (+ reg_covar I), precision Cholesky factors, mean log-likelihood.
 * shift (K,d): the centring point written by donk_gmm_em_step_f64 (NULL: the current means); means is output
 * (and the centring point when shift is NULL).
 * info (K) int32: 1 where a covariance is not positive definite (sklearn raises ValueError). */
int donk_gmm_m_finalize_f64(double n_total, int d, int K, const double* stats, const double* shift,
                            double reg_covar, double* weights, double* means, double* covariances,
                            double* prec_chol, double* lower_bound, int* info, void* stream);

/* Responsibilities (predict_proba) of rows X (n,d): resp (n,K).  Used by GMMPrior.eval for callers
 * that want the NIW object itself (prior_gmm.py:33). */
int donk_gmm_predict_proba_f64(int n, int d, int K, const double* X, const double* weights, const double* means,
                               const double* prec_chol, double* resp, void* stream);

/* Precision Cholesky factors from explicit precision matrices (sklearn precisions_init path). */
int donk_gmm_prec_chol_from_precisions_f64(int d, int K, const double* precisions, double* prec_chol,
                                           int* info, void* stream);

/* ---- k-means initialisation of the GMM prior ------------------------------------------------ */

/* The reference's first GMMPrior.update starts EM from k-means labels (sklearn/mixture/_base.py:119-128 via
 * donk/dynamics/prior/prior_gmm.py:21).  Lloyd passes on the device for pools that live there:
 * One pass over the rows X (n,d) with centres centers (K,d); every output is optional (NULL):
 *   labels (n) int32  nearest centre;   mind2 (n)  squared distance to it (k-means++ potential);
 *   dist (n,K)        squared distance to every centre (greedy k-means++ candidate scoring);
 *   sums (K*d + K)    [sum of the rows of each cluster (K,d) | row count (K)] of the labelling just computed, i.e.
 *                     the numerators of the next centres: one Lloyd iteration reads X once.  Needs `workspace`
 *                     of donk_kmeans_workspace_bytes(d, K) bytes.  Rows sharded over ranks: all-reduce `sums`. */
size_t donk_kmeans_workspace_bytes(int d, int K);
int donk_kmeans_pass_f64(int n, int d, int K, const double* X, const double* centers, int* labels, double* mind2,
                         double* dist, double* sums, void* workspace, size_t workspace_bytes, void* stream);
/* One round of the greedy k-means++ seeding scikit-learn's KMeans runs before Lloyd (sklearn/cluster/_kmeans.py:
 * _kmeans_plusplus, reached from GMMPrior.update through GaussianMixture(init_params="kmeans")): for Kc <= 16 candidate
 * centres cand (Kc,d) and the current nearest squared distances closest (n), in one read of the rows:
 *   dmin (n,Kc) = min(|x_n - cand_c|^2, closest_n)   and   pot (Kc) = sum_n dmin[n][c]  (the potential each candidate
 *   would leave; the winner is argmin pot and column `winner` of dmin is the next `closest`).
 * workspace: donk_kmeans_workspace_bytes(d, Kc) bytes. */
int donk_kmeanspp_round_f64(int n, int d, int Kc, const double* X, const double* cand, const double* closest, double* dmin,
                            double* pot, void* workspace, size_t workspace_bytes, void* stream);

/* ---- diagnostics ------------------------------------------------------------------------ */

/* FP64 pipe throughput probe (roofline denominator of the FP64-bound kernels): mode 0 = DFMA on the
 * CUDA cores, mode 1 = DMMA m8n8k4.  `flops` (host pointer) receives the operation count of the launch;
 * the caller times it with events on `stream`.  scratch: >= 1 double of device memory. */
int donk_fp64_probe(int mode, int blocks, int iters, double* scratch, double* flops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DONK_B200_H */
