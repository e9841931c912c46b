"""donk_b200 — B200-native trajectory-optimisation hot path of DiddiZ/donk.ai.

Drop-in names (same signatures as the reference):
    donk_b200.dynamics.linear_dynamics.fit_lr, donk_b200.dynamics.prior.GMMPrior,
    donk_b200.traj_opt.ILQR (step / sample_surface / optimize), donk_b200.traj_opt.lqg.{backward, forward,
    extended_costs_kl, kl_divergence_action, step_adjust}
Batched device API: donk_b200.batched.  All math runs in hand-written sm_100a CUDA kernels
(donk_b200/csrc) behind the C ABI of include/donk_b200.h; there is no CPU fallback.
"""
__version__ = "0.1"
