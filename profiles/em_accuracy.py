import os, sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
os.environ["DONK_GMM_MMA"] = sys.argv[1]
from donk_testutil import load_golden, relerr
from donk_b200.dynamics.prior import GMMPrior
from donk_b200.dynamics import to_transitions
g = load_golden("fitlr_traj00")
X, U = g["X"], g["U"]
prior = GMMPrior(8, random_state=0)
prior.update(to_transitions(X, U).reshape(-1, 32))
print("DONK_GMM_MMA", sys.argv[1], "means", relerr(prior.gmm.means_, g["gmm_means"]), "cov", relerr(prior.gmm.covariances_, g["gmm_covariances"]), "w", relerr(prior.gmm.weights_, g["gmm_weights"]))
