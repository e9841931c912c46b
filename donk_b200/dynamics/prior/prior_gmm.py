"""GMM dynamics prior (donk/dynamics/prior/prior_gmm.py:7-41) with the EM on the B200.

The reference wraps ``sklearn.mixture.GaussianMixture``; here the mixture lives on the device
(``donk_b200.batched.DeviceGMM``) and ``.gmm`` is a read-only view with sklearn's attribute names.
"""
from __future__ import annotations

import numpy as np

from donk_b200.dynamics.prior.prior import DynamicsPrior, NormalInverseWishart


class ConvergenceWarning(UserWarning):
    """EM stopped at ``max_iter`` (same role as ``sklearn.exceptions.ConvergenceWarning``)."""


class _MixtureView:
    """sklearn-style attribute names over the device mixture (numpy copies on access)."""

    def __init__(self, gmm, owner):
        self._g = gmm
        self._owner = owner
        self.warm_start = False

    n_components = property(lambda s: s._g.K)
    weights_ = property(lambda s: s._g.weights.cpu().numpy())
    means_ = property(lambda s: s._g.means.cpu().numpy())
    covariances_ = property(lambda s: s._g.covariances.cpu().numpy())
    precisions_cholesky_ = property(lambda s: s._g.precisions_cholesky.cpu().numpy())
    lower_bound_ = property(lambda s: s._g.lower_bound)
    lower_bounds_ = property(lambda s: list(s._g.lower_bounds))
    n_iter_ = property(lambda s: s._g.n_iter)
    converged_ = property(lambda s: s._g.converged)

    def predict_proba(self, X):
        return self._g.predict_proba(np.asarray(X, dtype=np.float64)).cpu().numpy()


class GMMPrior(DynamicsPrior):
    """A Gaussian-mixture based prior: ``GMMPrior(n_clusters, random_state=None)``."""

    def __init__(self, n_clusters: int, random_state=None, init: str = "auto") -> None:
        """``init``: where the k-means labels of the first ``update`` come from — ``"device"`` / ``"auto"`` (the default:
        ``batched.DeviceKMeans``, k-means++ and Lloyd on the GPU; labels agree with scikit-learn statistically, its
        random stream cannot be reproduced) or ``"sklearn"`` (opt-in, single rank only: scikit-learn's own ``KMeans`` on the
        host, the reference's exact labels — a test / migration aid, not part of the device path)."""
        from donk_b200.batched import DeviceGMM

        import os

        self.random_state = random_state
        # DONK_GMM_INIT=sklearn: process-wide opt-in used when the reference's own known-answer tests run against
        # this package (they pin scikit-learn's random stream through the k-means labels)
        self.init = os.environ.get("DONK_GMM_INIT", init) if init == "auto" else init
        self._gmm = DeviceGMM(n_clusters)
        self.gmm = _MixtureView(self._gmm, self)

    def update(self, XUX, init=None, group=None) -> None:
        """Refit the mixture to the transition pool ``XUX (n, d)`` (prior_gmm.py:14-24).

        First call: initial responsibilities from k-means labels, like sklearn's default
        ``init_params="kmeans"`` with ``n_init=1`` and this prior's ``random_state``
        (sklearn/mixture/_base.py:119-128); later calls warm-start from the current parameters
        (prior_gmm.py:24).  ``init=(weights, means, precisions)`` gives an explicit, RNG-free start;
        ``group`` is the ``torch.distributed`` group over which the rows of ``XUX`` are sharded.
        """
        g = self._gmm
        warm = self.gmm.warm_start and g.means is not None
        if isinstance(XUX, np.ndarray):
            XUX = np.asarray(XUX, dtype=np.float64)
        if not warm:
            if init is not None:
                w, m, prec = init
                g.set_params(w, m, precisions=prec)
            else:
                import torch.distributed as dist

                sharded = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
                if sharded and self.init == "sklearn":
                    # every rank would label ITS rows with its own k-means: cluster k would mean different things on
                    # different ranks and the all-reduced statistics would mix unrelated clusters
                    raise ValueError("init='sklearn' labels a single-rank pool; with rows sharded over ranks use "
                                     "init='device' (centres are seeded on rank 0 and broadcast)")
                use_device = self.init in ("device", "auto") or sharded
                if use_device:
                    from donk_b200.batched import DeviceKMeans

                    km = DeviceKMeans(g.K, self.random_state).fit(XUX, group)
                    g.init_from_labels(XUX, km.labels, group)
                else:
                    g.init_from_labels(XUX, self._kmeans_labels(XUX, g.K), group)
        g.fit(XUX, warm_start=warm, group=group)
        self.N = g.N
        self.gmm.warm_start = True

    def _kmeans_labels(self, XUX, K):
        # Host-side initial labelling with scikit-learn's own KMeans: reproduces the reference's labels exactly
        # (its random stream).  Large / device-resident pools use batched.DeviceKMeans instead (init="device").
        from sklearn.cluster import KMeans
        from sklearn.utils import check_random_state

        Xh = XUX.cpu().numpy() if hasattr(XUX, "cpu") else np.asarray(XUX)
        if Xh.shape[0] < K:
            raise ValueError(f"Expected n_samples >= n_components but got n_components = {K}, "
                             f"n_samples = {Xh.shape[0]}")
        return KMeans(n_clusters=K, n_init=1, random_state=check_random_state(self.random_state)).fit(Xh).labels_

    def eval(self, XUX) -> NormalInverseWishart:
        """Mixture-weighted NIW prior for the transitions ``XUX (N, d)`` (prior_gmm.py:26-41), evaluated on the device.

        A non-informative NIW updated with the responsibility-weighted cluster mean / covariance at the prior's
        strength ``N_p`` is ``NIW(mu_p, N_p Phi_p, N_p, N_p - (1 + d))``: built directly.
        """
        mu0, Phi, n_mean, n_covar = self._gmm.eval_prior(np.asarray(XUX, dtype=np.float64))
        return NormalInverseWishart(mu0.cpu().numpy(), Phi.cpu().numpy(), n_mean, n_covar)
