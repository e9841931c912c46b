"""N>1 host logic on CPU: world_size-2 gloo processes over the emulation backend.

EM rows are sharded over ranks and the packed statistics all-reduced (the one exchange step of the path);
ILQR conditions are sharded with no collective.  Results must match the single-rank run.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from donk_testutil import ROOT, relerr


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), DONK_GMM_MMA="1")  # the sharded DMMA path
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from donk_b200 import _lib, synthetic
    from donk_b200.batched import DeviceGMM
    from emu.build_emu import backend

    _lib.install_test_backend(backend())
    X, w0, m0, p0 = synthetic.gmm_pool(1200, 8, 3, seed=11)
    rows = np.array_split(np.arange(1200), world)[rank]
    gm = DeviceGMM(3, tol=0.0, max_iter=4).set_params(w0, m0, precisions=p0)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gm.fit(X[rows])
    assert gm.N == 1200
    # sharded k-means (centres broadcast from rank 0, per-cluster sums all-reduced) and the GMMPrior front end on it
    from donk_b200.batched import DeviceKMeans
    from donk_b200.dynamics.prior import GMMPrior

    Xk, _, _, _ = synthetic.gmm_pool(900, 6, 3, seed=5)
    rows_k = np.array_split(np.arange(900), world)[rank]
    km = DeviceKMeans(3, random_state=0).fit(Xk[rows_k])
    prior = GMMPrior(3, random_state=0, init="device")
    prior.update(Xk[rows_k])
    assert prior.N == 900
    # conditions sharded over ranks, no collective: each rank steps its own slice
    from donk_b200.batched import BatchedILQR

    sys.path.insert(0, str(ROOT))
    import oracle

    conds = [synthetic.condition(23, i, 4, 6, 2, 12) for i in range(4)]
    mine = [c for i, c in enumerate(conds) if i % world == rank]
    fits = [oracle.fit_lr(c.X, c.U) for c in mine]
    x0 = [oracle.state_fit(c.X[:, 0]) for c in mine]
    st = lambda xs: np.stack(xs)  # noqa: E731
    il = BatchedILQR(st([f[0] for f in fits]), st([f[1] for f in fits]), st([f[2] for f in fits]),
                     st([c.prev_K for c in mine]), st([c.prev_k for c in mine]), st([c.C for c in mine]),
                     st([c.c for c in mine]), st([c.cc for c in mine]), st([m for m, _ in x0]),
                     st([S for _, S in x0]), prev_covar=st([c.prev_covar for c in mine]))
    kl = il.step(torch.ones(len(mine), 1, dtype=torch.float64)).kl_div[:, 0].numpy()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), means=gm.means.numpy(), cov=gm.covariances.numpy(),
             w=gm.weights.numpy(), lb=gm.lower_bound, centers=km.centers.numpy(), km_iter=km.n_iter,
             labels=km.labels.numpy(), pmeans=prior.gmm.means_, plb=prior.gmm.lower_bound_, kl=kl)
    dist.barrier()
    dist.destroy_process_group()


def test_em_sharded_rows_match_single_rank(tmp_path, backend, monkeypatch):
    if backend.device_type != "cpu":
        pytest.skip("multi-rank logic is covered on CPU with gloo; the GPU run uses torchrun + NCCL (bench.py)")
    import warnings

    from donk_b200 import synthetic
    from donk_b200.batched import DeviceGMM

    monkeypatch.setenv("DONK_GMM_MMA", "1")  # the sharded DMMA path, here on a small pool
    X, w0, m0, p0 = synthetic.gmm_pool(1200, 8, 3, seed=11)
    ref = DeviceGMM(3, tol=0.0, max_iter=4).set_params(w0, m0, precisions=p0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref.fit(X)
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        z = np.load(tmp_path / f"r{r}.npz")
        assert relerr(z["means"], ref.means.numpy()) < 1e-12
        assert relerr(z["cov"], ref.covariances.numpy()) < 1e-11
        assert relerr(z["w"], ref.weights.numpy()) < 1e-12
        assert abs(float(z["lb"]) - ref.lower_bound) < 1e-12 * abs(ref.lower_bound)
    a, b = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    assert np.array_equal(a["cov"], b["cov"])  # parameters stay replicated bit for bit
    # k-means: both ranks end with the same centres; every local row sits with its nearest centre
    assert np.array_equal(a["centers"], b["centers"]) and int(a["km_iter"]) == int(b["km_iter"])
    Xk, _, _, _ = synthetic.gmm_pool(900, 6, 3, seed=5)
    for r, z in enumerate((a, b)):
        rows_k = np.array_split(np.arange(900), 2)[r]
        dist2 = ((Xk[rows_k][:, None, :] - z["centers"][None]) ** 2).sum(-1)
        assert np.array_equal(z["labels"], dist2.argmin(1))
    assert np.array_equal(a["pmeans"], b["pmeans"]) and float(a["plb"]) == float(b["plb"])
    # ILQR: the union of the ranks' slices equals the single-rank batch
    import oracle
    from oracle.lqg import ILQRProblem

    conds = [synthetic.condition(23, i, 4, 6, 2, 12) for i in range(4)]
    for i, c in enumerate(conds):
        prob = ILQRProblem(*oracle.fit_lr(c.X, c.U), c.C, c.c, c.cc, c.prev_K, c.prev_k, c.prev_covar,
                           np.linalg.inv(c.prev_covar), *oracle.state_fit(c.X[:, 0]))
        got = (a, b)[i % 2]["kl"][i // 2]
        assert abs(float(got) - oracle.ilqr_step(prob, 1.0).kl_div) < 1e-9
