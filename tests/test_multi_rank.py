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
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
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
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), means=gm.means.numpy(), cov=gm.covariances.numpy(),
             w=gm.weights.numpy(), lb=gm.lower_bound)
    dist.barrier()
    dist.destroy_process_group()


def test_em_sharded_rows_match_single_rank(tmp_path, backend):
    if backend.device_type != "cpu":
        pytest.skip("multi-rank logic is covered on CPU with gloo; the GPU run uses torchrun + NCCL (bench.py)")
    import warnings

    from donk_b200 import synthetic
    from donk_b200.batched import DeviceGMM

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
