"""CPU, world_size 2 over gloo: the row-sharded reduction used when one
process drives each GPU.  Each rank forms the sufficient statistics of its
own contiguous row shard (here with the oracle, standing in for the CUDA
pass), the flat [G | p | yy] buffer is summed with ONE allreduce, and both
ranks must end with the statistics of the full data set."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as orc
        from revrand_b200 import _engine
        rs = np.random.RandomState(0)
        N, d, K = 501, 3, 9
        X, y = rs.randn(N, d), rs.randn(N)
        W = np.random.RandomState(1).randn(d, K)
        assert _engine.world() == (rank, world)
        lo, hi = _engine.shard_rows(N, rank, world)
        Phi = orc.trig_features(X[lo:hi], W, 1.3)
        D = 2 * K
        flat = torch.zeros(D * D + D + 1, dtype=torch.float64)
        flat[:D * D] = torch.from_numpy(Phi.T.dot(Phi).ravel())
        flat[D * D:D * D + D] = torch.from_numpy(Phi.T.dot(y[lo:hi]))
        flat[-1] = float(y[lo:hi].dot(y[lo:hi]))
        _engine.allreduce_sum_(flat)
        Pf = orc.trig_features(X, W, 1.3)
        ok = (np.allclose(flat[:D * D].numpy().reshape(D, D), Pf.T.dot(Pf))
              and np.allclose(flat[D * D:D * D + D].numpy(), Pf.T.dot(y))
              and np.isclose(flat[-1].item(), y.dot(y)))
        # the posterior solve with the inverse distributed over the ranks
        # must equal the one-process float64 inverse on every rank
        A = Pf.T.dot(Pf)
        lam = np.full(D, 1.7)
        post = _engine.solve_posterior(torch.from_numpy(A), torch.from_numpy(Pf.T.dot(y)),
                                       0.3, torch.from_numpy(lam))
        Cref = np.linalg.inv(np.diag(1.0 / lam) + A / 0.3)
        ok = ok and np.allclose(post.C.numpy(), Cref, rtol=1e-9, atol=1e-12)
        ok = ok and np.allclose(post.m.numpy(), Cref.dot(Pf.T.dot(y)) / 0.3)
        ok = ok and np.isclose(post.trgc.item(), np.sum(A * Cref))
        # same through the GEMM-rich blocked path used on the GPU for D >= 1024
        _engine._BLOCK_INV_MIN, _engine._BLOCK_INV_LEAF = 8, 4
        _engine._BLOCK_INV_CUDA_ONLY = False
        _engine._TRI_INV_SHARD_MIN = 8     # top levels of L^-1 split over the two ranks
        post2 = _engine.solve_posterior(torch.from_numpy(A), torch.from_numpy(Pf.T.dot(y)),
                                        0.3, torch.from_numpy(lam))
        ok = ok and np.allclose(post2.C.numpy(), Cref, rtol=1e-9, atol=1e-12)
        # ... and the float32 image of C the gradient pass reads: row slabs of
        # Li^T Li formed by the ranks (two slabs each, triangular part only), all-gathered
        ok = ok and post2._C32 is not None
        ok = ok and np.allclose(post2.C32().numpy(), Cref, rtol=1e-5, atol=1e-7 * np.abs(Cref).max())
        # ranks of a sharded fit share rank 0's random starts, and a basis that
        # differs between ranks is refused
        mine = np.random.RandomState(100 + rank)
        first_of_rank0 = np.random.RandomState(100).randn(3)
        _engine.sync_random_state(mine)
        ok = ok and np.array_equal(mine.randn(3), first_of_rank0)
        _engine.assert_same_on_all_ranks(np.abs(W).sum(), "W")
        try:
            _engine.assert_same_on_all_ranks(np.abs(W).sum() + rank, "W")
            ok = False
        except Exception as e:
            ok = ok and "differs between ranks" in str(e)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_row_sharded_allreduce_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
