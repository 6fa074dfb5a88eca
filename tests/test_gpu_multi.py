"""GPU, two ranks over NCCL (skipped on a one-GPU box): the row-sharded
evaluation must reproduce the one-process evaluation -- the value pass forms
its statistics from 24-bit integers exactly, so sharding only changes the order
of a few float64 roundings."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

N, D_IN, K = 150001, 21, 192
POINT = dict(var=0.05, reg=1.3, ls=3.0)


def _inputs():
    rs = np.random.RandomState(3)
    X = rs.randn(N, D_IN).astype(np.float32)
    y = (np.sin(X.astype(np.float64).dot(rs.randn(D_IN)) / 3.0) + 0.1 * rs.randn(N)).astype(np.float32)
    return X, y


def _evaluate():
    from revrand_b200 import config
    from revrand_b200.basis_functions import RandomMatern32
    from revrand_b200.slm import _SLMProblem
    config.ENGINE = "tcgen05"
    X, y = _inputs()
    basis = RandomMatern32(nbases=K, Xdim=D_IN, random_state=1)
    prob = _SLMProblem(basis, X, y)
    r = prob.evaluate(POINT["var"], [POINT["reg"]], [POINT["ls"]], want_grad=True)
    return dict(logdet=float(r["logdet"]), trgc=float(r["trgc"]), sqerr=float(r["sqerr"]),
                q=float(r["q"][0]), m=r["m"].cpu().numpy(), g=np.asarray(r["g"][0]),
                diagC=r["post"].diagC.cpu().numpy())


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        ret[rank] = _evaluate()
    finally:
        dist.destroy_process_group()


def test_two_rank_evaluation_equals_one_rank():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    one = _evaluate()
    port = 29600 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    for rank in (0, 1):
        two = ret[rank]
        # G-derived quantities: exact integer statistics, float64 solve
        for k in ("logdet", "trgc"):
            assert abs(two[k] - one[k]) <= 1e-9 * abs(one[k]), (rank, k, two[k], one[k])
        np.testing.assert_allclose(two["diagC"], one["diagC"], rtol=1e-9)
        # Phi^T y: every shard quantises y against its own max |y| (2^-23 per row)
        assert np.linalg.norm(two["m"] - one["m"]) <= 1e-7 * np.linalg.norm(one["m"])
        assert abs(two["q"] - one["q"]) <= 1e-7 * abs(one["q"])
        # residuals: fp32 fitted values summed with float atomics
        assert abs(two["sqerr"] - one["sqerr"]) <= 1e-6 * abs(one["sqerr"])
        # the gradient pass sums fp32 tile results in a different order per shard
        assert np.linalg.norm(two["g"] - one["g"]) <= 1e-5 * np.linalg.norm(one["g"])
    # both ranks hold the same numbers
    assert ret[0]["logdet"] == ret[1]["logdet"]
    np.testing.assert_array_equal(ret[0]["m"], ret[1]["m"])
