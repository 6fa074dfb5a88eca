#!/usr/bin/env python
"""Config 5 of BASELINE.json: BasisCat(RandomRBF(K) + LinearBasis(onescol=True))
hyper-parameter sweep, K in {512 .. 8192}, N = 1e7 rows sharded over the ranks
(8 x B200): log-ML evaluations (value + gradients) per second, roofline fraction
of the value pass and per-phase times, one JSON line per K.

  torchrun --nproc-per-node 8 bench.py --workload config5 --gpus 8 [--K 512,1024,...]

Also the single-process driver of ``scripts/eval_breakdown.py``-style phase
timings at any N (``--workload config5 --N 1000000 --K 2048`` on one GPU).
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def shard_synthetic(N, d, lo, hi, seed=0):
    """Rows [lo, hi) of the config-5 data set, generated shard by shard (the
    full 1e7 x 21 matrix is never needed on one rank): block b of 2^20 rows uses
    RandomState(seed + b)."""
    B = 1 << 20
    w = np.random.RandomState(seed).randn(d)
    Xs, ys = [], []
    for b in range(lo // B, (hi + B - 1) // B):
        rs = np.random.RandomState(seed + 1 + b)
        Xb = rs.randn(B, d).astype(np.float32)
        yb = (np.sin(Xb.astype(np.float64).dot(w) / 3.0) + 0.1 * rs.randn(B)).astype(np.float32)
        a, e = max(lo, b * B) - b * B, min(hi, (b + 1) * B) - b * B
        Xs.append(Xb[a:e])
        ys.append(yb[a:e])
    return np.concatenate(Xs), np.concatenate(ys)


class _Shard(object):
    """What _SLMProblem needs from X when the rank only holds its own rows."""

    def __init__(self, X, N_total, lo):
        self.X, self.shape, self.lo = X, (N_total, X.shape[1]), lo

    def __getitem__(self, sl):
        if isinstance(sl, slice):
            a = 0 if sl.start is None else sl.start
            b = self.shape[0] if sl.stop is None else sl.stop
            if a >= self.lo and b - self.lo <= self.X.shape[0]:
                return self.X[a - self.lo:b - self.lo]
            if b <= 1:                       # the one-row probe of regularizer_diagonal
                return self.X[:1]
        raise IndexError("row range outside this rank's shard")


def main(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    import bench
    from revrand_b200 import Parameter, Positive, _cabi, _engine
    from revrand_b200.basis_functions import LinearBasis, RandomRBF
    from revrand_b200.slm import _SLMProblem

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world,
                                device_id=torch.device("cuda", local))
    N, d = args.N, args.d
    Ks = [int(k) for k in str(args.Ks).split(",")]
    lo, hi = _engine.shard_rows(N, rank, world)
    Xl, yl = shard_synthetic(N, d, lo, hi)
    pk = bench.peaks()

    def tmax(ms):
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        return tmax(a.elapsed_time(b)), out

    for K in Ks:
        basis = RandomRBF(nbases=K, Xdim=d, random_state=1, lenscale=Parameter(4.0, Positive())) \
            + LinearBasis(onescol=True)
        prob = _SLMProblem(basis, _Shard(Xl, N, lo), _YShard(yl, N, lo))
        D = prob.D
        plan, st = prob.plan, prob.stats
        regs, hyp, var = [1.0, 1.0], [4.0], 0.05
        prob.evaluate(var, regs, hyp, want_grad=True)          # warm-up (allocations, plans)
        ph = {}

        kept = prob._kept_buffer()     # the fp16 feature image shared by the two passes

        def f_value():
            st.zero_()
            if kept is not None:
                _engine.slm_suffstats_keep(plan, prob.Xd, prob.yd, st, kept, want_yy=False)
            else:
                _engine.slm_suffstats(plan, prob.Xd, prob.yd, st, engine=prob.engine,
                                      want_yy=False)
        ph["value_pass_ms"], _ = timed(f_value)
        ph["allreduce_stats_ms"], _ = timed(lambda: _engine.allreduce_sum_(st.flat))
        lam = torch.ones(D, dtype=torch.float64, device="cuda")
        ph["solve_ms"], post = timed(lambda: _engine.solve_posterior(st.G, st.p, var, lam))
        m32 = post.m.float().contiguous()
        ph["c32_ms"], C32 = timed(lambda: post.C32())

        def f_grad():
            prob.rflat.zero_()
            if kept is not None:
                _engine.slm_gradpass_kept(plan, prob.Xd, prob.yd, m32, C32, prob.R, prob.sqerr,
                                          kept)
            else:
                _engine.slm_gradpass(plan, prob.Xd, prob.yd, m32, C32, prob.R, prob.sqerr,
                                     engine=prob.engine)
        ph["gradient_pass_ms"], _ = timed(f_grad)
        ph["allreduce_grad_ms"], _ = timed(lambda: _engine.allreduce_sum_(prob.rflat))
        del post, C32
        steps = max(1, args.steps if K <= 2048 else min(args.steps, 2))
        tot = 0.0
        for i in range(steps):
            ms, r = timed(lambda: prob.evaluate(var, regs, [4.0 + i], want_grad=True))
            tot += ms
        ms_eval = tot / steps
        n_local = hi - lo
        flops = 2.0 * n_local * D * D + 2.0 * n_local * d * K
        ach = flops / (ph["value_pass_ms"] * 1e-3) / 1e12
        line = {
            "metric": "log-ML evals/sec, BasisCat(RandomRBF(K)+LinearBasis) sweep",
            "value": 1e3 / ms_eval, "unit": "evals/s", "n_gpus": world, "steps": steps,
            "ms_per_step": ms_eval, "higher_is_better": True, "scaling": "strong",
            "dtype": "s8", "data": "synthetic",
            "config": {"workload": "config5: SLM + BasisCat(RandomRBF(nbases=%d) + LinearBasis"
                                   "(onescol)), N=%d, d=%d, D=%d, value+grad eval" % (K, N, d, D)},
            "phases_ms": ph,
            "roofline": {"bound": "tensor", "kernel": "value pass (t3_syrk_kernel + generator)",
                         "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s",
                         "frac": ach / pk["tflops"], "ms_per_launch": ph["value_pass_ms"],
                         "algorithmic_flops_per_launch": flops,
                         "peak_source": pk["src"] + " bf16 sustained"},
            "check": {"logdet": float(r["logdet"]), "m_norm": float(r["m"].norm().item())},
            "workspace_gb": float(_engine._workspace[local].numel()) / 2 ** 30
            if local in _engine._workspace else None,
            "kept_features_gb": None if kept is None else float(kept.numel()) / 2 ** 30,
        }
        if rank == 0:
            print(json.dumps(line), flush=True)
        del prob, plan, st, kept
        _engine._workspace.clear()
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


class _YShard(object):
    def __init__(self, y, N_total, lo):
        self.y, self.lo, self.shape = y, lo, (N_total,)

    def __getitem__(self, sl):
        a = 0 if sl.start is None else sl.start
        b = self.shape[0] if sl.stop is None else sl.stop
        return self.y[a - self.lo:b - self.lo]
