#!/usr/bin/env python
"""Whole ``StandardLinearModel.fit`` calls (random starts + L-BFGS-B), the caller of
the hot path: evaluations per second as the optimiser actually issues them.

  python bench.py --workload fit

* config 1 of BASELINE.json: RandomRBF(nbases=256) on a 1-D sine, N=1000,
  nstarts=100 (the reference's regression-demo plumbing; SIMT engine, launch
  latency bound);
* a mid-size fit: RandomMatern32(nbases=512), N=40000, d=21, ARD lengthscales,
  nstarts=20, at most 40 L-BFGS-B iterations (tensor-core engine).
One JSON line.
"""

from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def timed_fit(slm, X, y):
    import torch
    calls = {"n": 0, "grad": 0}
    orig = slm._elbo

    def counted(Xa, ya, var, reg, hyp, want_grad=True):
        calls["n"] += 1
        calls["grad"] += int(bool(want_grad))
        return orig(Xa, ya, var, reg, hyp, want_grad=want_grad)
    slm._elbo = counted
    orig_batch = slm._elbo_values

    def counted_batch(Xa, ya, points):
        calls["n"] += len(points)
        before = calls["n"]
        out = orig_batch(Xa, ya, points)
        calls["n"] = before          # (points re-run one by one were already counted)
        return out
    slm._elbo_values = counted_batch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    slm.fit(X, y)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return dt, calls


def main(args):
    import __graft_entry__ as g
    g.build()
    import torch
    import revrand_b200 as rr
    from revrand_b200 import Parameter, Positive
    from revrand_b200 import basis_functions as bf
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    out = {}
    rs = np.random.RandomState(0)
    X = np.sort(rs.uniform(-5, 5, size=(1000, 1)), axis=0)
    y = np.sin(X[:, 0]) + 0.1 * rs.randn(1000)
    for rep in range(2):      # first fit pays the allocations
        slm = rr.StandardLinearModel(basis=bf.RandomRBF(nbases=256, Xdim=1, random_state=1),
                                     nstarts=100, maxiter=200, random_state=2)
        dt, calls = timed_fit(slm, X, y)
    out["config1"] = {"seconds": dt, "evaluations": calls["n"], "with_gradients": calls["grad"],
                      "fit_evals_per_s": calls["n"] / dt, "elbo": float(slm.obj_),
                      "message": str(slm.opt_message_),
                      "reference": "unmodified reference, same seeds: 128 evaluations in 22.3 s "
                                   "on this container's CPU, ELBO 884.1130586"}
    N, d, K = 40000, 21, 512
    X = rs.randn(N, d)
    y = np.sin(X.dot(rs.randn(d)) / 3.0) + 0.1 * rs.randn(N)
    for rep in range(2):
        slm = rr.StandardLinearModel(
            basis=bf.RandomMatern32(nbases=K, Xdim=d, random_state=1,
                                    lenscale=Parameter(3.0 * np.ones(d), Positive())),
            nstarts=20, maxiter=40, random_state=2)
        dt, calls = timed_fit(slm, X, y)
    out["midsize"] = {"N": N, "d": d, "nbases": K, "seconds": dt, "evaluations": calls["n"],
                      "with_gradients": calls["grad"], "fit_evals_per_s": calls["n"] / dt,
                      "elbo": float(slm.obj_), "message": str(slm.opt_message_)}
    # config 2 of BASELINE.json as a (short) fit: 16 random starts + at most 6 L-BFGS-B
    # iterations at N=1e6, d=21, K=2048 -- with and without pipelined random starts
    from bench import synthetic
    from revrand_b200 import config
    N, d, K = 1000000, 21, 2048
    X, y = synthetic(N, d)
    for tag, pipe in (("config2_pipelined", True), ("config2_sequential", False)):
        config.PIPELINE_STARTS = pipe
        for rep in range(2):
            slm = rr.StandardLinearModel(
                basis=bf.RandomMatern32(nbases=K, Xdim=d, random_state=1),
                nstarts=16, maxiter=6, random_state=2)
            dt, calls = timed_fit(slm, X, y)
        out[tag] = {"N": N, "d": d, "nbases": K, "seconds": dt, "evaluations": calls["n"],
                    "with_gradients": calls["grad"], "fit_evals_per_s": calls["n"] / dt,
                    "elbo": float(slm.obj_), "message": str(slm.opt_message_)}
    config.PIPELINE_STARTS = True
    print(json.dumps({"metric": "StandardLinearModel.fit evaluations/sec", "unit": "evals/s",
                      "value": out["config2_pipelined"]["fit_evals_per_s"], "fits": out}),
          flush=True)
