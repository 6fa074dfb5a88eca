#!/usr/bin/env python
"""Per-phase timings of one config-2 evaluation with the feature image kept between
the two passes against regenerating it (CUDA events, median of --reps)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=1000000)
    ap.add_argument("--K", type=int, default=2048)
    ap.add_argument("--d", type=int, default=21)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--once", action="store_true", help="one evaluation per mode (for ncu)")
    a = ap.parse_args()
    import torch
    import bench
    from revrand_b200 import _cabi, _engine, config
    from revrand_b200.basis_functions import RandomMatern32
    from revrand_b200.slm import _SLMProblem
    X, y = bench.synthetic(a.N, a.d) if hasattr(bench, "synthetic") else (None, None)
    if X is None:
        rs = np.random.RandomState(0)
        X = rs.randn(a.N, a.d).astype(np.float32)
        y = (np.sin(X.astype(np.float64).dot(rs.randn(a.d)) / 3.0) + 0.1 * rs.randn(a.N)).astype(np.float32)
    basis = RandomMatern32(nbases=a.K, Xdim=a.d, random_state=1)
    prob = _SLMProblem(basis, X, y)
    plan, st = prob.plan, prob.stats
    var, regs, hyp = 0.02, [1.0], [4.0]
    kept = prob._kept_buffer()
    assert kept is not None
    r = prob.evaluate(var, regs, hyp)            # warm-up
    if a.once:
        config.KEEP_FEATURES_MAX_BYTES = 0
        prob._kept, prob._kept_tried = None, True
        prob.evaluate(var, regs, hyp)
        return
    lam = torch.ones(prob.D, dtype=torch.float64, device="cuda")
    post = _engine.solve_posterior(st.G.clone(), st.p.clone(), var, lam)
    m32, C32 = post.m.float().contiguous(), post.C32()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn):
        ts = []
        for _ in range(a.reps):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    def v_plain():
        st.zero_()
        _engine.slm_suffstats(plan, prob.Xd, prob.yd, st, engine=prob.engine, want_yy=False)

    def v_keep():
        st.zero_()
        _engine.slm_suffstats_keep(plan, prob.Xd, prob.yd, st, kept, want_yy=False)

    def g_regen():
        prob.rflat.zero_()
        _engine.slm_gradpass(plan, prob.Xd, prob.yd, m32, C32, prob.R, prob.sqerr, engine=prob.engine)

    def g_kept():
        prob.rflat.zero_()
        _engine.slm_gradpass_kept(plan, prob.Xd, prob.yd, m32, C32, prob.R, prob.sqerr, kept)

    out = {"value_pass_ms": timed(v_plain), "value_pass_keep_ms": timed(v_keep),
           "gradient_pass_regen_ms": timed(g_regen), "gradient_pass_kept_ms": timed(g_kept),
           "solve_ms": timed(lambda: _engine.solve_posterior(st.G, st.p, var, lam).C32())}
    # whole evaluations, the two modes interleaved (same thermal / power state)
    ts = {True: [], False: []}
    prob._kept_tried = True
    for i in range(4 * a.reps):
        keep = (i % 2 == 0)
        prob._kept = kept if keep else None
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        prob.evaluate(var, regs, hyp)
        e1.record()
        torch.cuda.synchronize()
        ts[keep].append(e0.elapsed_time(e1))
    out["evaluate_keep_ms"] = float(np.mean(ts[True][1:]))
    out["evaluate_regen_ms"] = float(np.mean(ts[False][1:]))
    out["evaluate_keep_all"] = [round(v, 2) for v in ts[True]]
    out["evaluate_regen_all"] = [round(v, 2) for v in ts[False]]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
