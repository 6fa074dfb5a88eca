"""GPU diagnostic: per-phase CUDA-event timing of one SLM log-ML evaluation at
the config-2 shape (N, d, K from argv; defaults 1e6, 21, 2048)."""
import json
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.build()
import torch
from revrand_b200 import _engine as eng
from revrand_b200.basis_functions import RandomMatern32
from revrand_b200.slm import _SLMProblem
from bench import synthetic

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 21
K = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
X, y = synthetic(N, d)
basis = RandomMatern32(nbases=K, Xdim=d, random_state=1)
prob = _SLMProblem(basis, X, y)
plan, st = prob.plan, prob.stats


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), out


plan.set_lenscales([4.0])
res = {}
def f_suff():
    st.zero_(); eng.slm_suffstats(plan, prob.Xd, prob.yd, st, engine=prob.engine)
res["suffstats_ms"], _ = timed(f_suff)
lam = torch.ones(plan.D, dtype=torch.float64, device="cuda")
res["solve_value_only_ms"], _ = timed(lambda: eng.solve_posterior(st.G, st.p, 0.02, lam, need_C=False))
res["solve_ms"], post = timed(lambda: eng.solve_posterior(st.G, st.p, 0.02, lam))
res["c32_ms"], C32 = timed(lambda: post.C32())
m32 = post.m.float().contiguous()
def f_res():
    prob.rflat.zero_(); return eng.slm_residual(plan, prob.Xd, prob.yd, m32, sqerr=prob.sqerr)
res["residual_ms"], _ = timed(f_res)
def f_grad():
    prob.rflat.zero_(); eng.slm_gradpass(plan, prob.Xd, prob.yd, m32, C32, prob.R, prob.sqerr, engine=prob.engine)
res["gradpass_ms"], _ = timed(f_grad)
res["full_eval_ms"], _ = timed(lambda: prob.evaluate(0.02, [1.0], [4.0], want_grad=True))
res["shape"] = dict(N=N, d=d, K=K)
print(json.dumps(res))
