"""Time only the gradient pass at the config-2 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.build()
import numpy as np, torch
from revrand_b200 import _engine as eng
from revrand_b200.basis_functions import RandomMatern32
from revrand_b200.slm import _SLMProblem
from bench import synthetic
X, y = synthetic(1000000, 21)
prob = _SLMProblem(RandomMatern32(nbases=2048, Xdim=21, random_state=1), X, y)
plan = prob.plan
plan.set_lenscales([4.0])
D = plan.D
m32 = torch.randn(D, device="cuda") * 0.01
A = torch.randn(D, D, device="cuda") * 0.01
C32 = (A @ A.T + torch.eye(D, device="cuda")).contiguous()
ts = []
for _ in range(5):
    prob.rflat.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.slm_gradpass(plan, prob.Xd, prob.yd, m32, C32, prob.R, prob.sqerr, engine=prob.engine)
    b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print("gradpass ms", " ".join("%.2f" % t for t in ts), "overlap" if not os.environ.get("RR_GP_NO_OVERLAP") else "serial")
