"""Run only the fused value pass a few times at the config-2 shape (for ncu)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.build()
import torch
from revrand_b200 import _engine as eng
from revrand_b200.basis_functions import RandomMatern32
from revrand_b200.slm import _SLMProblem
from bench import synthetic

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
X, y = synthetic(N, 21)
prob = _SLMProblem(RandomMatern32(nbases=2048, Xdim=21, random_state=1), X, y)
prob.plan.set_lenscales([4.0])
for _ in range(reps):
    prob.stats.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.slm_suffstats(prob.plan, prob.Xd, prob.yd, prob.stats, engine=prob.engine)
    b.record()
    torch.cuda.synchronize()
    print("suffstats ms", a.elapsed_time(b))
