#!/usr/bin/env python
"""sum Err^2 at the six config-2 points three ways, against the float64 oracle golden:
from the kept fp16 image (gradient pass), from the float64 statistics, from the fp32
residual pass."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from revrand_b200 import _engine as eng
from revrand_b200 import basis_functions as bf
from revrand_b200.slm import _SLMProblem
g = np.load("tests/golden/config2.npz")
N, d, K = int(g["N"]), int(g["d"]), int(g["K"])
X, y = bench.synthetic(N, d)
prob = _SLMProblem(bf.RandomMatern32(nbases=K, Xdim=d, random_state=1), X, y)
for i, (ls, var) in enumerate(g["points"]):
    r = prob.evaluate(float(var), [float(g["reg"])], [float(ls)], want_grad=True)
    st, m = prob.stats, r["m"]
    s_stats = float((prob.yy - 2.0 * st.p.dot(m) + m.dot(st.G @ m)).item())
    s_res = float(eng.slm_residual(prob.plan, prob.Xd, prob.yd, m.float().contiguous()).item())
    ref = float(g["sqerr"][i])
    print("ls=%g var=%g  kept %.2e  stats %.2e  fp32 residual pass %.2e   (sqerr/yy = %.3f)"
          % (ls, var, abs(r["sqerr"] - ref) / ref, abs(s_stats - ref) / ref, abs(s_res - ref) / ref,
             ref / prob.yy))
