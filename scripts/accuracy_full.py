"""Full-size (config 2: N=1e6, d=21, K=2048) agreement of the fused tcgen05
engine with the SIMT engine (fp32 features, float64 accumulation; itself within
~1e-7 of the float64 oracle at the sizes the oracle can run) at the six
evaluation points of bench.py."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.build()
import numpy as np
import torch
from revrand_b200 import config, _engine as eng
from revrand_b200.basis_functions import RandomMatern32
from revrand_b200.slm import _SLMProblem
from bench import synthetic, EVAL_POINTS, REG

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
X, y = synthetic(N, 21)
basis = RandomMatern32(nbases=K, Xdim=21, random_state=1)
res = {}
for engine in ("tcgen05", "tcgen05_fine", "simt"):
    config.ENGINE = engine
    prob = _SLMProblem(basis, X, y)
    for (ls, var) in EVAL_POINTS:
        r = prob.evaluate(var, [REG], [ls], want_grad=True)
        post = r["post"]
        res[(engine, ls, var)] = dict(m=post.m.cpu().numpy(), dC=post.diagC.cpu().numpy(),
                                      logdet=float(r["logdet"]), trgc=float(r["trgc"]),
                                      sqerr=float(r["sqerr"]), g=np.array(r["g"]).ravel(),
                                      cond=float(post.cond_est), dmax=float((post.diagC).max()))
    del prob
    torch.cuda.empty_cache()
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
for (ls, var) in EVAL_POINTS:
  for eng_name in ("tcgen05", "tcgen05_fine"):
    a, b = res[(eng_name, ls, var)], res[("simt", ls, var)]
    print("%-12s ls=%-4g var=%-5g cond_est %.1e | m %.1e  diagC %.1e  logdet %.1e  tr(GC) %.1e  sqerr %.1e  dls(sum) %.1e"
          % (eng_name, ls, var, b["cond"], rel(a["m"], b["m"]), rel(a["dC"], b["dC"]),
             abs(a["logdet"] - b["logdet"]) / abs(b["logdet"]), abs(a["trgc"] - b["trgc"]) / abs(b["trgc"]),
             abs(a["sqerr"] - b["sqerr"]) / abs(b["sqerr"]), abs(a["g"].sum() - b["g"].sum()) / abs(b["g"].sum())),
          flush=True)
