"""Config-5 shaped evaluation: BasisCat(RandomRBF(K) + LinearBasis(onescol)) at
N=1e6, d=21 through the tcgen05 engines (affine columns as pseudo-frequency
slots / extra reduction columns); per-evaluation time for a few K."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.build()
import numpy as np
import torch
from revrand_b200 import basis_functions as bf, Parameter, Positive
from revrand_b200.slm import _SLMProblem
from bench import synthetic

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
X, y = synthetic(N, 21)
for K in (512, 2048):
    basis = bf.RandomRBF(nbases=K, Xdim=21, random_state=1) + bf.LinearBasis(onescol=True)
    prob = _SLMProblem(basis, X, y)
    assert prob.uses_tcgen05()
    ts = []
    for i in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = prob.evaluate(0.02, [1.0, 1.0], [4.0], want_grad=True)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("K=%d D=%d: value+grad eval ms %s" % (K, prob.D, " ".join("%.1f" % t for t in ts)), flush=True)
    del prob
    torch.cuda.empty_cache()
