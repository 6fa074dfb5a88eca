"""Posterior-mean / log-ML / gradient errors of the tcgen05 engine against the
float64 oracle at a few (N, K, lengthscale, var) points."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.build()
import numpy as np
import revrand_b200 as rr
from revrand_b200 import Parameter, Positive, config
from revrand_b200 import basis_functions as bf
from oracle import oracle as orc

def synth(N, d, seed):
    rs = np.random.RandomState(seed)
    X = rs.randn(N, d).astype(np.float32).astype(np.float64)
    w = rs.randn(d)
    return X, np.sin(X.dot(w) / 3.0) + 0.1 * rs.randn(N)

for (N, d, K, ls, var) in [(20011, 21, 200, 3.0, 0.05), (60000, 21, 256, 4.0, 0.02),
                           (60000, 21, 256, 1.0, 0.02), (30000, 3, 256, 1.0, 0.02),
                           (30000, 1, 256, 0.6, 0.01)]:
    X, y = synth(N, d, 11)
    b = bf.RandomMatern32(nbases=K, Xdim=d, random_state=4, lenscale=Parameter(float(ls), Positive()))
    blocks = [dict(kind="trig", W=b.W, lenscale=ls, cols=None)]
    ref = orc.slm_elbo(X, y, var, [1.0], blocks)
    out = []
    for engine in ("tcgen05", "simt"):
        config.ENGINE = engine
        slm = rr.StandardLinearModel(basis=b); slm.obj_ = -np.inf
        nelbo, (dv, dr, dl) = slm._elbo(X, y, var, 1.0, ls)
        rel = lambda a, r: float(np.linalg.norm(np.ravel(a) - np.ravel(r)) / np.linalg.norm(np.ravel(r)))
        out.append("%s: m %.1e elbo %.1e diagC %.1e dvar %.1e dls %.1e" % (
            engine, rel(slm.weights_, ref["m"]), abs(nelbo - ref["neg_elbo"]) / abs(ref["neg_elbo"]),
            rel(slm.covariance_.diagonal(), ref["C"].diagonal()), abs(dv - ref["dvar"]) / abs(ref["dvar"]),
            abs(dl - ref["dhyp"][0]) / abs(ref["dhyp"][0])))
    iC = np.diag(np.ones(2 * K)) + None if False else None
    print("N=%d d=%d K=%d ls=%g var=%g | %s | %s" % (N, d, K, ls, var, out[0], out[1]), flush=True)
