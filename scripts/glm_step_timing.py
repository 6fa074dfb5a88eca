"""Config 4 (BASELINE.json): GLM Poisson + RandomRBF(1024), minibatch 8192 of
N=1e6, d=21, K_mix=10, L=50: time of one SVI step (one _elbo call: device part
rr_glm_step + host assembly).  Reference CPU: 3.4 s / step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.build()
import numpy as np
import torch
import revrand_b200 as rr
from revrand_b200 import basis_functions as bf, likelihoods as lk, Parameter, Positive

rs = np.random.RandomState(0)
M, d, K, Kmix, L = 8192, 21, 1024, 10, 50
X = rs.randn(M, d)
y = rs.poisson(np.exp(np.sin(X[:, 0]))).astype(float)
basis = bf.RandomRBF(nbases=K, Xdim=d, random_state=1, lenscale=Parameter(1.0, Positive()))
glm = rr.GeneralizedLinearModel(likelihood=lk.Poisson('exp'), basis=basis, K=Kmix, nsamples=L,
                                batch_size=M, random_state=2)
D = 2 * K
glm.B_, glm.D_, glm._it = 1e6 / M, D, 1     # _it > 0: no ELBO logging on this step
m = rs.randn(D, Kmix) * 0.1
C = np.abs(rs.randn(D, Kmix)) * 0.1 + 0.05
ts = []
for i in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = glm._elbo(m, C, 1.0, [], 1.0, X, y)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
print("glm step wall ms (M=8192, K=1024, Kmix=10, L=50):", " ".join("%.1f" % (1e3 * t) for t in ts))
