"""Time the one-GPU posterior solve at D=4096 for several leaf sizes of the
blocked inverse."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from revrand_b200 import _engine as eng
D = 4096
rs = np.random.RandomState(0)
A = torch.from_numpy(rs.randn(D, 2 * D)).cuda()
G = (A @ A.T) / D
p = torch.from_numpy(rs.randn(D)).cuda()
lam = torch.ones(D, dtype=torch.float64, device="cuda")
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
for leaf in (128, 256, 512, 1024, 2048):
    eng._BLOCK_INV_LEAF = leaf
    print("leaf %4d: solve (need_C) %.2f ms, value-only %.2f ms" % (
        leaf, timed(lambda: eng.solve_posterior(G, p, 0.02, lam, need_C=True)),
        timed(lambda: eng.solve_posterior(G, p, 0.02, lam, need_C=False))), flush=True)
L = torch.linalg.cholesky(G / 0.02 + torch.eye(D, dtype=torch.float64, device="cuda"))
print("potrf alone %.2f ms; potri %.2f ms" % (timed(lambda: torch.linalg.cholesky_ex(G / 0.02 + torch.eye(D, dtype=torch.float64, device="cuda"))),
                                            timed(lambda: torch.cholesky_inverse(L))))
