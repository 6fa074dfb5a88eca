"""Where the end-to-end call (host arrays in, scalars out) spends its time beyond
the device-resident evaluation: upload, scales, host assembly."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from revrand_b200 import StandardLinearModel, config
from revrand_b200.basis_functions import RandomMatern32
from bench import synthetic, EVAL_POINTS, REG

N, d, K = 1000000, 21, 2048
X, y = synthetic(N, d)
print("X dtype", X.dtype, "y dtype", y.dtype)
basis = RandomMatern32(nbases=K, Xdim=d, random_state=1)
Xh, yh = torch.from_numpy(X).pin_memory(), torch.from_numpy(y).pin_memory()
slm = StandardLinearModel(basis=basis)
slm.obj_ = -np.inf
config.CACHE_DEVICE_DATA = False


def wall(fn, reps=4):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    return float(np.median(ts))


ls, var = EVAL_POINTS[1]
print("e2e _elbo(host X, y): %.2f ms" % wall(lambda: slm._elbo(Xh, yh, var, REG, ls)))
prob = slm._cached_problem
print("upload only: %.2f ms" % wall(lambda: prob.upload(Xh, yh)))
print("Xd.copy_ only: %.2f ms" % wall(lambda: prob.Xd.copy_(Xh, non_blocking=True)))
print("col scale only: %.2f ms" % wall(lambda: prob._refresh_col_scale()))
print("evaluate (resident): %.2f ms" % wall(lambda: prob.evaluate(var, [REG], [ls])))
print("evaluate value-only: %.2f ms" % wall(lambda: prob.evaluate(var, [REG], [ls], want_grad=False)))
config.CACHE_DEVICE_DATA = True
slm._problem_key = None
print("e2e _elbo cached data: %.2f ms" % wall(lambda: slm._elbo(Xh, yh, var, REG, ls)))
print("fingerprint: %.2f ms" % wall(lambda: slm._fingerprint(Xh, yh)))
# ---- stages of the uncached call ---------------------------------------------------
config.CACHE_DEVICE_DATA = False


def stage_times():
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    p = slm._get_problem(Xh, yh)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    r = p.evaluate(var, [REG], [ls], want_grad=True)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return 1e3 * (t1 - t0), 1e3 * (t2 - t1)


for _ in range(4):
    print("get_problem %.2f ms, evaluate %.2f ms" % stage_times())
for _ in range(4):
    t0 = time.perf_counter()
    slm._elbo(Xh, yh, var, REG, ls)
    torch.cuda.synchronize()
    print("_elbo %.2f ms (obj_ %.10f)" % (1e3 * (time.perf_counter() - t0), slm.obj_))
