"""Debug: build the library with -DRR_T2_TRACE, run the fused value pass once at
the config-2 shape and print the per-slab event timeline of CTA 0 (clock64
deltas).  Events: gens (warp 4 / last gen warp): 0 wait-U start, 1 U ready,
2 U loaded, 5/6 Phi-stage wait start/end, 3 body done, 4 Phi published;
MMA warp: 8 loop top, 9 Phi ready, 10 projection(t+2) issued, 11 Gram issued;
loader: 12 tile wait, 13 tile free, 14 tile published."""
import ctypes as C
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "revrand_b200", "csrc")
LIB = os.path.join(ROOT, "revrand_b200", "lib", "librevrand_b200_trace.so")
os.makedirs(os.path.dirname(LIB), exist_ok=True)
if not os.path.exists(LIB) or "--rebuild" in sys.argv:
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode",
                           "arch=compute_100a,code=sm_100a", "-lineinfo", "-DRR_T2_TRACE"]
                          + [a for a in sys.argv[1:] if a.startswith("-D")] + [
                           "-Xcompiler", "-fPIC", "-shared", "-o", LIB]
                          + sorted(glob.glob(os.path.join(CSRC, "*.cu"))), cwd=CSRC)
if "--build-only" in sys.argv:
    sys.exit(0)
import numpy as np
import torch
from revrand_b200 import _cabi
_cabi.LIB_PATH = LIB
from revrand_b200 import _engine as eng
from revrand_b200.basis_functions import RandomMatern32
from revrand_b200.slm import _SLMProblem
from bench import synthetic

X, y = synthetic(1000000, 21)
prob = _SLMProblem(RandomMatern32(nbases=2048, Xdim=21, random_state=1), X, y)
prob.plan.set_lenscales([4.0])
for _ in range(2):
    prob.stats.zero_()
    eng.slm_suffstats(prob.plan, prob.Xd, prob.yd, prob.stats, engine=prob.engine)
torch.cuda.synchronize()
lib = _cabi.load()
lib.rr_debug_t2_trace.restype = C.c_int
NS, SL = 48, 32
buf = (C.c_longlong * (NS * SL))()
m = lib.rr_debug_t2_trace(buf, NS * SL)
tr = np.array(buf, dtype=np.int64).reshape(NS, SL)
t0 = tr[0][tr[0] > 0].min()
names = {0: "g0.waitU", 1: "g0.Uready", 2: "g0.Uloaded", 5: "g0.stWait", 6: "g0.stFree",
         3: "g0.bodyDone", 4: "g0.published",
         16: "gL.waitU", 17: "gL.Uready", 18: "gL.Uloaded", 21: "gL.stWait", 22: "gL.stFree",
         19: "gL.bodyDone", 20: "gL.published",
         8: "mma.top", 9: "mma.phiReady", 10: "mma.projIssued", 11: "mma.gramIssued",
         12: "ld.tileWait", 13: "ld.tileFree", 14: "ld.published"}
for s in range(NS):
    ev = sorted((int(tr[s][k] - t0), names[k]) for k in names if tr[s][k] > 0)
    print("slab %d: " % s + "  ".join("%s@%d" % (n, t) for t, n in ev))
per = np.diff(tr[:, 4])
print("period (g0.published) mean %.0f min %d max %d" % (per.mean(), per.min(), per.max()))
for a, b, lab in [(0, 1, "g0 wait for U"), (1, 2, "g0 U load"), (2, 5, "g0 trig until first store"),
                  (5, 6, "g0 wait for Phi stage"), (6, 3, "g0 rest of body"), (3, 4, "g0 fence+publish"),
                  (8, 9, "mma wait Phi"), (9, 10, "mma issue projection (waits X,U)"),
                  (10, 11, "mma issue Gram")]:
    dlt = tr[:, b] - tr[:, a]
    print("%-36s mean %7.0f  min %6d  max %6d" % (lab, dlt.mean(), dlt.min(), dlt.max()))
