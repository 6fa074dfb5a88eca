#!/usr/bin/env python
"""Value pass of config 2 with and without skipping the below-diagonal columns of the
first tile of every row panel (REVRAND_B200_T3_DIAG_SKIP), the two modes interleaved."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from revrand_b200 import _engine
from revrand_b200.basis_functions import RandomMatern32
from revrand_b200.slm import _SLMProblem
N, d, K = 1000000, 21, int(os.environ.get("K", "2048"))
X, y = bench.synthetic(N, d)
prob = _SLMProblem(RandomMatern32(nbases=K, Xdim=d, random_state=1), X, y)
prob.plan.set_lenscales([4.0])
st = prob.stats
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = {"0": [], "1": []}
ref = None
for i in range(24):
    mode = "1" if i % 2 == 0 else "0"
    os.environ["REVRAND_B200_T3_DIAG_SKIP"] = mode
    st.zero_()
    flush.zero_()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _engine.slm_suffstats(prob.plan, prob.Xd, prob.yd, st, engine=prob.engine, want_yy=False)
    b.record()
    torch.cuda.synchronize()
    ts[mode].append(a.elapsed_time(b))
    if ref is None:
        ref = st.G.clone()
    assert torch.equal(ref, st.G)
print(json.dumps({"skip_ms": float(np.mean(ts["1"][2:])), "full_ms": float(np.mean(ts["0"][2:])),
                  "skip_all": [round(v, 2) for v in ts["1"]], "full_all": [round(v, 2) for v in ts["0"]]}))
