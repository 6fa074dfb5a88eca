#!/usr/bin/env python
"""Kernel-time budget of the row-sharded posterior solve as ONE rank of `--world` sees
it, on a single GPU: `_engine.world` is faked and the all-gathers are replaced by local
copies (so: compute + copies, no NVLink time).  Prints a torch-profiler table grouped by
kernel and the total device time of one solve."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from revrand_b200 import _engine as eng

ap = argparse.ArgumentParser()
ap.add_argument("--D", type=int, default=4096)
ap.add_argument("--world", type=int, default=8)
a = ap.parse_args()
t = torch
D, ws = a.D, a.world
g = t.Generator(device="cuda").manual_seed(0)
A = t.randn(D, D + 64, dtype=t.float64, device="cuda", generator=g)
G = A @ A.T
p = t.randn(D, dtype=t.float64, device="cuda", generator=g)
lam = t.ones(D, dtype=t.float64, device="cuda")

eng.world = lambda: (0, ws)


def fake_gather(out, inp, *k, **kw):
    out.view(ws, -1).copy_(inp.reshape(1, -1).expand(ws, -1))


t.distributed.all_gather_into_tensor = fake_gather


def solve():
    post = eng.solve_posterior(G, p, float(D), lam, need_C=True, defer_check=True)
    return post.C32()


for _ in range(3):
    solve()
t.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    solve()
    e1.record()
    t.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("solve (rank 0 of %d, D=%d, no network): %.3f ms (median of 5)" % (ws, D, float(np.median(ts))))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    solve()
    t.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
