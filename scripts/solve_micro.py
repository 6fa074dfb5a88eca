#!/usr/bin/env python
"""Micro-timings of the float64 pieces of the posterior solve (CUDA events, median)."""
import json
import sys
import numpy as np
import torch

t = torch
dev = "cuda"


def timed(fn, reps=7):
    fn()
    ts = []
    for _ in range(reps):
        t.cuda.synchronize()
        a, b = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        t.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return round(float(np.median(ts)), 4)


out = {}
g = t.Generator(device=dev).manual_seed(0)
for n in (512, 1024, 2048, 4096):
    A = t.randn(n, n + 64, dtype=t.float64, device=dev, generator=g)
    S = A @ A.T + n * t.eye(n, dtype=t.float64, device=dev)
    out["potrf_%d" % n] = timed(lambda: t.linalg.cholesky_ex(S))
    L = t.linalg.cholesky_ex(S)[0]
    B = t.randn(n, n, dtype=t.float64, device=dev, generator=g)
    # X L^T = B  (right-side solve, as L21 = A21 L11^-T)
    out["trsm_right_%d" % n] = timed(lambda: t.linalg.solve_triangular(L.T, B, upper=True, left=False))
    out["gemm_nt_%d" % n] = timed(lambda: B @ B.T)
    Bs = B[: n // 8]
    out["trsm_right_%d_rows8th" % n] = timed(lambda: t.linalg.solve_triangular(L.T, Bs, upper=True, left=False))
    out["gemm_nt_%d_cols8th" % n] = timed(lambda: B @ Bs.T)
    Li = t.linalg.solve_triangular(L, t.eye(n, dtype=t.float64, device=dev), upper=False)
    out["gemm_via_inverse_%d" % n] = timed(lambda: B @ Li.T)
print(json.dumps(out))
