#!/usr/bin/env python
"""Round-2 quick GPU check: int8 self-tests, small parity, value-pass timing at
config 2 for the int8 engine and the round-1 fused kind::f16 kernel."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import oracle as orc  # noqa: E402
from revrand_b200 import _cabi, _engine  # noqa: E402
from revrand_b200 import basis_functions as bf  # noqa: E402


def log(*a):
    print(*a, flush=True)


def main():
    torch.cuda.set_device(0)
    for kb in (1, 3, 40):
        t0 = time.time()
        bad = _engine.tcgen05_i8_selftest(kb)
        log("i8 selftest kblocks=%d mismatches=%d (%.1f s)" % (kb, bad, time.time() - t0))
    rs = np.random.RandomState(0)
    for (N, d, K) in [(129, 3, 70), (5000, 21, 64), (40000, 21, 200)]:
        X = rs.randn(N, d).astype(np.float32).astype(np.float64)
        y = np.sin(X[:, 0]) + 0.1 * rs.randn(N)
        b = bf.RandomMatern32(nbases=K, Xdim=d, random_state=1)
        plan = b._plan(d, [2.0])
        Phi = orc.trig_features(X, b.W, 2.0)
        Gref, pref = Phi.T.dot(Phi), Phi.T.dot(y.astype(np.float32).astype(float))
        Xd, yd = _engine.to_device(X), _engine.to_device(y)
        for name, e in (("i8", _cabi.RR_ENGINE_TCGEN05), ("simt", _cabi.RR_ENGINE_SIMT)):
            st = _engine.SuffStats(plan.D)
            _engine.slm_suffstats(plan, Xd, yd, st, engine=e)
            torch.cuda.synchronize()
            G = st.G.cpu().numpy()
            log("N=%d d=%d K=%d %s: G relerr %.2e  p relerr %.2e  asym %.1e" % (
                N, d, K, name, np.linalg.norm(G - Gref) / np.linalg.norm(Gref),
                np.linalg.norm(st.p.cpu().numpy() - pref) / np.linalg.norm(pref),
                np.abs(G - G.T).max()))
    # ---- timing at config 2 ------------------------------------------------------
    N, d, K = 1000000, 21, 2048
    X = rs.randn(N, d).astype(np.float32)
    y = np.sin(X[:, 0]).astype(np.float32)
    b = bf.RandomMatern32(nbases=K, Xdim=d, random_state=1)
    plan = b._plan(d, [4.0])
    Xd, yd = _engine.to_device(X), _engine.to_device(y)
    st = _engine.SuffStats(plan.D)
    flops = 2.0 * N * (2 * K) ** 2 + 2.0 * N * d * K
    res = {}
    for name, e in (("i8", _cabi.RR_ENGINE_TCGEN05), ("fused16", _cabi.RR_ENGINE_TCGEN05_FUSED16)):
        for _ in range(2):
            st.zero_()
            _engine.slm_suffstats(plan, Xd, yd, st, engine=e, want_yy=False)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            st.zero_()
            a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _engine.slm_suffstats(plan, Xd, yd, st, engine=e, want_yy=False)
            bb.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(bb))
        res[name] = st.G.clone()
        log("config2 value pass %-8s: %s ms  -> %.0f TFLOP/s algorithmic" % (
            name, " ".join("%.2f" % t for t in ts), flops / (min(ts) * 1e-3) / 1e12))
        log("   trace/N = %.9f" % (st.G.diagonal().sum().item() / N))
    log("i8 vs fused16 G: rel diff %.2e" % ((res["i8"] - res["fused16"]).norm() / res["i8"].norm()).item())


if __name__ == "__main__":
    main()
