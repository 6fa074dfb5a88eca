#!/usr/bin/env python
"""Generate ``tests/golden/config2.npz``: the float64 posterior, log marginal
likelihood and gradients of BASELINE config 2 (StandardLinearModel +
RandomMatern32(nbases=2048), N=1e6, d=21) at the six evaluation points that
``bench.py`` times.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).  Run in the build container
(needs ``/root/reference`` for the seeded frequency matrix W; ~30 min on 8
cores):

    python oracle/gen_golden_config2.py [--N 1000000] [--K 2048] [--threads 8]

The arithmetic is the oracle's row-chunked restatement of
``StandardLinearModel._elbo`` (revrand/slm.py:142-199): pass 1 accumulates
G = Phi^T Phi and Phi^T y with the oracle's own feature map (which restates
basis_functions.py:859-864); the solve is ``orc.solve_posdef``
(mathfun/linalg.py:84-125); pass 2 forms the residuals (slm.py:161-162) and
the lengthscale gradients (slm.py:193-197).  Two evaluation points share a
lengthscale, so they share the trigonometric work of both passes.

The gradient is stored for ALL d input dimensions ("ARD components at a common
lengthscale"): component i is  d(-ELBO)/d l_i = -(m^T dPhi_i^T Err -
sum(dPhi_i^T Phi o C)) / var  with dPhi_i of basis_functions.py:897-899.  The
trace term is evaluated as sum(dPhi_i o (Phi C)) -- the same number, without
the d Gram-sized products the reference forms.  The reference's own return
value for a scalar lengthscale is component 0 (Appendix B #1 of SURVEY.md);
the mathematically complete isotropic derivative is the sum.  Before the big
run the script checks this pass-2 formulation against ``orc.slm_elbo`` (ARD,
monolithic, itself pinned to the unmodified reference by gen_golden.py) on a
small case and aborts on disagreement.
"""

import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("REVRAND_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
if not hasattr(np, "asscalar"):
    np.asscalar = lambda a: a.item()  # noqa: E731

from oracle import oracle as orc  # noqa: E402


def synthetic(N, d, seed=0):
    """Same generator as bench.synthetic."""
    rs = np.random.RandomState(seed)
    X = rs.randn(N, d).astype(np.float32)
    w = rs.randn(d)
    y = (np.sin(X.astype(np.float64).dot(w) / 3.0)
         + 0.1 * rs.randn(N)).astype(np.float32)
    return X, y


EVAL_LS = [1.0, 4.0, 10.0]
EVAL_VAR = [0.02, 1.0]
REG = 1.0


def evaluate_points(X, y, W, ls, variances, reg, chunk, log=print):
    """All evaluation points that share lengthscale ``ls``."""
    N, d = X.shape
    K = W.shape[1]
    D = 2 * K
    blocks = [dict(kind="trig", W=W, lenscale=ls, cols=None)]
    G = np.zeros((D, D))
    p = np.zeros(D)
    t0 = time.time()
    for s in range(0, N, chunk):
        Phi = orc.concat_features(X[s:s + chunk], blocks)
        G += Phi.T.dot(Phi)
        p += Phi.T.dot(y[s:s + chunk])
    log("  ls=%g pass 1: %.0f s" % (ls, time.time() - t0))
    Ld, slices = orc.regularizer_diagonal(X[:1], blocks, [reg])
    iL = 1.0 / Ld
    posts = []
    for var in variances:
        iC = np.diag(iL) + G / var
        C, logdetiC = orc.solve_posdef(iC, np.eye(D))
        m = C.dot(p) / var
        posts.append(dict(var=var, C=C, m=m, logdet=logdetiC,
                          trgc=(G * C).sum(), sqerr=0.0,
                          R=np.zeros((d, K))))
    t0 = time.time()
    sK = np.sqrt(K)
    for s in range(0, N, chunk):
        Xc, yc = X[s:s + chunk], y[s:s + chunk]
        Phi = orc.concat_features(Xc, blocks)
        cosb, sinb = Phi[:, :K], Phi[:, K:]
        for P in posts:
            Err = yc - Phi.dot(P["m"])
            P["sqerr"] += (Err ** 2).sum()
            T = np.outer(Err, P["m"]) - Phi.dot(P["C"])
            # dPhi_i[n, :] = x_ni * (-W_i / l^2) o [-sin | cos] / sqrt(K); Phi
            # already carries the 1 / sqrt(K)
            Q = -sinb * T[:, :K] + cosb * T[:, K:]
            P["R"] += Xc.T.dot(Q)
        del Phi
    log("  ls=%g pass 2: %.0f s" % (ls, time.time() - t0))
    out = []
    for P in posts:
        var, m, C = P["var"], P["m"], P["C"]
        dC = C.diagonal().copy()
        ELBO = -0.5 * (N * np.log(2 * np.pi * var) + P["sqerr"] / var
                       + P["trgc"] / var + ((m ** 2 + dC) * iL).sum()
                       + P["logdet"] + np.log(Ld).sum() - D)
        dvar = 0.5 * (-N + (P["sqerr"] + P["trgc"]) / var) / var
        dreg = -0.5 * (((m ** 2 + dC) * iL ** 2).sum() - iL.sum())
        # -(m^T dPhi_i^T Err - sum(dPhi_i^T Phi o C)) / var with
        # dPhi_i = x_i (-W_i / l^2) o [-sin | cos] / sqrt(K):
        #   = (1 / (var l^2)) sum_k W_ik R_ik,  R = X^T Q,  T = Err (x) m - Phi C
        dl = (W * P["R"]).sum(axis=1) / (var * ls ** 2)
        out.append(dict(ls=ls, var=var, neg_elbo=-ELBO, m=m, diagC=dC,
                        logdet=P["logdet"], trgc=P["trgc"], sqerr=P["sqerr"],
                        dvar=-dvar, dreg=dreg, dl=dl,
                        cond=float(np.linalg.cond(np.diag(iL) + G / var))
                        if D <= 512 else np.nan))
    return out


def selfcheck():
    """Pin the pass-2 formulation used here to the oracle's monolithic ARD
    ``slm_elbo`` (reference formulation with explicit dPhi)."""
    rs = np.random.RandomState(5)
    N, d, K = 1500, 4, 24
    X = rs.randn(N, d)
    y = np.sin(X.sum(axis=1)) + 0.1 * rs.randn(N)
    W = rs.randn(d, K)
    for ls, var in [(0.7, 0.05), (3.0, 1.0)]:
        mine = evaluate_points(X, y, W, ls, [var], 1.3, chunk=400,
                               log=lambda *_: None)[0]
        blocks = [dict(kind="trig", W=W, lenscale=np.full(d, ls), cols=None)]
        ref = orc.slm_elbo(X, y, var, [1.3], blocks)
        chk = [("neg_elbo", mine["neg_elbo"], ref["neg_elbo"]),
               ("m", mine["m"], ref["m"]),
               ("diagC", mine["diagC"], ref["C"].diagonal()),
               ("dvar", mine["dvar"], ref["dvar"]),
               ("dreg", mine["dreg"], ref["dreg"][0]),
               ("dl", mine["dl"], ref["dhyp"][0])]
        for name, a, b in chk:
            if not np.allclose(a, b, rtol=1e-9, atol=1e-11):
                raise SystemExit("SELFCHECK MISMATCH %s: %r vs %r" % (name, a, b))
        # scalar lengthscale through the oracle's (reference-compatible) path:
        # component 0 only
        blocks = [dict(kind="trig", W=W, lenscale=ls, cols=None)]
        iso = orc.slm_elbo_chunked(X, y, var, [1.3], blocks, chunk=500)
        if not np.allclose(iso["dhyp"][0], mine["dl"][0], rtol=1e-9, atol=1e-11):
            raise SystemExit("SELFCHECK MISMATCH isotropic compat component")
    print("selfcheck OK: pass-2 formulation == oracle slm_elbo (ARD) to 1e-9")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=1000000)
    ap.add_argument("--d", type=int, default=21)
    ap.add_argument("--K", type=int, default=2048)
    ap.add_argument("--chunk", type=int, default=20000)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "config2.npz"))
    args = ap.parse_args()
    if args.threads:
        from threadpoolctl import threadpool_limits
        threadpool_limits(args.threads)
    selfcheck()
    from revrand import basis_functions as rbf  # the unmodified reference
    W = rbf.RandomMatern32(nbases=args.K, Xdim=args.d, random_state=1).W
    X32, y32 = synthetic(args.N, args.d)
    X, y = X32.astype(np.float64), y32.astype(np.float64)
    res = []
    for ls in EVAL_LS:
        print("lengthscale %g" % ls, flush=True)
        res += evaluate_points(X, y, W, ls, EVAL_VAR, REG, args.chunk,
                               log=lambda s: print(s, flush=True))
    out = {"N": args.N, "d": args.d, "K": args.K, "reg": REG,
           "W_checksum": float(np.abs(W).sum()),
           "points": np.array([[r["ls"], r["var"]] for r in res])}
    for key in ("neg_elbo", "logdet", "trgc", "sqerr", "dvar", "dreg"):
        out[key] = np.array([r[key] for r in res])
    for key in ("m", "diagC", "dl"):
        out[key] = np.stack([r[key] for r in res])
    np.savez_compressed(args.out, **out)
    for r in res:
        print("ls=%g var=%g -ELBO=%.10g |m|=%.6g dl0=%.6g dl_sum=%.6g"
              % (r["ls"], r["var"], r["neg_elbo"], np.linalg.norm(r["m"]),
                 r["dl"][0], r["dl"].sum()))
    print("wrote", args.out)


if __name__ == "__main__":
    main()
