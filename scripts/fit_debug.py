#!/usr/bin/env python
"""Trace every evaluation of a config-1 fit (arguments, value, gradient norm)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import revrand_b200 as rr  # noqa: E402
from revrand_b200 import basis_functions as bf  # noqa: E402


def main():
    torch.cuda.set_device(0)
    rs = np.random.RandomState(0)
    X = np.sort(rs.uniform(-5, 5, size=(1000, 1)), axis=0)
    y = np.sin(X[:, 0]) + 0.1 * rs.randn(1000)
    slm = rr.StandardLinearModel(basis=bf.RandomRBF(nbases=256, Xdim=1, random_state=1),
                                 nstarts=100, maxiter=200, random_state=2)
    orig = slm._elbo
    n = [0]

    def traced(Xa, ya, var, reg, hyp, want_grad=True):
        r = orig(Xa, ya, var, reg, hyp, want_grad=want_grad)
        n[0] += 1
        if n[0] <= 3 or n[0] > 98 or not np.isfinite(r[0]):
            print(n[0], "grad" if want_grad else "val ", float(var), float(reg), float(hyp), r[0],
                  None if r[1] is None else [float(np.ravel(g)[0]) for g in r[1]], flush=True)
        return r
    slm._elbo = traced
    slm.fit(X, y)
    print("evals", n[0], "ELBO", slm.obj_, slm.var_, slm.regularizer_, slm.hypers_)
    # the reference (unmodified, same seeds): 128 evaluations, ELBO 884.1131,
    # var 0.0093828, reg 2.60454, lenscale 2.31472; start 41 is the best candidate:
    print("reference best start: var 0.010302932803516844 reg 0.9925232900583055 "
          "ls 0.3498019106085425 -> -818.7285065746694")
    r = orig(X, y, 0.010302932803516844, 0.9925232900583055, 0.3498019106085425)
    print("ours at that point:", r)
    r = orig(X, y, 8.21398677751374e-22, 1.717836909102504e-05, 4.2058912125146726e+36)
    print("ours at the reference's 2nd L-BFGS point (ref 5.3744275379158945e+23):", r)
    prob = slm._get_problem(X, y)
    e = prob.evaluate(8.21398677751374e-22, [1.717836909102504e-05], [4.2058912125146726e+36])
    print({k: (v if np.ndim(v) == 0 else np.asarray(v).ravel()[:4]) for k, v in e.items()
           if k in ("logdet", "trgc", "sqerr", "cond_est", "q", "g")})
    print("m finite:", bool(torch.isfinite(e["m"]).all()), "max |m|", float(e["m"].abs().max()))


if __name__ == "__main__":
    main()
