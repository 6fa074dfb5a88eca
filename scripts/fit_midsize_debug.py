#!/usr/bin/env python
"""Directional finite-difference check of the mid-size ARD fit's objective/gradient
and a trace of its L-BFGS-B evaluations."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import revrand_b200 as rr  # noqa: E402
from revrand_b200 import Parameter, Positive  # noqa: E402
from revrand_b200 import basis_functions as bf  # noqa: E402

torch.cuda.set_device(0)
rs = np.random.RandomState(0)
rs.uniform(-5, 5, size=(1000, 1)); rs.randn(1000)
N, d, K = 40000, 21, 512
X = rs.randn(N, d)
y = np.sin(X.dot(rs.randn(d)) / 3.0) + 0.1 * rs.randn(N)
slm = rr.StandardLinearModel(
    basis=bf.RandomMatern32(nbases=K, Xdim=d, random_state=1,
                            lenscale=Parameter(3.0 * np.ones(d), Positive())),
    nstarts=20, maxiter=40, random_state=2)
orig = slm._elbo
n = [0]


def traced(Xa, ya, var, reg, hyp, want_grad=True):
    r = orig(Xa, ya, var, reg, hyp, want_grad=want_grad)
    n[0] += 1
    if want_grad:
        print(n[0], float(var), float(reg), np.round(np.asarray(hyp), 3)[:4], r[0],
              float(r[1][0]), float(r[1][1]), np.asarray(r[1][2])[:4], flush=True)
    return r


slm._elbo = traced
slm.fit(X, y)
print("message:", slm.opt_message_, "ELBO", slm.obj_)
# finite differences in log-space at the first gradient point
var, reg, ls = slm.var_, slm.regularizer_, np.asarray(slm.hypers_, dtype=float)
slm.obj_ = -np.inf
f0, (dv, dr, dl) = orig(X, y, var, reg, ls)
g = np.concatenate([[dv * var], [dr * reg], np.asarray(dl) * ls])      # d/d log
print("f0", f0, "|g_log|", np.linalg.norm(g), g[:5])
for which in range(5):
    e = np.zeros_like(g)
    e[which] = 1.0
    for eps in (1e-2, 1e-3, 1e-4):
        z = np.log(np.concatenate([[var], [reg], ls])) + eps * e
        zm = np.log(np.concatenate([[var], [reg], ls])) - eps * e
        fp = orig(X, y, float(np.exp(z[0])), float(np.exp(z[1])), np.exp(z[2:]))[0]
        fm = orig(X, y, float(np.exp(zm[0])), float(np.exp(zm[1])), np.exp(zm[2:]))[0]
        print("coord %d eps %.0e: FD %.6g  analytic %.6g" % (which, eps, (fp - fm) / (2 * eps), g[which]))
