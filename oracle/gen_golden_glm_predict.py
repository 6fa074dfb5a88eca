#!/usr/bin/env python
"""Generate ``tests/golden/glm_predict.npz`` from the UNMODIFIED reference:
``GeneralizedLinearModel.predict_moments / predict_cdf / predict_interval``
(revrand/glm.py:349-418, 468-570, 669-694) for a hand-set variational posterior
and seeded weight draws, for every likelihood; pins the oracle restatements
(``orc.glm_predict_*``).  Build container only (needs /root/reference).

    python oracle/gen_golden_glm_predict.py
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("REVRAND_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
if not hasattr(np, "asscalar"):
    np.asscalar = lambda a: a.item()  # noqa: E731

import revrand  # noqa: E402
from revrand import basis_functions as rbf  # noqa: E402
from revrand import likelihoods as rlik  # noqa: E402
from revrand.btypes import Parameter, Positive  # noqa: E402

from oracle import oracle as orc  # noqa: E402
from tests.golden import cases  # noqa: E402

LIK = dict(gaussian=(rlik.Gaussian, orc.LIK_GAUSSIAN), bernoulli=(rlik.Bernoulli, orc.LIK_BERNOULLI),
           binomial=(rlik.Binomial, orc.LIK_BINOMIAL),
           poisson_exp=(lambda: rlik.Poisson('exp'), orc.LIK_POISSON_EXP),
           poisson_softplus=(lambda: rlik.Poisson('softplus'), orc.LIK_POISSON_SOFTPLUS))


def check(name, a, b, **kw):
    tol = dict(rtol=1e-9, atol=1e-9)
    tol.update(kw)
    if not np.allclose(a, b, equal_nan=True, **tol):
        raise SystemExit("ORACLE MISMATCH %s: max abs err %g"
                         % (name, np.nanmax(np.abs(np.asarray(a) - np.asarray(b)))))


def main():
    sh = cases.GLM_PREDICT
    out = {}
    for name in cases.GLM_PREDICT_LIKS:
        inp = cases.glm_predict_inputs(name)
        mk, lid = LIK[name]
        basis = rbf.RandomRBF(nbases=sh["K"], Xdim=sh["d"], random_state=31,
                              lenscale=Parameter(inp["ls"], Positive()))
        glm = revrand.GeneralizedLinearModel(likelihood=mk(), basis=basis, K=sh["Kmix"])
        glm.weights_, glm.covariance_ = inp["w"], inp["C"]
        glm.basis_hypers_, glm.regularizer_ = inp["ls"], 1.0
        glm.like_hypers_ = inp["var"] if name == "gaussian" else []
        largs = (inp["n"],) if name == "binomial" else ()
        glm.random_ = np.random.RandomState(sh["seed"])
        Ey, Vy = glm.predict_moments(inp["X"], nsamples=sh["S"], likelihood_args=largs)
        p, pmin, pmax = glm.predict_cdf(inp["X"], sh["quantile"], nsamples=sh["S"],
                                        likelihood_args=largs)
        ql, qu = glm.predict_interval(inp["X"], sh["percentile"], nsamples=sh["S"],
                                      likelihood_args=largs, multiproc=False)
        # the oracle on the same draws
        rs = np.random.RandomState(sh["seed"])
        Phi = orc.trig_features(inp["X"], basis.W, inp["ls"])
        D = 2 * sh["K"]

        def draws():
            k = rs.randint(0, sh["Kmix"], size=(sh["S"],))
            w = inp["w"][:, k] + rs.randn(D, sh["S"]) * np.sqrt(inp["C"][:, k])
            return Phi.dot(w)
        arg = inp["n"] if name == "binomial" else None
        par = inp["var"] if name == "gaussian" else None
        oEy, oVy = orc.glm_predict_moments(draws(), lid, arg)
        check(name + "/Ey", oEy, Ey)
        check(name + "/Vy", oVy, Vy)
        op = orc.glm_predict_cdf(draws(), lid, sh["quantile"], par, arg)
        for a, b, nm in zip(op, (p, pmin, pmax), ("p", "pmin", "pmax")):
            check(name + "/" + nm, a, b)
        oql, oqu = orc.glm_predict_interval(draws(), lid, sh["percentile"], par, arg)
        check(name + "/ql", oql, ql, atol=1e-6)
        check(name + "/qu", oqu, qu, atol=1e-6)
        for key, val in (("Ey", Ey), ("Vy", Vy), ("p", p), ("pmin", pmin), ("pmax", pmax),
                         ("ql", ql), ("qu", qu)):
            out[name + "/" + key] = val
    path = os.path.join(ROOT, "tests", "golden", "glm_predict.npz")
    np.savez_compressed(path, **out)
    print("glm_predict.npz: %d arrays -> %s" % (len(out), path))
    for name in cases.GLM_PREDICT_LIKS:
        print(name, "ql[:4]", out[name + "/ql"][:4], "qu[:4]", out[name + "/qu"][:4])


if __name__ == "__main__":
    main()
