#!/usr/bin/env python
"""Generate ``tests/golden/bases2.npz`` from the UNMODIFIED reference: feature
maps and gradients of PolynomialBasis, RadialBasis, SigmoidalBasis and
FastFoodGM (revrand/basis_functions.py:496-576, 616-815, 1386-1562), and pin
the oracle restatements of them.  Build container only (needs /root/reference).

    python oracle/gen_golden_bases2.py
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

from revrand import basis_functions as rbf  # noqa: E402
from revrand.btypes import Bound, Parameter, Positive  # noqa: E402

from oracle import oracle as orc  # noqa: E402
from tests.golden import cases  # noqa: E402


def check(name, a, b):
    if not np.allclose(a, b, rtol=1e-9, atol=1e-11):
        raise SystemExit("ORACLE MISMATCH %s: max abs err %g"
                         % (name, np.max(np.abs(np.asarray(a) - np.asarray(b)))))


def main():
    out = {}
    for (d, N) in cases.BASES2_SHAPES:
        X, C, ls_iso, ls_ard, mean = cases.bases2_inputs(d, N)
        tag = "d%d" % d
        b = rbf.PolynomialBasis(order=cases.BASES2_ORDER, include_bias=True)
        Phi = b.transform(X)
        check(tag + "/poly", orc.polynomial_features(X, cases.BASES2_ORDER), Phi)
        out[tag + "/poly/Phi"] = Phi
        for ard in (False, True):
            ls = ls_ard if ard else ls_iso
            lsp = Parameter(np.asarray(ls, dtype=float) if ard else float(ls), Positive())
            key = tag + ("/ard" if ard else "/iso")
            for name, cls, f, g in (("radial", rbf.RadialBasis, orc.radial_features,
                                     orc.radial_feature_grads),
                                    ("sigmoid", rbf.SigmoidalBasis, orc.sigmoidal_features,
                                     orc.sigmoidal_feature_grads)):
                rb = cls(centres=C, lenscale=lsp)
                Phi, dPhi = rb.transform(X, ls), rb.grad(X, ls)
                check(key + "/" + name, f(X, C, ls), Phi)
                check(key + "/" + name + "/grad", g(X, C, ls), dPhi)
                out[key + "/" + name + "/Phi"] = Phi
                out[key + "/" + name + "/dPhi"] = dPhi
        gm = rbf.FastFoodGM(nbases=cases.BASES2_NBASES, Xdim=d, random_state=3,
                            mean=Parameter(mean.copy(), Bound()),
                            lenscale=Parameter(ls_ard.copy(), Positive()))
        Phi = gm.transform(X, mean, ls_ard)
        dm, dl = gm.grad(X, mean, ls_ard)
        check(tag + "/gm", orc.fastfood_gm_features(X, gm.B, gm.G, gm.PI, gm.S, mean, ls_ard), Phi)
        om, ol = orc.fastfood_gm_feature_grads(X, gm.B, gm.G, gm.PI, gm.S, mean, ls_ard)
        check(tag + "/gm/dmean", om, dm)
        check(tag + "/gm/dlen", ol, dl)
        out[tag + "/gm/B"] = gm.B.astype(np.int8)
        out[tag + "/gm/G"] = gm.G
        out[tag + "/gm/PI"] = gm.PI.astype(np.int32)
        out[tag + "/gm/S"] = gm.S
        out[tag + "/gm/Phi"] = Phi
        out[tag + "/gm/dmean"] = dm
        out[tag + "/gm/dlen"] = dl
    path = os.path.join(ROOT, "tests", "golden", "bases2.npz")
    np.savez_compressed(path, **out)
    print("bases2.npz: %d arrays -> %s" % (len(out), path))


if __name__ == "__main__":
    main()
