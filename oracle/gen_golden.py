#!/usr/bin/env python
"""Generate ``tests/golden/*.npz`` from the UNMODIFIED reference and pin the
oracle restatement against it.

Run in the build container only (needs ``/root/reference``):

    python oracle/gen_golden.py

The reference is imported from ``/root/reference`` with a one-line
``numpy.asscalar`` shim (removed in numpy >= 1.23, used at
revrand/utils/base.py:285); nothing in the reference is modified or copied.
Outputs are float64.  The script fails if the oracle disagrees with the
reference beyond summation-order noise.
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

import revrand  # noqa: E402  (the reference)
from revrand import basis_functions as rbf  # noqa: E402
from revrand import likelihoods as rlik  # noqa: E402
from revrand.btypes import Parameter, Positive  # noqa: E402

from oracle import oracle as orc  # noqa: E402
from tests.golden import cases  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
TOL = dict(rtol=1e-9, atol=1e-11)


def check(name, a, b, **kw):
    tol = dict(TOL)
    tol.update(kw)
    if not np.allclose(a, b, **tol):
        err = np.max(np.abs(np.asarray(a) - np.asarray(b)))
        raise SystemExit("ORACLE MISMATCH %s: max abs err %g" % (name, err))


def make_ref_basis(cls, K, d, seed, ard, ls_value, reg=None, apply_ind=None):
    kw = {}
    if apply_ind is not None:
        kw["apply_ind"] = apply_ind
    if reg is not None:
        kw["regularizer"] = Parameter(reg, Positive())
    lsp = Parameter(np.asarray(ls_value, dtype=float) if ard
                    else float(ls_value), Positive())
    return getattr(rbf, cls)(nbases=K, Xdim=d, lenscale=lsp,
                             random_state=seed, **kw)


def ref_block(b):
    """Oracle block description from a reference basis object."""
    cols = getattr(b, "apply_ind", None)
    if isinstance(b, rbf.FastFoodRBF):
        return dict(kind="fastfood", B=b.B, G=b.G, PI=b.PI, S=b.S, cols=cols)
    if isinstance(b, rbf._RandomKernelBasis):
        return dict(kind="trig", W=b.W, cols=cols)
    if isinstance(b, rbf.LinearBasis):
        return dict(kind="linear", onescol=b.onescol, cols=cols)
    if isinstance(b, rbf.BiasBasis):
        return dict(kind="bias", offset=b.offset, cols=cols)
    raise TypeError(b)


def gen_bases():
    out = {}
    for cls in cases.RANDOM_BASES:
        for (d, K, N) in cases.BASIS_SHAPES:
            for seed in cases.BASIS_SEEDS:
                X, ls_iso, ls_ard = cases.basis_case_inputs(d, K, N, seed)
                for ard in (False, True):
                    ls = ls_ard if ard else ls_iso
                    b = make_ref_basis(cls, K, d, seed, ard, ls)
                    key = cases.basis_case_key(cls, d, K, N, seed, ard)
                    Phi = b.transform(X, ls)
                    dPhi = b.grad(X, ls)
                    blk = ref_block(b)
                    blk["lenscale"] = ls
                    check(key + "/Phi", orc.block_features(X, blk), Phi)
                    check(key + "/dPhi", orc.block_grads(X, blk), dPhi)
                    D = Phi.shape[1]
                    probe = cases.probe_matrix(N, D, seed)
                    if isinstance(b, rbf.FastFoodRBF):
                        if not ard:
                            out[key + "/B"] = b.B.astype(np.int8)
                            out[key + "/G"] = b.G
                            out[key + "/PI"] = b.PI.astype(np.int32)
                            out[key + "/S"] = b.S
                    elif not ard:
                        out[key + "/W"] = b.W
                    out[key + "/Phi"] = Phi
                    if dPhi.ndim == 3 and dPhi.size > 20000:
                        out[key + "/dPhi_probe"] = np.einsum(
                            "nj,njp->p", probe, dPhi)
                    else:
                        out[key + "/dPhi"] = dPhi
    np.savez_compressed(os.path.join(OUT, "bases.npz"), **out)
    print("bases.npz: %d arrays" % len(out))


def build_case_basis(case):
    bases, hypers, regs = [], [], []
    for cls, kw in case["blocks"]:
        if cls in ("LinearBasis", "BiasBasis"):
            k2 = {k: v for k, v in kw.items() if k != "reg"}
            b = getattr(rbf, cls)(
                regularizer=Parameter(kw["reg"], Positive()), **k2)
        else:
            ai = kw.get("apply_ind")
            d_eff = len(ai) if ai is not None else case["d"]
            ls = cases.block_lenscale(kw, d_eff)
            b = make_ref_basis(cls, kw["K"], d_eff, kw["seed"], kw["ard"], ls,
                               reg=kw["reg"], apply_ind=ai)
            hypers.append(ls)
        bases.append(b)
        regs.append(kw["reg"])
    basis = bases[0]
    for b in bases[1:]:
        basis = basis + b
    return basis, bases, hypers, regs


def gen_slm():
    out = {}
    for name, case in cases.SLM_CASES.items():
        X, y = cases.slm_case_inputs(case)
        basis, bases, hypers, regs = build_case_basis(case)
        slm = revrand.StandardLinearModel(basis=basis)
        slm.obj_ = -np.inf
        single = len(bases) == 1
        reg_arg = regs[0] if single else list(regs)
        hyp_arg = (hypers[0] if len(hypers) == 1 else list(hypers))
        nelbo, (dvar, dreg, dhyp) = slm._elbo(X, y, case["var"], reg_arg,
                                             hyp_arg)
        blocks = []
        hi = 0
        for b in bases:
            blk = ref_block(b)
            if blk["kind"] in ("trig", "fastfood"):
                blk["lenscale"] = hypers[hi]
                hi += 1
            blocks.append(blk)
        o = orc.slm_elbo(X, y, case["var"], regs, blocks)
        oc = orc.slm_elbo_chunked(X, y, case["var"], regs, blocks, chunk=97)
        dreg_l = [dreg] if single else list(dreg)
        dhyp_l = [dhyp] if len(hypers) == 1 else list(dhyp)
        for tag, oo in (("mono", o), ("chunk", oc)):
            check(name + "/neg_elbo/" + tag, oo["neg_elbo"], nelbo)
            check(name + "/dvar/" + tag, oo["dvar"], dvar)
            for a, b_ in zip(oo["dreg"], dreg_l):
                check(name + "/dreg/" + tag, a, b_)
            for a, b_ in zip(oo["dhyp"], dhyp_l):
                check(name + "/dhyp/" + tag, a, b_, rtol=1e-7, atol=1e-8)
            check(name + "/m/" + tag, oo["m"], slm.weights_, rtol=1e-7,
                  atol=1e-10)
            check(name + "/C/" + tag, oo["C"], slm.covariance_, rtol=1e-7,
                  atol=1e-12)
        out[name + "/neg_elbo"] = np.float64(nelbo)
        out[name + "/dvar"] = np.float64(dvar)
        out[name + "/dreg"] = np.asarray(dreg_l, dtype=float)
        for i, g in enumerate(dhyp_l):
            out[name + "/dhyp%d" % i] = np.asarray(g, dtype=float)
        out[name + "/m"] = slm.weights_
        out[name + "/diagC"] = slm.covariance_.diagonal().copy()
        # predictive moments at seeded query points (slm.py:219-244)
        Xs = np.random.RandomState(5000 + case["seed"]).randn(50, case["d"])
        slm.var_, slm.regularizer_, slm.hypers_ = case["var"], reg_arg, hyp_arg
        Ey, Vy = slm.predict_moments(Xs)
        oEy, oVy = orc.slm_predict_moments(Xs, blocks, o["m"], o["C"],
                                           case["var"])
        check(name + "/Ey", oEy, Ey, rtol=1e-7, atol=1e-10)
        check(name + "/Vy", oVy, Vy, rtol=1e-7, atol=1e-10)
        out[name + "/Ey"] = Ey
        out[name + "/Vy"] = Vy
    np.savez_compressed(os.path.join(OUT, "slm.npz"), **out)
    print("slm.npz: %d arrays" % len(out))


class _InjectedNoise(object):
    """Stands in for ``glm.random_`` so glm.py:300 draws the seeded noise."""

    def __init__(self, eps):
        self.eps = list(eps)

    def randn(self, L, D):
        e = self.eps.pop(0)
        assert e.shape == (L, D)
        return e


def gen_glm():
    out = {}
    sh = cases.GLM_SHAPE
    lik_ids = dict(gaussian=orc.LIK_GAUSSIAN, bernoulli=orc.LIK_BERNOULLI,
                   binomial=orc.LIK_BINOMIAL, poisson_exp=orc.LIK_POISSON_EXP,
                   poisson_softplus=orc.LIK_POISSON_SOFTPLUS)
    for name, spec in cases.GLM_CASES.items():
        inp = cases.glm_case_inputs(name)
        basis = make_ref_basis("RandomRBF", sh["K"], sh["d"], 31, True,
                               inp["ls"], reg=inp["reg"])
        lik = getattr(rlik, spec["lik"])(**spec["lik_kwargs"])
        glm = revrand.GeneralizedLinearModel(likelihood=lik, basis=basis,
                                             K=sh["Kmix"], nsamples=sh["L"])
        glm.B_ = inp["B"]
        glm.D_ = 2 * sh["K"]
        glm._GeneralizedLinearModel__it = -1
        glm.random_ = _InjectedNoise(inp["eps"])
        lpars = spec["lpar"] if spec["lpar"] is not None else []
        largs = (inp["n"],) if spec["largs"] == "n" else ()
        nelbo, (dm, dC, dreg, dlp, dbp) = glm._elbo(
            inp["m"], inp["C"], inp["reg"], lpars, inp["ls"], inp["X"],
            inp["y"], *largs)
        blocks = [dict(kind="trig", W=basis.W, lenscale=inp["ls"], cols=None)]
        o = orc.glm_elbo(inp["m"], inp["C"], [inp["reg"]], lik_ids[name],
                         spec["lpar"], inp["X"], inp["y"], blocks, inp["eps"],
                         inp["B"], lik_arg=inp["n"])
        check(name + "/neg_elbo", o["neg_elbo"], nelbo)
        check(name + "/dm", o["dm"], dm)
        check(name + "/dC", o["dC"], dC)
        check(name + "/dreg", o["dreg"][0], dreg)
        check(name + "/dbp", o["dbpars"][0], dbp, rtol=1e-7, atol=1e-9)
        if spec["lpar"] is not None:
            check(name + "/dlp", o["dlpar"], dlp[0])
            out[name + "/dlpar"] = np.float64(dlp[0])
        out[name + "/neg_elbo"] = np.float64(nelbo)
        out[name + "/dm"] = dm
        out[name + "/dC"] = dC
        out[name + "/dreg"] = np.float64(dreg)
        out[name + "/dbpars"] = np.asarray(dbp, dtype=float)
    np.savez_compressed(os.path.join(OUT, "glm.npz"), **out)
    print("glm.npz: %d arrays" % len(out))


def gen_misc():
    """Known-answer items the reference's own tests pin near this path."""
    out = {}
    # hadamard doctest, revrand/mathfun/linalg.py:202-206
    from revrand.mathfun.linalg import hadamard, solve_posdef
    Y = np.array([[1., 2., 3., 4.], [0., 1., 0., 1.]])
    Hn = hadamard(Y, ordering=False)
    check("hadamard", orc.hadamard_unordered(Y), Hn)
    out["hadamard_in"], out["hadamard_out"] = Y, Hn
    rs = np.random.RandomState(7)
    Y2 = rs.randn(5, 32)
    out["hadamard32_in"], out["hadamard32_out"] = Y2, hadamard(Y2, False)
    # solve_posdef on PD and near-singular (tests/test_mathfun.py:53-80)
    A = rs.randn(6, 6)
    A = A.dot(A.T) + 0.5 * np.eye(6)
    Xr, ld = solve_posdef(A, np.eye(6))
    Xo, ldo = orc.solve_posdef(A, np.eye(6))
    check("solve_posdef", Xo, Xr)
    check("solve_posdef_ld", ldo, ld)
    out["pd_A"], out["pd_inv"], out["pd_logdet"] = A, Xr, np.float64(ld)
    v = rs.randn(6, 2)
    As = v.dot(v.T) + 1e-13 * np.eye(6)
    Xr, ld = solve_posdef(As, np.eye(6))
    Xo, ldo = orc.solve_posdef(As, np.eye(6))
    check("solve_posdef_sing", Xo, Xr, rtol=1e-5, atol=1e-3)
    out["sing_A"], out["sing_inv"], out["sing_logdet"] = As, Xr, np.float64(ld)
    np.savez_compressed(os.path.join(OUT, "misc.npz"), **out)
    print("misc.npz: %d arrays" % len(out))


def config1_fit():
    """BASELINE config 1 end to end through the unmodified reference: the
    learned hyper-parameters quoted in tests/test_gpu_parity.py::
    test_slm_fit_config1_end_to_end (printed, not stored)."""
    from revrand import StandardLinearModel
    from revrand.basis_functions import RandomRBF
    rs = np.random.RandomState(0)
    X = np.sort(rs.uniform(-5, 5, size=(1000, 1)), axis=0)
    y = np.sin(X[:, 0]) + 0.1 * rs.randn(1000)
    slm = StandardLinearModel(basis=RandomRBF(nbases=256, Xdim=1, random_state=1),
                              nstarts=20, maxiter=200, random_state=2)
    slm.fit(X, y)
    print("config1 fit: var_ = %r, regularizer_ = %r, hypers_ = %r, obj_ = %r"
          % (slm.var_, slm.regularizer_, slm.hypers_, slm.obj_))


if __name__ == "__main__":
    gen_bases()
    gen_slm()
    gen_glm()
    gen_misc()
    config1_fit()
    print("oracle pinned against reference; fixtures written to", OUT)
