"""Shared builders for the parity tests: the same seeded cases as
tests/golden/cases.py, instantiated with revrand_b200 objects and as oracle
block descriptions."""

import numpy as np

from revrand_b200 import Parameter, Positive
from revrand_b200 import basis_functions as bf
from tests.golden import cases


def make_basis(cls, K, d, seed, ard, ls_value, reg=None, apply_ind=None):
    kw = {}
    if apply_ind is not None:
        kw["apply_ind"] = apply_ind
    if reg is not None:
        kw["regularizer"] = Parameter(reg, Positive())
    lsp = Parameter(np.asarray(ls_value, dtype=float) if ard
                    else float(ls_value), Positive())
    return getattr(bf, cls)(nbases=K, Xdim=d, lenscale=lsp, random_state=seed,
                            **kw)


def build_case_basis(case):
    """(basis, bases, hypers, regs) for an SLM golden case."""
    bases, hypers, regs = [], [], []
    for cls, kw in case["blocks"]:
        if cls in ("LinearBasis", "BiasBasis"):
            k2 = {k: v for k, v in kw.items() if k != "reg"}
            b = getattr(bf, cls)(regularizer=Parameter(kw["reg"], Positive()),
                                 **k2)
        else:
            ai = kw.get("apply_ind")
            d_eff = len(ai) if ai is not None else case["d"]
            ls = cases.block_lenscale(kw, d_eff)
            b = make_basis(cls, kw["K"], d_eff, kw["seed"], kw["ard"], ls,
                           reg=kw["reg"], apply_ind=ai)
            hypers.append(ls)
        bases.append(b)
        regs.append(kw["reg"])
    basis = bases[0]
    for b in bases[1:]:
        basis = basis + b
    return basis, bases, hypers, regs


def oracle_blocks(bases, hypers):
    """Oracle block dicts for revrand_b200 basis objects."""
    out, hi = [], 0
    for b in bases:
        cols = getattr(b, "apply_ind", None)
        if isinstance(b, bf.FastFoodRBF):
            blk = dict(kind="fastfood", B=b.B, G=b.G, PI=b.PI, S=b.S, cols=cols,
                       lenscale=hypers[hi])
            hi += 1
        elif isinstance(b, bf._RandomKernelBasis):
            blk = dict(kind="trig", W=b.W, cols=cols, lenscale=hypers[hi])
            hi += 1
        elif isinstance(b, bf.LinearBasis):
            blk = dict(kind="linear", onescol=b.onescol, cols=cols)
        elif isinstance(b, bf.BiasBasis):
            blk = dict(kind="bias", offset=b.offset, cols=cols)
        else:
            raise TypeError(b)
        out.append(blk)
    return out


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def max_phase(basis, X, ls):
    """max |x . W / l| over the batch (radians) for a random-frequency basis;
    0 for bases without an explicit frequency matrix (FastFood: Gaussian-like
    phases of O(10) rad).  Used to scale fp32 phase tolerances."""
    W = getattr(basis, "W", None)
    if W is None:
        return 0.0
    lsf = np.broadcast_to(np.atleast_1d(np.asarray(ls, dtype=float)), (W.shape[0],))
    return float(np.max(np.abs(np.asarray(X, dtype=float).dot(W / lsf[:, None]))))
