"""CPU oracle for the random-feature log-marginal-likelihood hot path.

TEST INFRASTRUCTURE ONLY.  This module is a float64 numpy restatement of the
reference algorithm (NICTA/revrand @ 4c1881b).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker / the timed CPU
baseline -- never as part of the shipped product path (``revrand_b200/``
must fail loudly without its CUDA extension and never routes through here).

Pinning: every function below is checked against outputs of the *unmodified*
reference imported from ``/root/reference`` by ``oracle/gen_golden.py``; the
resulting vectors are committed under ``tests/golden/`` and re-checked by
``tests/test_oracle_golden.py`` (which does not need ``/root/reference``).
The reference's own tests hold no value-level golden vectors for this path
(SURVEY.md section 8c), so the committed fixtures are the pin.

Each function cites the reference file:line it restates (paths relative to
the reference checkout).
"""

from __future__ import annotations

import math

import numpy as np
from scipy.linalg import LinAlgError, cho_solve, cholesky, svd
from scipy.special import expit, gammaln

CHOLTHRESH = 1e-5  # revrand/mathfun/linalg.py:31


# --------------------------------------------------------------------------
# Feature maps
# --------------------------------------------------------------------------

def _as_lenscale(lenscale, d):
    """Scalar -> length-1 array, as revrand/basis_functions.py:605-613."""
    ls = np.atleast_1d(np.asarray(lenscale, dtype=float))
    if ls.shape not in ((1,), (d,)):
        raise ValueError("lenscale must be scalar or (d,)")
    return ls


def trig_features(X, W, lenscale):
    """[cos(X W/l) | sin(X W/l)] / sqrt(K).

    revrand/basis_functions.py:859-864 (``_RandomKernelBasis.transform``).
    """
    N, d = X.shape
    K = W.shape[1]
    ls = _as_lenscale(lenscale, d)[:, None]
    WX = X @ (W / ls)
    return np.hstack((np.cos(WX), np.sin(WX))) / math.sqrt(K)


def trig_feature_grads(X, W, lenscale):
    """d Phi / d lenscale, shape (N, 2K) for scalar l, (N, 2K, d) for ARD.

    revrand/basis_functions.py:888-901.  Reproduces the reference exactly,
    including the scalar-lenscale behaviour where only input dimension 0
    contributes (the loop at :896 runs once).
    """
    N, d = X.shape
    K = W.shape[1]
    ls = _as_lenscale(lenscale, d)[:, None]
    WX = X @ (W / ls)
    msin, cos = -np.sin(WX), np.cos(WX)
    out = []
    for i, l in enumerate(ls):
        dWX = np.outer(X[:, i], -W[i, :] / l ** 2)
        out.append(np.hstack((dWX * msin, dWX * cos)) / math.sqrt(K))
    return np.dstack(out) if len(ls) != 1 else out[0]


def hadamard_unordered(Y):
    """Unnormalised-by-halves Walsh-Hadamard transform, natural order.

    revrand/mathfun/linalg.py:182-220 with ``ordering=False``: log2(n)
    butterfly stages, each multiplying by [[1,1],[1,-1]]/2, so the overall
    operator is H_n / n.
    """
    nv, n = Y.shape
    steps = int(round(math.log2(n)))
    assert 2 ** steps == n
    H2 = np.array([[1.0, 1.0], [1.0, -1.0]]) / 2.0
    for _ in range(steps):
        Y = np.transpose(Y.reshape(nv, 2, n // 2), (0, 2, 1)).dot(H2)
    return Y.reshape(nv, n)


def fastfood_vx(X, B, G, PI, S):
    """V X for the FastFood projection; revrand/basis_functions.py:1356-1371."""
    N, d0 = X.shape
    k, d2 = B.shape
    Xp = np.zeros((N, d2))
    Xp[:, :d0] = X
    blocks = []
    for b, g, pi, s in zip(B, G, PI, S):
        v = hadamard_unordered(Xp * b[None, :])
        v = v[:, pi] * g[None, :]
        v = hadamard_unordered(v) * s[None, :] * math.sqrt(d2)
        blocks.append(v)
    return np.hstack(blocks)


def fastfood_features(X, B, G, PI, S, lenscale):
    """revrand/basis_functions.py:1285-1289 (``FastFoodRBF.transform``)."""
    ls = _as_lenscale(lenscale, X.shape[1])
    VX = fastfood_vx(X / ls, B, G, PI, S)
    n = B.size
    return np.hstack((np.cos(VX), np.sin(VX))) / math.sqrt(n)


def fastfood_feature_grads(X, B, G, PI, S, lenscale):
    """revrand/basis_functions.py:1311-1329 (``FastFoodRBF.grad``)."""
    d = X.shape[1]
    ls = _as_lenscale(lenscale, d)
    VX = fastfood_vx(X / ls, B, G, PI, S)
    msin, cos = -np.sin(VX), np.cos(VX)
    n = B.size
    out = []
    for i, l in enumerate(ls):
        ind = np.zeros(d)
        ind[i] = 1.0 / l ** 2
        dVX = -fastfood_vx(X * ind, B, G, PI, S)
        out.append(np.hstack((dVX * msin, dVX * cos)) / math.sqrt(n))
    return np.dstack(out) if len(ls) != 1 else out[0]


def polynomial_features(X, order, include_bias=True):
    """[1, X^1 .. X^order] with the powers of one input column adjacent.

    revrand/basis_functions.py:537-566 (``PolynomialBasis.transform``).
    """
    N, d = X.shape
    Phi = (X[:, :, None] ** (np.arange(order) + 1)).reshape(N, d * order)
    if include_bias:
        Phi = np.hstack((np.ones((N, 1)), Phi))
    return Phi


def radial_features(X, C, lenscale):
    """exp(-|| (x - c) / (2 l^2) ||^2): the reference divides by ``2 l^2``
    BEFORE squaring (revrand/basis_functions.py:665-688), reproduced as is."""
    ls = _as_lenscale(lenscale, X.shape[1])
    den = 2.0 * ls ** 2
    diff = X[:, None, :] / den - C[None, :, :] / den
    return np.exp(-(diff ** 2).sum(axis=2))


def radial_feature_grads(X, C, lenscale):
    """revrand/basis_functions.py:690-722: Phi * (x_i - c_i)^2 / l_i^6 per
    lengthscale; a scalar lengthscale only sees input dimension 0."""
    ls = _as_lenscale(lenscale, X.shape[1])
    Phi = radial_features(X, C, lenscale)
    out = []
    for i, l in enumerate(ls):
        ldist = (X[:, [i]] / l ** 3 - C[:, [i]].T / l ** 3) ** 2
        out.append(Phi * ldist)
    return np.dstack(out) if len(ls) != 1 else out[0]


def sigmoidal_features(X, C, lenscale):
    """expit(|| (x - c) / l ||); revrand/basis_functions.py:770-790."""
    ls = _as_lenscale(lenscale, X.shape[1])
    diff = X[:, None, :] / ls - C[None, :, :] / ls
    return expit(np.sqrt((diff ** 2).sum(axis=2)))


def sigmoidal_feature_grads(X, C, lenscale):
    """revrand/basis_functions.py:792-815: -|x_i - c_i| / l_i^2 Phi (1 - Phi)."""
    ls = _as_lenscale(lenscale, X.shape[1])
    Phi = sigmoidal_features(X, C, lenscale)
    out = []
    for i, l in enumerate(ls):
        ldist = np.abs(X[:, [i]] / l ** 2 - C[:, [i]].T / l ** 2)
        out.append(-ldist * Phi * (1 - Phi))
    return np.dstack(out) if len(ls) != 1 else out[0]


def fastfood_gm_features(X, B, G, PI, S, mean, lenscale):
    """One Gaussian spectral-mixture component: [cos(VX + Xm) | sin(VX + Xm) |
    cos(VX - Xm) | sin(VX - Xm)] / sqrt(2 n).

    revrand/basis_functions.py:1458-1472 (``FastFoodGM.transform``).
    """
    VX = fastfood_vx(X / lenscale, B, G, PI, S)
    mX = X.dot(mean)[:, None]
    n = B.size
    return np.hstack((np.cos(VX + mX), np.sin(VX + mX),
                      np.cos(VX - mX), np.sin(VX - mX))) / math.sqrt(2 * n)


def fastfood_gm_feature_grads(X, B, G, PI, S, mean, lenscale):
    """(d Phi / d mean, d Phi / d lenscale), each (N, 4n, d) (2-D for d = 1).

    revrand/basis_functions.py:1474-1527 (``FastFoodGM.grad``).
    """
    d = X.shape[1]
    VX = fastfood_vx(X / lenscale, B, G, PI, S)
    mX = X.dot(mean)[:, None]
    sp, sm = -np.sin(VX + mX), -np.sin(VX - mX)
    cp, cm = np.cos(VX + mX), np.cos(VX - mX)
    n = B.size
    dmean, dlen = [], []
    for i, l in enumerate(np.atleast_1d(lenscale)):
        dmX = X[:, [i]]
        dmean.append(np.hstack((dmX * sp, dmX * cp, -dmX * sm, -dmX * cm))
                     / math.sqrt(2 * n))
        ind = np.zeros(d)
        ind[i] = 1.0 / l ** 2
        dVX = -fastfood_vx(X * ind, B, G, PI, S)
        dlen.append(np.hstack((dVX * sp, dVX * cp, dVX * sm, dVX * cm))
                    / math.sqrt(2 * n))
    if d != 1:
        return np.dstack(dmean), np.dstack(dlen)
    return dmean[0], dlen[0]


def linear_features(X, onescol=True):
    """revrand/basis_functions.py:468-485."""
    return np.hstack((np.ones((len(X), 1)), X)) if onescol else X


def bias_features(X, offset=1.0):
    """revrand/basis_functions.py:415-432."""
    return np.ones((len(X), 1)) * offset


# A *block* is a dict describing one basis of a concatenation:
#   {"kind": "trig", "W": (d,K), "lenscale": scalar|(d,), "cols": idx|None}
#   {"kind": "fastfood", "B","G","PI","S", "lenscale", "cols"}
#   {"kind": "linear", "onescol": bool, "cols"}
#   {"kind": "bias", "offset": float, "cols"}
# "cols" mirrors ``apply_ind`` (revrand/basis_functions.py:70-105).

def _slice(X, blk):
    cols = blk.get("cols")
    return X if cols is None else X[:, cols]


def block_features(X, blk):
    Xs = _slice(X, blk)
    kind = blk["kind"]
    if kind == "trig":
        return trig_features(Xs, blk["W"], blk["lenscale"])
    if kind == "fastfood":
        return fastfood_features(Xs, blk["B"], blk["G"], blk["PI"], blk["S"],
                                 blk["lenscale"])
    if kind == "linear":
        return linear_features(Xs, blk.get("onescol", True))
    if kind == "bias":
        return bias_features(Xs, blk.get("offset", 1.0))
    raise ValueError(kind)


def block_grads(X, blk):
    Xs = _slice(X, blk)
    kind = blk["kind"]
    if kind == "trig":
        return trig_feature_grads(Xs, blk["W"], blk["lenscale"])
    if kind == "fastfood":
        return fastfood_feature_grads(Xs, blk["B"], blk["G"], blk["PI"],
                                      blk["S"], blk["lenscale"])
    return None  # parameter-free: revrand/basis_functions.py:252-274 -> []


def concat_features(X, blocks):
    """``BasisCat.transform``; revrand/basis_functions.py:1599-1627."""
    return np.hstack([block_features(X, b) for b in blocks])


def concat_grads(X, blocks):
    """``BasisCat.grad``: one zero-padded array per parameterised base.

    revrand/basis_functions.py:1629-1677.
    """
    dims = [block_features(X[:1], b).shape[1] for b in blocks]
    ends = np.cumsum([0] + dims)
    D = int(ends[-1])
    N = X.shape[0]
    out = []
    for i, b in enumerate(blocks):
        g = block_grads(X, b)
        if g is None:
            continue
        shape = (N, D) if g.ndim < 3 else (N, D, g.shape[2])
        full = np.zeros(shape)
        full[:, ends[i]:ends[i + 1]] = g
        out.append(full)
    return out


def regularizer_diagonal(X, blocks, regs):
    """diag(Lambda) and per-base column slices.

    revrand/basis_functions.py:307-340 and :1712-1748.
    """
    dims = [block_features(X[:1], b).shape[1] for b in blocks]
    ends = np.cumsum([0] + dims)
    diag = np.concatenate([np.full(n, float(r)) for n, r in zip(dims, regs)])
    slices = [slice(int(a), int(b)) for a, b in zip(ends[:-1], ends[1:])]
    return diag, slices


# --------------------------------------------------------------------------
# Linear algebra
# --------------------------------------------------------------------------

def solve_posdef(A, b):
    """Cholesky solve with SVD fallback; revrand/mathfun/linalg.py:84-125.

    Returns (A^-1 b, logdet A).  Falls back to the clamped-SVD solve of
    :128-179 when the factorisation fails or any diagonal of the factor is
    below CHOLTHRESH.
    """
    try:
        L = cholesky(A, lower=False)
        if np.any(L.diagonal() < CHOLTHRESH):
            raise LinAlgError("unstable cholesky")
        X = cho_solve((L, False), b)
        logdet = 2.0 * np.sum(np.log(L.diagonal()))
    except LinAlgError:
        U, s, V = svd(A)
        sc = np.maximum(s, 1e-15)
        ss = 1.0 / np.sqrt(sc)
        X = (U * ss[None, :]).dot((ss[:, None] * V)).dot(b)
        logdet = np.sum(np.log(s))
    return X, logdet


# --------------------------------------------------------------------------
# Standard linear model: ELBO (= exact log marginal likelihood) + gradients
# --------------------------------------------------------------------------

def _apply_grad(fun, g):
    """revrand/basis_functions.py:109-152 restricted to ndarray inputs."""
    if g.ndim <= 2:
        return fun(g)
    return np.array([fun(g[:, :, i]) for i in range(g.shape[2])])


def slm_elbo(X, y, var, regs, blocks):
    """Monolithic restatement of ``StandardLinearModel._elbo``.

    revrand/slm.py:142-199.  ``regs`` is a list with one regulariser per
    block.  Returns a dict with the negative ELBO, its gradients in the
    reference's layout (-dvar, dreg per block, dhyp per parameterised block)
    and the posterior (m, C).
    """
    Phi = concat_features(X, blocks)
    G = Phi.T.dot(Phi)
    N, D = Phi.shape
    Ld, slices = regularizer_diagonal(X, blocks, regs)
    iL = 1.0 / Ld
    iC = np.diag(iL) + G / var
    C, logdetiC = solve_posdef(iC, np.eye(D))
    m = C.dot(Phi.T.dot(y)) / var
    TrGC = (G * C).sum()
    Err = y - Phi.dot(m)
    sqErr = (Err ** 2).sum()
    ELBO = -0.5 * (N * np.log(2 * np.pi * var) + sqErr / var + TrGC / var
                   + ((m ** 2 + C.diagonal()) * iL).sum() + logdetiC
                   + np.log(Ld).sum() - D)
    dvar = 0.5 * (-N + (sqErr + TrGC) / var) / var
    dreg = [-0.5 * (((m[s] ** 2 + C[s, s].diagonal()) * iL[s] ** 2).sum()
                    - iL[s].sum()) for s in slices]

    def dhyp(dPhi):
        return -(m.T.dot(Err.dot(dPhi)) - (dPhi.T.dot(Phi) * C).sum()) / var

    dhyps = [_apply_grad(dhyp, g) for g in concat_grads(X, blocks)]
    return {"neg_elbo": -ELBO, "dvar": -dvar, "dreg": dreg, "dhyp": dhyps,
            "m": m, "C": C, "G": G, "Phiy": Phi.T.dot(y), "sqErr": sqErr,
            "TrGC": TrGC, "logdetiC": logdetiC}


def slm_elbo_chunked(X, y, var, regs, blocks, chunk=20000, grads=True,
                     budget=None):
    """Row-chunked sufficient-statistics restatement of ``_elbo``.

    Pass 1 accumulates G, Phi^T y over row chunks (slm.py:145-146,157);
    the solve follows slm.py:150-157; pass 2 recomputes each chunk to get
    the residuals (slm.py:161-162) and the hyper-gradients (slm.py:193-197)
    chunk by chunk, which are sums over rows.  Equal to :func:`slm_elbo` up
    to summation order; used where Phi does not fit in memory and as the
    timed CPU baseline in ``bench.py``.

    ``budget=(seconds_pass1, seconds_pass2)`` is for TIMING runs only: each
    pass stops taking new chunks once its budget is spent, and the returned
    ``"timing"`` entry says how many rows each pass covered and how long it
    and the solve took (the numerical results then describe the rows seen).
    """
    import time
    N = X.shape[0]
    D = concat_features(X[:1], blocks).shape[1]
    G = np.zeros((D, D))
    p = np.zeros(D)
    t0 = time.perf_counter()
    rows1 = 0
    for s in range(0, N, chunk):
        Phi = concat_features(X[s:s + chunk], blocks)
        G += Phi.T.dot(Phi)
        p += Phi.T.dot(y[s:s + chunk])
        rows1 = min(N, s + chunk)
        if budget is not None and time.perf_counter() - t0 > budget[0]:
            break
    t_pass1 = time.perf_counter() - t0
    t0 = time.perf_counter()
    Ld, slices = regularizer_diagonal(X[:1], blocks, regs)
    iL = 1.0 / Ld
    iC = np.diag(iL) + G / var
    C, logdetiC = solve_posdef(iC, np.eye(D))
    m = C.dot(p) / var
    TrGC = (G * C).sum()
    t_solve = time.perf_counter() - t0
    sqErr = 0.0
    dh = None
    t0 = time.perf_counter()
    rows2 = 0
    for s in range(0, rows1, chunk):
        Xc, yc = X[s:s + chunk], y[s:s + chunk]
        Phi = concat_features(Xc, blocks)
        Err = yc - Phi.dot(m)
        sqErr += (Err ** 2).sum()
        if grads:
            def dhyp(dPhi):
                return -(m.T.dot(Err.dot(dPhi))
                         - (dPhi.T.dot(Phi) * C).sum()) / var
            part = [_apply_grad(dhyp, g) for g in concat_grads(Xc, blocks)]
            dh = part if dh is None else [a + b for a, b in zip(dh, part)]
        rows2 = min(rows1, s + chunk)
        if budget is not None and time.perf_counter() - t0 > budget[1]:
            break
    t_pass2 = time.perf_counter() - t0
    Nn = rows1
    ELBO = -0.5 * (Nn * np.log(2 * np.pi * var) + sqErr / var + TrGC / var
                   + ((m ** 2 + C.diagonal()) * iL).sum() + logdetiC
                   + np.log(Ld).sum() - D)
    dvar = 0.5 * (-Nn + (sqErr + TrGC) / var) / var
    dreg = [-0.5 * (((m[s] ** 2 + C[s, s].diagonal()) * iL[s] ** 2).sum()
                    - iL[s].sum()) for s in slices]
    return {"neg_elbo": -ELBO, "dvar": -dvar, "dreg": dreg,
            "dhyp": dh if dh is not None else [], "m": m, "C": C, "G": G,
            "Phiy": p, "sqErr": sqErr, "TrGC": TrGC, "logdetiC": logdetiC,
            "timing": {"rows_pass1": rows1, "rows_pass2": rows2,
                       "t_pass1": t_pass1, "t_solve": t_solve,
                       "t_pass2": t_pass2}}


def slm_predict_moments(Xs, blocks, m, C, var):
    """revrand/slm.py:239-244."""
    Phi = concat_features(Xs, blocks)
    return Phi.dot(m), (Phi.dot(C) * Phi).sum(axis=1) + var


# --------------------------------------------------------------------------
# Likelihoods (revrand/likelihoods.py) and special functions
# --------------------------------------------------------------------------

def softplus(f):
    """log(1+exp(f)), stable; revrand/mathfun/special.py:91-124."""
    return np.logaddexp(0.0, f)


def safesoftplus(f):
    """revrand/mathfun/special.py:138-142 (floor 1e-100)."""
    return np.maximum(softplus(f), 1e-100)


def logsumexp0(X):
    """Column-wise logsumexp; revrand/mathfun/special.py:22-51 (axis=0)."""
    mx = X.max(axis=0)
    return np.log(np.exp(X - mx[None, :]).sum(axis=0)) + mx


LIK_GAUSSIAN, LIK_BERNOULLI, LIK_BINOMIAL, LIK_POISSON_EXP, \
    LIK_POISSON_SOFTPLUS = range(5)


def lik_loglike(lik, y, f, arg=None):
    """loglike; revrand/likelihoods.py:46-66,171-192,298-322,456-484."""
    y, f = np.broadcast_arrays(y, f)
    if lik == LIK_GAUSSIAN:
        return -0.5 * (np.log(2 * np.pi * arg) + (y - f) ** 2 / arg)
    if lik == LIK_BERNOULLI:
        return y * f - softplus(f)
    if lik == LIK_BINOMIAL:
        n = np.broadcast_to(arg, f.shape)
        return (gammaln(n + 1) - gammaln(y + 1) - gammaln(n - y + 1)
                + y * f - n * softplus(f))
    if lik == LIK_POISSON_EXP:
        return y * f - np.exp(f) - gammaln(y + 1)
    if lik == LIK_POISSON_SOFTPLUS:
        g = softplus(f)
        return y * np.log(g) - g - gammaln(y + 1)
    raise ValueError(lik)


def lik_df(lik, y, f, arg=None):
    """d loglike / d f; revrand/likelihoods.py:84-104,211-233,346-368,501-521."""
    y, f = np.broadcast_arrays(y, f)
    if lik == LIK_GAUSSIAN:
        return (y - f) / arg
    if lik == LIK_BERNOULLI:
        return y - expit(f)
    if lik == LIK_BINOMIAL:
        return y - expit(f) * np.broadcast_to(arg, f.shape)
    if lik == LIK_POISSON_EXP:
        return y - np.exp(f)
    if lik == LIK_POISSON_SOFTPLUS:
        return expit(f) * (y / safesoftplus(f) - 1)
    raise ValueError(lik)


def lik_dp(lik, y, f, arg=None):
    """d loglike / d likelihood-parameter (Gaussian var only).

    revrand/likelihoods.py:370-396; parameter-free likelihoods return []
    (:106-127).
    """
    if lik == LIK_GAUSSIAN:
        y, f = np.broadcast_arrays(y, f)
        iv = 1.0 / arg
        return 0.5 * (((y - f) * iv) ** 2 - iv)
    return None


def lik_Ey(lik, f, arg=None):
    """revrand/likelihoods.py:68-82,194-209,324-344,486-499."""
    if lik == LIK_GAUSSIAN:
        return f
    if lik == LIK_BERNOULLI:
        return expit(f)
    if lik == LIK_BINOMIAL:
        return expit(f) * arg
    if lik == LIK_POISSON_EXP:
        return np.exp(f)
    if lik == LIK_POISSON_SOFTPLUS:
        return softplus(f)
    raise ValueError(lik)


# --------------------------------------------------------------------------
# Generalised linear model: AEVB ELBO step
# --------------------------------------------------------------------------

def lik_cdf(lik, q, f, lik_param=None, arg=None):
    """Likelihood CDFs: revrand/likelihoods.py:129-146 (Bernoulli), :235-254
    (Binomial), :398-419 (Gaussian), :523-541 (Poisson)."""
    from scipy.stats import bernoulli, binom, norm, poisson
    if lik == LIK_GAUSSIAN:
        return norm.cdf(q, loc=f, scale=np.sqrt(lik_param))
    if lik == LIK_BERNOULLI:
        return bernoulli.cdf(q, expit(f))
    if lik == LIK_BINOMIAL:
        return binom.cdf(q, n=arg, p=expit(f))
    if lik == LIK_POISSON_EXP:
        return poisson.cdf(q, mu=np.exp(f))
    if lik == LIK_POISSON_SOFTPLUS:
        return poisson.cdf(q, mu=softplus(f))
    raise ValueError(lik)


def glm_predict_moments(F, lik, arg=None):
    """Ey, Vy from latent draws F (N, S); revrand/glm.py:404-418."""
    a = None if arg is None else np.asarray(arg)[:, None]
    ys = lik_Ey(lik, F, a)
    Ey = ys.mean(axis=1)
    return Ey, ((ys - Ey[:, None]) ** 2).mean(axis=1)


def glm_predict_cdf(F, lik, quantile, lik_param=None, arg=None):
    """(mean, min, max) over the draws of cdf(quantile | f); revrand/glm.py:499-516."""
    a = None if arg is None else np.asarray(arg)[:, None]
    ps = lik_cdf(lik, quantile, F, lik_param, a)
    return ps.mean(axis=1), ps.min(axis=1), ps.max(axis=1)


def glm_predict_interval(F, lik, percentile, lik_param=None, arg=None):
    """Per-row brentq roots of the Monte-Carlo CDF; revrand/glm.py:669-694
    (``_rootfinding``) applied to every row as :546-570 does."""
    from scipy.optimize import brentq
    N = F.shape[0]
    lp = (1 - percentile) / 2
    up = 1 - lp
    ql, qu = np.empty(N), np.empty(N)
    for n in range(N):
        fn = F[n]
        an = None if arg is None else arg[n]

        def gap(q, pct):
            return lik_cdf(lik, q, fn, lik_param, an).mean() - pct
        Eyn = lik_Ey(lik, fn, an).mean()
        lb, ub = -1000 * max(Eyn, 1), 1000 * max(Eyn, 1)
        for out, pct in ((ql, lp), (qu, up)):
            try:
                out[n] = brentq(gap, a=lb, b=ub, args=(pct,))
            except ValueError:
                out[n] = np.nan
    return ql, qu


def qmatrix(m, C):
    """log N(m_i; m_j, diag(C_i + C_j)); revrand/glm.py:697-712."""
    K = m.shape[1]
    out = np.empty((K, K))
    for i in range(K):
        for j in range(K):
            v = C[:, i] + C[:, j]
            out[i, j] = -0.5 * (np.log(2 * np.pi * v)
                                + (m[:, i] - m[:, j]) ** 2 / v).sum()
    return out


def glm_elbo(m, C, regs, lik, lik_param, X, y, blocks, eps, B, lik_arg=None,
             calc_ll=True):
    """Restatement of ``GeneralizedLinearModel._elbo`` + ``_reparam_k``.

    revrand/glm.py:205-322.  ``eps`` (K_mix, L, D) is the reparameterisation
    noise the reference would draw at glm.py:300 for k = 0..K_mix-1 in order;
    ``B`` is the batch magnification N / batch_size (glm.py:160).
    ``lik_param`` is the learnable likelihood parameter (Gaussian var) or
    None; ``lik_arg`` a fixed per-row argument (Binomial n) or None.
    """
    D, K = m.shape
    L = eps.shape[1]
    arg = lik_param if lik_param is not None else lik_arg
    Phi = concat_features(X, blocks)
    Ld, slices = regularizer_diagonal(X[:1], blocks, regs)
    iL = 1.0 / Ld[:, None]
    logNkl = qmatrix(m, C)
    logzk = logsumexp0(logNkl)
    dm = np.empty_like(m)
    dC = np.empty_like(C)
    Ell = np.empty(K)
    dlp = 0.0
    EdPhi = np.zeros_like(Phi)
    for k in range(K):
        e = eps[k]
        Sk = np.sqrt(C[:, k])
        ws = m[:, k] + Sk * e
        fs = ws.dot(Phi.T)
        dfs = lik_df(lik, y, fs, arg)
        Edws = dfs.dot(Phi)
        Edm = Edws.sum(axis=0) / L
        EdC = (Edws * e / Sk).sum(axis=0) / L
        EdPhi += dfs.T.dot(ws) / L / K
        dp = lik_dp(lik, y, fs, arg)
        if dp is not None:
            dlp -= dp.sum() / L / K
        Ell[k] = lik_loglike(lik, y, fs, arg).sum() / L if calc_ll else np.inf
        Nkl_zk = np.exp(logNkl[:, k] - logzk[k])
        Nkl_zl = np.exp(logNkl[:, k] - logzk)
        alpha = Nkl_zk + Nkl_zl
        mkmj = m[:, k][:, None] - m
        iCkCj = 1.0 / (C[:, k][:, None] + C)
        dm[:, k] = (B * Edm - m[:, k] / Ld + (iCkCj * mkmj).dot(alpha)) / K
        dC[:, k] = (B * EdC - 1.0 / Ld
                    + (iCkCj - (mkmj * iCkCj) ** 2).dot(alpha)) / (2 * K)
    dreg = [-0.5 * (((m[s] ** 2 + C[s]) * iL[s] ** 2).sum() / K - iL[s].sum())
            for s in slices]
    dbp = [_apply_grad(lambda dPhi: -(EdPhi * dPhi).sum(), g)
           for g in concat_grads(X, blocks)]
    ELBO = -np.inf
    if calc_ll:
        ELBO = (Ell.sum() * B - 0.5 * D * K * np.log(2 * np.pi)
                - 0.5 * K * np.log(Ld).sum()
                - 0.5 * ((m ** 2 + C) * iL).sum()
                - logzk.sum() + np.log(K)) / K
    return {"neg_elbo": -ELBO, "dm": -dm, "dC": -dC, "dreg": dreg,
            "dlpar": dlp if lik_param is not None else None, "dbpars": dbp,
            "EdPhi": EdPhi}
