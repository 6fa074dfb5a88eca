"""Seeded case definitions shared by ``oracle/gen_golden.py`` (which runs the
unmodified reference on them) and the parity tests (which rebuild the same
inputs without needing ``/root/reference``).

Everything here is deterministic in the seeds; only *outputs* of the
reference are stored in ``tests/golden/*.npz``.
"""

import numpy as np

# name, reference class name, constructor kwargs beyond (nbases, Xdim)
RANDOM_BASES = ["RandomRBF", "RandomLaplace", "RandomCauchy",
                "RandomMatern32", "RandomMatern52", "OrthogonalRBF",
                "FastFoodRBF"]

# (d, K, N)
BASIS_SHAPES = [(1, 16, 64), (3, 16, 64), (21, 96, 48)]
BASIS_SEEDS = [0, 1]


def basis_case_inputs(d, K, N, seed):
    rs = np.random.RandomState(1000 + seed)
    X = rs.randn(N, d)
    ls_iso = 0.7 + 0.5 * rs.rand()
    ls_ard = 0.5 + rs.rand(d)
    return X, ls_iso, ls_ard


def probe_matrix(N, D, seed):
    """Seeded (N, D) probe used to contract large gradient tensors."""
    return np.random.RandomState(4000 + seed).randn(N, D)


def basis_case_key(cls, d, K, N, seed, ard):
    return "%s_d%d_K%d_N%d_s%d_%s" % (cls, d, K, N, seed,
                                     "ard" if ard else "iso")


# ---- SLM cases ------------------------------------------------------------
# Each: name -> dict(N, d, blocks spec, var, regs, seed)
# blocks spec entries: (class name, kwargs) with kwargs possibly containing
# "ard": bool, "K": int, "seed": int, "apply_ind": list|None

SLM_CASES = {
    "rbf_iso_d1": dict(N=300, d=1, var=0.3, seed=3, blocks=[
        ("RandomRBF", dict(K=32, seed=11, ard=False, ls=0.8, reg=1.3))]),
    "rbf_iso_d3": dict(N=300, d=3, var=0.05, seed=4, blocks=[
        ("RandomRBF", dict(K=16, seed=12, ard=False, ls=1.5, reg=0.7))]),
    "matern32_ard_d5": dict(N=500, d=5, var=0.02, seed=5, blocks=[
        ("RandomMatern32", dict(K=64, seed=13, ard=True, ls=2.0, reg=1.0))]),
    "cauchy_ard_d21": dict(N=400, d=21, var=0.1, seed=6, blocks=[
        ("RandomCauchy", dict(K=96, seed=14, ard=True, ls=4.0, reg=2.0))]),
    "rbf_plus_linear": dict(N=400, d=5, var=0.07, seed=7, blocks=[
        ("RandomRBF", dict(K=32, seed=15, ard=True, ls=1.2, reg=0.9)),
        ("LinearBasis", dict(onescol=True, reg=3.0))]),
    "two_trig_plus_bias": dict(N=350, d=4, var=0.2, seed=8, blocks=[
        ("RandomMatern52", dict(K=24, seed=16, ard=False, ls=0.9, reg=1.1)),
        ("RandomRBF", dict(K=40, seed=17, ard=True, ls=1.7, reg=0.6,
                           apply_ind=[0, 2])),
        ("BiasBasis", dict(offset=1.0, reg=5.0))]),
    "fastfood_ard_d6": dict(N=300, d=6, var=0.1, seed=9, blocks=[
        ("FastFoodRBF", dict(K=30, seed=18, ard=True, ls=1.4, reg=1.0))]),
    "config1_sine": dict(N=1000, d=1, var=0.01, seed=10, blocks=[
        ("RandomRBF", dict(K=256, seed=1, ard=False, ls=0.6, reg=1.0))]),
}


def slm_case_inputs(case):
    rs = np.random.RandomState(2000 + case["seed"])
    N, d = case["N"], case["d"]
    X = rs.randn(N, d)
    w = rs.randn(d)
    y = np.sin(X.dot(w) / 1.5) + 0.1 * rs.randn(N)
    return X, y


def block_lenscale(kw, d_eff):
    """The lenscale evaluation point for a block (deterministic)."""
    if kw.get("ard"):
        return kw["ls"] * (1.0 + 0.1 * np.arange(d_eff))
    return kw["ls"]


# ---- GLM cases ------------------------------------------------------------

GLM_CASES = {
    "gaussian": dict(lik="Gaussian", lik_kwargs={}, lpar=0.3, largs=None),
    "bernoulli": dict(lik="Bernoulli", lik_kwargs={}, lpar=None, largs=None),
    "binomial": dict(lik="Binomial", lik_kwargs={}, lpar=None, largs="n"),
    "poisson_exp": dict(lik="Poisson", lik_kwargs={"tranfcn": "exp"},
                        lpar=None, largs=None),
    "poisson_softplus": dict(lik="Poisson",
                             lik_kwargs={"tranfcn": "softplus"}, lpar=None,
                             largs=None),
}
GLM_SHAPE = dict(N=2000, M=96, d=4, K=24, Kmix=3, L=7, seed=21)


def glm_case_inputs(name):
    sh = GLM_SHAPE
    rs = np.random.RandomState(3000 + sh["seed"])
    M, d = sh["M"], sh["d"]
    X = rs.randn(M, d)
    f = np.sin(X[:, 0]) + 0.3 * X[:, 1]
    n = None
    if name == "gaussian":
        y = f + 0.2 * rs.randn(M)
    elif name == "bernoulli":
        y = (rs.rand(M) < 1 / (1 + np.exp(-2 * f))).astype(float)
    elif name == "binomial":
        n = rs.randint(3, 9, size=M).astype(float)
        y = rs.binomial(n.astype(int), 1 / (1 + np.exp(-f))).astype(float)
    else:
        y = rs.poisson(np.exp(f)).astype(float)
    D = 2 * sh["K"]
    m = 0.3 * rs.randn(D, sh["Kmix"])
    C = 0.05 + 0.2 * rs.rand(D, sh["Kmix"])
    eps = rs.randn(sh["Kmix"], sh["L"], D)
    ls = 1.0 + 0.2 * np.arange(d)
    return dict(X=X, y=y, n=n, m=m, C=C, eps=eps, ls=ls, reg=1.7,
                B=sh["N"] / M)


# ---- the remaining bases ----------------------------------
# (d, N); centres are seeded (M, d) draws; FastFoodGM uses nbases = 12
BASES2_SHAPES = [(1, 40), (3, 40), (6, 32)]
BASES2_M = 7
BASES2_NBASES = 12
BASES2_ORDER = 3


def bases2_inputs(d, N):
    rs = np.random.RandomState(7000 + d)
    X = rs.randn(N, d)
    C = rs.randn(BASES2_M, d)
    ls_iso = 0.8 + 0.4 * rs.rand()
    ls_ard = 0.6 + rs.rand(d)
    mean = 0.5 * rs.randn(d)
    return X, C, ls_iso, ls_ard, mean


# ---- GLM predictive paths ----------------------------
GLM_PREDICT = dict(N=48, d=3, K=16, Kmix=3, S=150, seed=77, quantile=1.5, percentile=0.9)
GLM_PREDICT_LIKS = ["gaussian", "bernoulli", "binomial", "poisson_exp", "poisson_softplus"]


def glm_predict_inputs(name):
    sh = GLM_PREDICT
    rs = np.random.RandomState(8000 + GLM_PREDICT_LIKS.index(name))
    X = rs.randn(sh["N"], sh["d"])
    D = 2 * sh["K"]
    w = 0.6 * rs.randn(D, sh["Kmix"])
    C = 0.02 + 0.05 * rs.rand(D, sh["Kmix"])
    ls = 0.8 + 0.4 * rs.rand(sh["d"])
    n = rs.randint(3, 12, size=sh["N"]).astype(float)
    return dict(X=X, w=w, C=C, ls=ls, n=n, var=0.35)
