"""CPU: host-side logic of the drop-in boundary (no compute calls)."""

import ctypes
import os
import pickle
import re

import numpy as np
import pytest
from scipy.optimize import minimize
from scipy.stats import gamma
from sklearn.base import clone

import revrand_b200 as rr
from revrand_b200 import Bound, Parameter, Positive, _cabi, _engine
from revrand_b200 import basis_functions as bf
from revrand_b200 import likelihoods as lk
from revrand_b200.optimize import (Adam, AdaDelta, AdaGrad, Momentum,
                                   SGDUpdater, Layout, flatten_values,
                                   gen_batch, logtrick_minimizer, logtrick_sgd,
                                   sgd, structured_minimizer, structured_sgd)
from oracle import oracle as orc
from tests.golden import cases
from tests import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "revrand_b200.h")).read()
    declared = set(re.findall(r"\b(rr_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"rr_status", "rr_plan", "rr_likelihood"}
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)
    lib = _cabi.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.rr_version() >= 100
    assert ctypes.sizeof(_cabi.RRPlan) == 16 + 10 * ctypes.sizeof(ctypes.c_void_p)
    # shape queries need no GPU
    assert lib.rr_tcgen05_supported(21, 2048, 0, 4096) == 1
    assert lib.rr_tcgen05_supported(21, 2048, 22, 4118) == 1   # affine columns ride along
    assert lib.rr_tcgen05_supported(21, 2048, 22, 4000) == 0   # inconsistent plan
    assert lib.rr_engine_auto_min_rows() > 0
    assert lib.rr_workspace_bytes(_cabi.RR_OP_GRADPASS, 10 ** 6, 21, 2048, 4096,
                                  0, 0, _cabi.RR_ENGINE_AUTO) > 0


def test_binding_signatures_match_the_header():
    """Argument COUNT and pointer/integer/float kind of every ctypes signature
    against the prototype in include/revrand_b200.h (a changed prototype with a
    stale binding would otherwise only show up as garbage on the GPU box)."""
    hdr = open(os.path.join(ROOT, "include", "revrand_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = dict(re.findall(r"\b(rr_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S))
    assert set(protos) == set(_cabi.SIGNATURES)

    def kind_of_c(arg):
        arg = " ".join(arg.split())
        if "*" in arg:
            return "ptr"
        if arg.startswith(("float", "double")):
            return "float"
        return "int"

    def kind_of_ctypes(t):
        if t in (ctypes.c_float, ctypes.c_double):
            return "float"
        if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents"):
            return "ptr"
        return "int"

    for name, (res, args) in _cabi.SIGNATURES.items():
        cargs = [a for a in protos[name].split(",") if a.strip() and a.strip() != "void"]
        assert len(cargs) == len(args), (name, cargs, args)
        assert [kind_of_c(a) for a in cargs] == [kind_of_ctypes(t) for t in args], name


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    b = bf.RandomRBF(nbases=4, Xdim=2, random_state=0)
    with pytest.raises(_cabi.RevrandB200Error):
        b.transform(np.zeros((3, 2)))
    with pytest.raises(_cabi.RevrandB200Error):
        rr.StandardLinearModel(basis=b).fit(np.zeros((5, 2)), np.zeros(5))


def test_weight_samplers_match_reference_draws(golden):
    g = golden["bases"]
    for cls in cases.RANDOM_BASES:
        for (d, K, N) in cases.BASIS_SHAPES:
            for seed in cases.BASIS_SEEDS:
                b = helpers.make_basis(cls, K, d, seed, False, 1.0)
                key = cases.basis_case_key(cls, d, K, N, seed, False)
                if cls == "FastFoodRBF":
                    assert np.array_equal(b.B, g[key + "/B"])
                    assert np.array_equal(b.PI, g[key + "/PI"])
                    assert np.array_equal(b.G, g[key + "/G"])
                    assert np.array_equal(b.S, g[key + "/S"])
                else:
                    assert np.array_equal(b.W, g[key + "/W"]), key


def test_fastfood_dense_image_equals_structured_projection():
    rs = np.random.RandomState(3)
    for d, K in ((6, 30), (21, 96), (1, 5)):
        b = bf.FastFoodRBF(nbases=K, Xdim=d, random_state=5)
        X = rs.randn(17, d)
        VX = orc.fastfood_vx(X, b.B, b.G, b.PI, b.S)
        np.testing.assert_allclose(X.dot(b._freqs()), VX, rtol=1e-10, atol=1e-10)
        assert b.get_dim(X) == 2 * b.n


def test_parameter_and_bounds():
    assert Positive().lower == 1e-14 and Positive(5).upper == 5
    with pytest.raises(ValueError):
        Bound(2, 1)
    with pytest.raises(ValueError):
        Parameter(-1., Positive())
    p = Parameter(gamma(1.), Positive(), shape=(3,))
    assert p.is_random and p.shape == (3,) and np.allclose(p.value, 1.0)
    draws = p.rvs(np.random.RandomState(0))
    assert draws.shape == (3,) and np.all(draws > 0)
    assert not Parameter().has_value and Parameter(1.).is_scalar
    assert Bound(0, 1).check(0.5) and not Bound(0.1, 1).check(0.05)
    b2 = pickle.loads(pickle.dumps(Positive(3.)))
    assert isinstance(b2, Positive) and b2.upper == 3.


def test_basis_protocol_shapes_and_routing():
    d = 4
    base = bf.RandomRBF(nbases=5, Xdim=d, random_state=0) + \
        bf.RandomMatern32(nbases=3, Xdim=2, random_state=1,
                          lenscale=Parameter(np.ones(2), Positive()),
                          apply_ind=[0, 2]) + bf.LinearBasis(onescol=True)
    X = np.zeros((3, d))
    assert base.get_dim(X) == 10 + 6 + 5
    assert isinstance(base.params, list) and len(base.params) == 2
    assert len(base.regularizer) == 3
    diag, slices = base.regularizer_diagonal(X, 1., 2., 3.)
    assert diag.shape == (21,) and [s.start for s in slices] == [0, 10, 16]
    assert np.all(diag[10:16] == 2.)
    routed = base._route([1.5, np.array([2., 3.])])
    assert routed[0] == [1.5] and np.all(routed[1][0] == [2., 3.]) and routed[2] == []
    blocks = base._blocks(d, [1.5, np.array([2., 3.])])
    assert [b.kind for b in blocks] == ["trig", "trig", "extra"]
    assert list(blocks[1].cols) == [0, 2]
    # sum() of bases and single-base collapse of params
    s = sum([bf.RandomRBF(nbases=2, Xdim=d, random_state=0), bf.BiasBasis()])
    assert isinstance(s, bf.BasisCat) and isinstance(s.params, Parameter)
    assert not bf.LinearBasis().params.has_value
    with pytest.raises(ValueError):
        bf.RandomRBF(nbases=2, Xdim=3, lenscale=Parameter(np.ones(2), Positive()))
    with pytest.raises(ValueError):
        bf.RandomRBF(nbases=2, Xdim=3)._check_dim(2, None)


def test_apply_grad_structure():
    f = lambda g: g.sum()
    assert bf.apply_grad(f, []) == []
    assert bf.apply_grad(f, np.ones((3, 2))) == 6
    assert bf.apply_grad(f, np.ones((3, 2, 4))).shape == (4,)
    out = bf.apply_grad(f, (g for g in [np.ones((2, 2)), np.ones((2, 2, 3))]))
    assert out[0] == 4 and out[1].shape == (3,)
    assert bf.apply_grad(f, [np.ones((2, 2))]) == 4


def test_layout_roundtrip():
    params = [Parameter(1., Positive()), [Parameter(2., Positive()),
                                          Parameter(3., Positive())],
              Parameter(), Parameter(np.arange(6.).reshape(2, 3), Bound())]
    lay = Layout.of_parameters(params)
    vals = [1., [2., 3.], [], np.arange(6.).reshape(2, 3)]
    flat = flatten_values(vals)
    assert flat.shape == (9,)
    back = lay.unflatten(flat)
    assert back[0] == 1. and back[1] == [2., 3.] and back[2] == []
    assert np.array_equal(back[3], vals[3])


def _quadratic_data(rs):
    x = rs.randn(200)
    y = 0.5 * x ** 2 + 2.0 * x + 3.0
    return np.vstack((x, y)).T


def test_structured_logtrick_minimizer_recovers_quadratic():
    rs = np.random.RandomState(1)
    data = _quadratic_data(rs)

    def obj(a, bc, data):
        b, c = bc
        x, y = data[:, 0], data[:, 1]
        r = a * x ** 2 + b * x + c - y
        return (r ** 2).sum(), [2 * (r * x ** 2).sum(),
                                [2 * (r * x).sum(), 2 * r.sum()]]
    nmin = structured_minimizer(logtrick_minimizer(minimize))
    params = [Parameter(gamma(2.), Positive()),
              [Parameter(1., Positive()), Parameter(1., Bound())]]
    res = nmin(obj, params, args=(data,), method='L-BFGS-B', jac=True,
               nstarts=20, random_state=np.random.RandomState(0))
    a, (b, c) = res.x
    assert np.allclose([a, b, c], [0.5, 2.0, 3.0], atol=1e-3)


@pytest.mark.parametrize("upd", [Adam(alpha=0.1), AdaDelta(), AdaGrad(),
                                 Momentum(rho=0.5, eta=0.0005),
                                 SGDUpdater(eta=0.0005)])
def test_sgd_updaters_recover_quadratic(upd):
    rs = np.random.RandomState(2)
    data = _quadratic_data(rs)

    def grad(w, data):
        x, y = data[:, 0], data[:, 1]
        r = w[0] * x ** 2 + w[1] * x + w[2] - y
        return np.array([2 * (r * x ** 2).mean(), 2 * (r * x).mean(), 2 * r.mean()])
    res = sgd(grad, np.array([1., 1., 1.]), data, maxiter=8000, updater=upd,
              batch_size=20, random_state=np.random.RandomState(3))
    assert np.allclose(res.x, [0.5, 2.0, 3.0], atol=0.1)


def test_structured_sgd_with_bounds_and_logtrick():
    rs = np.random.RandomState(4)
    data = _quadratic_data(rs)

    def obj(a, b, c, data):
        x, y = data[:, 0], data[:, 1]
        r = a * x ** 2 + b * x + c - y
        return (r ** 2).mean(), [2 * (r * x ** 2).mean(), 2 * (r * x).mean(),
                                 2 * r.mean()]
    nsgd = structured_sgd(logtrick_sgd(sgd))
    params = [Parameter(gamma(1.), Positive()), Parameter(1., Positive()),
              Parameter(1., Bound(-10, 10))]
    res = nsgd(obj, params, data, eval_obj=True, maxiter=6000,
               updater=Adam(alpha=0.05), batch_size=20, nstarts=10,
               random_state=np.random.RandomState(5))
    assert np.allclose(res.x, [0.5, 2.0, 3.0], atol=0.05)


def test_gen_batch_draw_order_matches_reference_semantics():
    # permutations are drawn only when the previous one is exhausted
    data = np.arange(10)
    rs1, rs2 = np.random.RandomState(7), np.random.RandomState(7)
    got = np.concatenate([b[0] for b in gen_batch(data, 4, maxiter=5,
                                                  random_state=rs1)])
    perms = np.concatenate([rs2.permutation(10), rs2.permutation(10)])
    assert np.array_equal(got, perms)


def test_likelihood_numpy_protocol_matches_oracle():
    rs = np.random.RandomState(0)
    f = rs.randn(4, 7)
    y = rs.poisson(1.5, size=7).astype(float)
    n = (y + rs.randint(1, 4, size=7)).astype(float)
    for obj, lid, args in [(lk.Gaussian(), orc.LIK_GAUSSIAN, (0.4,)),
                           (lk.Bernoulli(), orc.LIK_BERNOULLI, ()),
                           (lk.Binomial(), orc.LIK_BINOMIAL, (n,)),
                           (lk.Poisson('exp'), orc.LIK_POISSON_EXP, ()),
                           (lk.Poisson('softplus'), orc.LIK_POISSON_SOFTPLUS, ())]:
        yy = np.minimum(y, 1.) if lid == orc.LIK_BERNOULLI else y
        a = args[0] if args else None
        np.testing.assert_allclose(obj.loglike(yy, f, *args),
                                   orc.lik_loglike(lid, yy, f, a), rtol=1e-10)
        np.testing.assert_allclose(obj.df(yy, f, *args),
                                   orc.lik_df(lid, yy, f, a), rtol=1e-10)
        np.testing.assert_allclose(obj.Ey(f, *args), orc.lik_Ey(lid, f, a),
                                   rtol=1e-10)
    assert lk.Bernoulli().dp(y, f) == []
    with pytest.raises(ValueError):
        lk.Poisson('log')


def test_sklearn_clone_and_pickle():
    b = bf.RandomRBF(nbases=5, Xdim=2, random_state=0) + bf.LinearBasis()
    slm = rr.StandardLinearModel(basis=b, var=Parameter(1., Positive()),
                                 nstarts=3, random_state=1)
    c = clone(slm)
    assert c.get_params()["nstarts"] == 3 and c.basis is not None
    glm = rr.GeneralizedLinearModel(likelihood=lk.Poisson(), basis=b, K=3)
    g2 = pickle.loads(pickle.dumps(clone(glm)))
    assert g2.K == 3 and isinstance(g2.likelihood, lk.Poisson)
    s2 = pickle.loads(pickle.dumps(slm))
    assert np.array_equal(s2.basis.bases[0].W, b.bases[0].W)


def test_shard_rows_partition():
    for N, W in ((10, 3), (7, 8), (1000000, 8)):
        spans = [_engine.shard_rows(N, r, W) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == N
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_blocked_spd_inverse_matches_numpy():
    """The GEMM-rich blocked inverse used on the GPU for D >= 1024
    (_engine.blocked_spd_inverse) against numpy, on CPU tensors, with a leaf
    size small enough to exercise three levels of recursion."""
    import torch
    rs = np.random.RandomState(3)
    n = 700
    A = rs.randn(n, 2 * n)
    S = A.dot(A.T) / n + 0.5 * np.eye(n)
    old = _engine._BLOCK_INV_LEAF
    _engine._BLOCK_INV_LEAF = 128
    try:
        L = torch.linalg.cholesky(torch.from_numpy(S))
        C = _engine.blocked_spd_inverse(L).numpy()
    finally:
        _engine._BLOCK_INV_LEAF = old
    ref = np.linalg.inv(S)
    assert np.max(np.abs(C - ref)) < 1e-10 * np.max(np.abs(ref))
    assert np.max(np.abs(C - C.T)) < 1e-12


def test_batched_triangular_inverse_levels_and_padding():
    """_engine._tri_inv_lower, level-by-level batched form: exact power-of-two
    blocking, identity padding, and the recursion it replaces give the same L^-1."""
    import torch
    rs = np.random.RandomState(5)
    old = _engine._BLOCK_INV_LEAF
    try:
        for n, leaf in ((512, 64), (700, 128), (130, 64), (37, 8)):
            A = rs.randn(n, 2 * n)
            L = torch.linalg.cholesky(torch.from_numpy(A.dot(A.T) / n + 0.5 * np.eye(n)))
            _engine._BLOCK_INV_LEAF = leaf
            out = torch.empty_like(L)
            _engine._tri_inv_lower(L, out)
            ref = np.linalg.inv(L.numpy())
            assert np.max(np.abs(out.numpy() - ref)) < 1e-11 * np.max(np.abs(ref)), (n, leaf)
            assert np.max(np.abs(np.triu(out.numpy(), 1))) == 0.0
            _engine._BLOCK_INV_BATCHED = False
            try:
                out2 = torch.empty_like(L)
                _engine._tri_inv_lower(L, out2)
            finally:
                _engine._BLOCK_INV_BATCHED = True
            assert np.max(np.abs(out2.numpy() - ref)) < 1e-11 * np.max(np.abs(ref))
    finally:
        _engine._BLOCK_INV_LEAF = old


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours)
    on a tiny workload: one JSON line with the keys of the bench contract."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "0", "--N", "4000", "--K", "32",
                          "--cpu-sample", "2000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def test_metrics_match_reference_docstring_examples():
    """revrand/metrics.py:25-31, 58-64, 91-95, 125-129 (doctest examples) plus
    closed-form values: mll is a LOSS (negative log-likelihood)."""
    from scipy.stats import norm
    from revrand_b200 import metrics
    rs = np.random.RandomState(3)
    y = rs.randn(100)
    assert metrics.smse(y, y) == 0.0
    assert metrics.smse(y, rs.randn(100)) >= 1.0
    mean_prob = -norm.logpdf(1e-2, loc=0)
    assert metrics.mll(y, y, 1) <= mean_prob
    assert metrics.mll(y, rs.randn(100), 1) >= mean_prob
    np.testing.assert_allclose(metrics.mll(y, y, 1.0), 0.5 * np.log(2 * np.pi))
    assert metrics.msll(y, y, 1, y) < 0
    assert metrics.msll(y, rs.randn(100), 1, y) >= 0
    np.testing.assert_allclose(
        metrics.msll(y, y, 1.0, y),
        metrics.mll(y, y, 1.0) + np.mean(norm.logpdf(y, y.mean(), y.std())))
    assert metrics.lins_ccc(y, y) > 0.99
    assert metrics.lins_ccc(y, np.zeros_like(y)) < 0.01


def test_glm_refuses_likelihoods_the_device_kernel_does_not_know():
    from revrand_b200 import glm, likelihoods as lk

    class Tweaked(lk.Gaussian):
        def df(self, y, f, var):
            return 2 * super().df(y, f, var)

    with pytest.raises(NotImplementedError):
        glm._check_device_likelihood(Tweaked(), [1.0], ())
    with pytest.raises(NotImplementedError):
        glm._check_device_likelihood(lk.Gaussian(), [1.0, 2.0], ())
    with pytest.raises(NotImplementedError):
        glm._check_device_likelihood(lk.Binomial(), [], (np.ones(3), np.ones(3)))
    glm._check_device_likelihood(lk.Binomial(), [], (np.ones(3),))
    glm._check_device_likelihood(lk.Gaussian(), [0.5], ())


def test_device_data_cache_key_sees_in_place_edits():
    """The device-resident copy behind direct ``_elbo`` calls is keyed on array
    identity AND a content fingerprint."""
    from revrand_b200.slm import StandardLinearModel
    X = np.random.RandomState(0).randn(500, 3)
    y = np.random.RandomState(1).randn(500)
    k0 = StandardLinearModel._fingerprint(X, y)
    assert k0 == StandardLinearModel._fingerprint(X, y)
    X[0, 0] += 1.0
    assert k0 != StandardLinearModel._fingerprint(X, y)
    X[0, 0] -= 1.0
    y[-1] = 7.0
    assert k0 != StandardLinearModel._fingerprint(X, y)


def test_glm_mixture_gradients_match_the_component_loop():
    """The vectorised mixture terms of the GLM step against the reference's loop
    over components (glm.py:249-260)."""
    from revrand_b200 import glm
    from revrand_b200.mathfun.special import logsumexp
    rs = np.random.RandomState(0)
    D, K, B = 12, 4, 7.5
    m, C = rs.randn(D, K), 0.1 + rs.rand(D, K)
    Lam = 0.5 + rs.rand(D)
    Edm, EdC = rs.randn(D, K), rs.randn(D, K)
    logNkl = glm._qmatrix(m, C)
    logzk = logsumexp(logNkl, axis=0)
    dm, dC = glm._mixture_gradients(m, C, Lam, logNkl, logzk, Edm, EdC, B)
    for k in range(K):
        Nkl_zk = np.exp(logNkl[:, k] - logzk[k])
        Nkl_zl = np.exp(logNkl[:, k] - logzk)
        alpha = Nkl_zk + Nkl_zl
        mkmj = m[:, k][:, None] - m
        iCkCj = 1. / (C[:, k][:, None] + C)
        rm = (B * Edm[:, k] - m[:, k] / Lam + (iCkCj * mkmj).dot(alpha)) / K
        rC = (B * EdC[:, k] - 1. / Lam + (iCkCj - (mkmj * iCkCj) ** 2).dot(alpha)) / (2 * K)
        np.testing.assert_allclose(dm[:, k], rm, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(dC[:, k], rC, rtol=1e-12, atol=1e-14)
