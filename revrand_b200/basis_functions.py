"""Basis (feature-map) objects with learnable hyper-parameters, evaluated on
the GPU.

Drop-in for the random-kernel part of revrand/basis_functions.py: same class
names, constructor arguments, ``transform`` / ``grad`` / ``get_dim`` /
``params`` / ``regularizer`` / ``regularizer_diagonal`` protocol and ``+``
concatenation (reference :160-380, :816-1383, :1569-1790).  The numerics run
in ``librevrand_b200.so``; each basis also describes itself as *plan blocks*
so the models can fuse the whole concatenation into one kernel without ever
materialising Phi.

Random frequency matrices are drawn on the host with numpy's legacy
``RandomState`` in exactly the reference's draw order ("identical seeded
random bases"): :952-954, :993-995, :1034-1045, :1051-1065, :1198-1208,
:1342-1354.
"""

from __future__ import annotations

import math

import numpy as np
from scipy.linalg import hadamard as _sylvester
from scipy.linalg import qr
from scipy.stats import gamma
from scipy.stats import norm as norm_dist
from sklearn.utils import check_random_state

from . import _engine as eng
from . import config
from .btypes import Bound, Parameter, Positive


def _issequence(obj):
    return isinstance(obj, (list, tuple)) or (
        hasattr(obj, "__next__") or hasattr(obj, "send"))


def apply_grad(fun, grad):
    """Map ``fun`` over a structured basis gradient.

    Same contract as revrand/basis_functions.py:109-152: sequences (lists,
    tuples, generators) map element-wise and collapse when they hold a single
    entry, an empty gradient gives ``[]``, 2-D arrays are passed to ``fun``
    and 3-D arrays are mapped over their last axis.
    """
    if _issequence(grad):
        out = [apply_grad(fun, g) for g in grad]
        return out[0] if len(out) == 1 else out
    if len(grad) == 0:
        return []
    nd = np.ndim(grad)
    if nd in (1, 2):
        return fun(grad)
    if nd == 3:
        return np.array([fun(grad[:, :, i]) for i in range(grad.shape[2])])
    raise ValueError("Only up to 3d gradients allowed!")


def require_model_support(basis):
    """Raise NotImplementedError if ``basis`` (or a member of a concatenation)
    is a feature map only, i.e. cannot run inside the fused model passes."""
    for b in getattr(basis, "bases", [basis]):
        if not getattr(b, "_model_ok", True):
            raise NotImplementedError(
                "%s is available as a feature map (transform / grad) only: the "
                "fused StandardLinearModel / GeneralizedLinearModel passes handle "
                "trigonometric, linear, bias and polynomial columns"
                % type(b).__name__)


def _as_2d(X):
    X = np.asarray(X, dtype=np.float64)
    if X.ndim != 2:
        raise ValueError("X must be a 2-D array of shape (N, d)")
    return X


class Basis(object):
    """Base class: identity features, no hyper-parameters.

    Subclasses set ``_n_hypers`` (number of positional hyper-parameters their
    ``transform`` / ``grad`` accept) and implement ``_blocks``.
    """

    _n_hypers = 0
    _params = Parameter()
    _regularizer = Parameter(gamma(1.), Positive())

    def __init__(self, regularizer=None, apply_ind=None):
        self._set_common(regularizer, apply_ind)

    # -- construction helpers ------------------------------------------------
    def _set_common(self, regularizer, apply_ind):
        if regularizer is not None:
            if not regularizer.is_scalar:
                raise ValueError("Regularizer parameters have to be scalar!")
            if regularizer.bounds.lower <= 0:
                raise ValueError("Regularizer has to be bounded below by 0!")
            self._regularizer = regularizer
        if np.isscalar(apply_ind):
            apply_ind = [apply_ind]
        self.apply_ind = apply_ind

    def _cols(self, d):
        """Input columns this basis reads (None = all d of them)."""
        if self.apply_ind is None:
            return None
        return np.arange(d)[self.apply_ind]

    def _view(self, X):
        return X if self.apply_ind is None else X[:, self.apply_ind]

    # -- plan blocks -----------------------------------------------------------
    def _blocks(self, d, hypers):
        """Plan blocks of this basis for inputs with ``d`` columns."""
        cols = self._cols(d)
        src = np.arange(d) if cols is None else cols
        return [eng.ExtraBlock(src, np.zeros(len(src)))]

    def _plan(self, d, hypers):
        plan = eng.FeaturePlan(self._blocks(d, hypers), d)
        return plan

    # -- public protocol ---------------------------------------------------------
    def transform(self, X, *hypers):
        """Phi(X): (N, D) float64 array computed on the GPU."""
        X = _as_2d(X)
        plan = self._plan(X.shape[1], list(hypers))
        Phi = eng.features(plan, eng.to_device(X))
        return Phi.double().cpu().numpy()

    def grad(self, X, *hypers):
        """Gradient of the features wrt each hyper-parameter ([] if none)."""
        return []

    def get_dim(self, X):
        if not hasattr(self, "_D"):
            d = np.shape(X)[1]
            self._D = int(sum(b.width for b in
                              self._blocks(d, self.params_values())))
        return self._D

    def params_values(self):
        ps = self.params if isinstance(self.params, list) else [self.params]
        return [p.value for p in ps if p.has_value]

    def regularizer_diagonal(self, X, regularizer=None):
        reg = self.regularizer.value if regularizer is None else regularizer
        return np.full(self.get_dim(X), reg, dtype=float), slice(None)

    @property
    def params(self):
        return self._params

    @property
    def regularizer(self):
        return self._regularizer

    def __add__(self, other):
        return BasisCat([self, other])

    def __radd__(self, other):
        return self if other == 0 else self.__add__(other)

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_plan_cache", None)
        return state

    def __repr__(self):
        return "{}(regularizer={})".format(type(self).__name__, self.regularizer)


class BiasBasis(Basis):
    """A single constant column (reference :387-440)."""

    def __init__(self, offset=1., regularizer=None, apply_ind=None):
        self.offset = offset
        self._set_common(regularizer, apply_ind)

    def _blocks(self, d, hypers):
        return [eng.ExtraBlock([-1], [self.offset])]

    def __repr__(self):
        return "{}(offset={}, regularizer={})".format(
            type(self).__name__, self.offset, self.regularizer)


class LinearBasis(Basis):
    """[1, X] (or X when ``onescol`` is False); reference :443-493."""

    def __init__(self, onescol=True, regularizer=None, apply_ind=None):
        self.onescol = onescol
        self._set_common(regularizer, apply_ind)

    def _blocks(self, d, hypers):
        cols = self._cols(d)
        src = list(np.arange(d) if cols is None else cols)
        val = [0.0] * len(src)
        if self.onescol:
            src, val = [-1] + src, [1.0] + val
        return [eng.ExtraBlock(src, val)]

    def __repr__(self):
        return "{}(onescol={}, regularizer={})".format(
            type(self).__name__, self.onescol, self.regularizer)


class PolynomialBasis(Basis):
    """[1, X^1, ..., X^order], the powers of one input column adjacent
    (reference :496-576).  Rides inside fused plans as affine columns with an
    integer power."""

    def __init__(self, order, include_bias=True, regularizer=None, apply_ind=None):
        if order < 0:
            raise ValueError("Polynomial order must be positive")
        self.order = order
        self.include_bias = include_bias
        self._set_common(regularizer, apply_ind)

    def _blocks(self, d, hypers):
        cols = self._cols(d)
        cols = list(np.arange(d) if cols is None else cols)
        src, val, pw = [], [], []
        if self.include_bias:
            src, val, pw = [-1], [1.0], [1]
        for c in cols:
            for p in range(1, self.order + 1):
                src.append(c)
                val.append(0.0)
                pw.append(p)
        return [eng.ExtraBlock(src, val, pw)]

    def __repr__(self):
        return "{}(order={}, include_bias={}, regularizer={})".format(
            type(self).__name__, self.order, self.include_bias, self.regularizer)


class _LengthScaleBasis(Basis):
    """Shared lengthscale handling (reference :579-613)."""

    _n_hypers = 1

    def _init_lenscale(self, lenscale):
        if lenscale.shape != (self.d,) and lenscale.shape != ():
            raise ValueError("Parameter dimension doesn't agree with X"
                             " dimensions!")
        self._params = lenscale

    def _check_dim(self, Xdim, lenscale):
        if Xdim != self.d:
            raise ValueError("Dimensions of data inconsistent!")
        if lenscale is None:
            lenscale = self.params.value
        # the reference evaluates, but ignores, bounds.check here (:603)
        if np.isscalar(lenscale):
            lenscale = np.array([lenscale], dtype=float)
        lenscale = np.asarray(lenscale, dtype=float)
        if (self.params.shape == () and len(lenscale) == 1) \
                or np.shape(lenscale) == self.params.shape:
            return lenscale
        raise ValueError("Dimension of input parameter is inconsistent!")

    def _eff_d(self, d):
        cols = self._cols(d)
        return d if cols is None else len(cols)


class RadialBasis(_LengthScaleBasis):
    """exp(-||(x - c) / (2 l^2)||^2) around fixed centres (reference :616-731;
    the division by 2 l^2 before squaring is the reference's, :686-688).

    ``transform`` / ``grad`` run on the device; the fused model passes only
    know trigonometric and affine columns, so a model built on this basis
    raises NotImplementedError.
    """
    _kind = 0
    _model_ok = False

    def __init__(self, centres, lenscale=Parameter(gamma(1.), Positive()),
                 regularizer=None, apply_ind=None):
        centres = np.asarray(centres, dtype=float)
        self.M, self.d = centres.shape
        self.C = centres
        self._init_lenscale(lenscale)
        self._set_common(regularizer, apply_ind)

    def get_dim(self, X):
        return self.M

    def _blocks(self, d, hypers):
        raise NotImplementedError(
            "%s is available as a feature map (transform / grad) only: the fused "
            "StandardLinearModel / GeneralizedLinearModel passes handle "
            "trigonometric, linear, bias and polynomial columns" % type(self).__name__)

    def transform(self, X, lenscale=None):
        Xv = self._view(_as_2d(X))
        ls = self._check_dim(Xv.shape[1], lenscale)
        Phi = eng.centre_features(eng.to_device(Xv), self.C, ls, self._kind)
        return Phi.double().cpu().numpy()

    def grad(self, X, lenscale=None):
        Xv = self._view(_as_2d(X))
        ls = self._check_dim(Xv.shape[1], lenscale)
        _, dPhi = eng.centre_features(eng.to_device(Xv), self.C, ls, self._kind,
                                      want_grad=True)
        return dPhi.double().cpu().numpy()

    def __repr__(self):
        return "{}(centres={}, lenscale={}, regularizer={})".format(
            type(self).__name__, self.C, self.params, self.regularizer)


class SigmoidalBasis(RadialBasis):
    """expit(||(x - c) / l||) around fixed centres (reference :734-815)."""
    _kind = 1


class _RandomKernelBasis(_LengthScaleBasis):
    """cos/sin random features of a shift-invariant kernel (reference
    :816-914).  Phi = [cos(XW/l) | sin(XW/l)] / sqrt(nbases)."""

    def __init__(self, nbases, Xdim, lenscale=Parameter(gamma(1.), Positive()),
                 regularizer=None, random_state=None, apply_ind=None):
        self.d = Xdim
        self.n = nbases
        self.random_state = random_state
        self._random = check_random_state(random_state)
        self.W = self._weightsamples()
        self._init_lenscale(lenscale)
        self._set_common(regularizer, apply_ind)

    def _weightsamples(self):
        raise NotImplementedError

    # frequency matrix seen by the kernels (d_eff, n)
    def _freqs(self):
        return self.W

    def _blocks(self, d, hypers):
        ls = self._check_dim(self._eff_d(d), hypers[0] if hypers else None)
        return [eng.TrigBlock(self._freqs(), ls, self._cols(d))]

    def transform(self, X, lenscale=None):
        return super(_RandomKernelBasis, self).transform(X, lenscale)

    def grad(self, X, lenscale=None):
        """d Phi / d lenscale: (N, 2n) for a scalar lengthscale, (N, 2n, d) for
        ARD (reference :866-901).  With ``config.REFERENCE_COMPAT`` (default)
        the scalar case reproduces the reference, in which only input
        dimension 0 contributes."""
        Xv = self._view(_as_2d(X))
        ls = self._check_dim(Xv.shape[1], lenscale)
        out = eng.trig_grad(eng.to_device(Xv), self._freqs(), ls,
                            compat=config.REFERENCE_COMPAT)
        return out.double().cpu().numpy()

    def __repr__(self):
        return "{}(nbases={}, Xdim={}, lenscale={}, regularizer={}, " \
            "random_state={})".format(type(self).__name__, self.n, self.d,
                                      self.params, self.regularizer,
                                      self.random_state)


class RandomRBF(_RandomKernelBasis):
    """RBF kernel: W ~ N(0, I)."""

    def _weightsamples(self):
        return self._random.randn(self.d, self.n)


class RandomLaplace(_RandomKernelBasis):
    """Laplace kernel: W ~ standard Cauchy."""

    def _weightsamples(self):
        return self._random.standard_cauchy(size=(self.d, self.n))


class RandomCauchy(_RandomKernelBasis):
    """Cauchy kernel: W is a Gaussian scale mixture, N(0,I) * sqrt(2 Gamma(1))
    per frequency (multivariate Laplace)."""

    def _weightsamples(self):
        g = self._random.randn(self.d, self.n)
        z = self._random.standard_gamma(1., size=(1, self.n))
        return g * np.sqrt(2 * z)


class _RandomMatern(_RandomKernelBasis):
    _p = None

    def _weightsamples(self):
        # multivariate Student-t with df = 2 (p + 1/2): N(0,I) * sqrt(df/chi2)
        df = 2 * (self._p + 0.5)
        g = self._random.randn(self.d, self.n)
        u = self._random.chisquare(df, size=(self.n,))
        return g * np.sqrt(df / u)


class RandomMatern32(_RandomMatern):
    """Matern 3/2 kernel."""
    _p = 1


class RandomMatern52(_RandomMatern):
    """Matern 5/2 kernel."""
    _p = 2


class OrthogonalRBF(_RandomKernelBasis):
    """Orthogonal random features for the RBF kernel; as in the reference
    (:1198-1208) the chi scaling is applied to the *input dimensions*."""

    def _weightsamples(self):
        reps = int(np.ceil(self.n / self.d))
        blocks = [qr(self._random.randn(self.d, self.d))[0] for _ in range(reps)]
        Q = np.hstack(blocks)
        S = np.sqrt(self._random.chisquare(df=self.d, size=self.d))
        return S[:, None] * Q[:, :self.n]


class FastFoodRBF(_LengthScaleBasis):
    """FastFood approximation of the RBF features (reference :1211-1383).

    ``transform`` runs the structured S H G Pi H B projection as an in-kernel
    Walsh-Hadamard butterfly; inside the fused model passes the same linear
    map is used through its dense (d, n) image ``_freqs()``.
    """

    def __init__(self, nbases, Xdim, lenscale=Parameter(gamma(1.), Positive()),
                 regularizer=None, random_state=None, apply_ind=None):
        self.random_state = random_state
        self._random = check_random_state(random_state)
        self.nbases = nbases
        self.d = Xdim
        self.d2 = 2 ** int(np.ceil(np.log2(Xdim)))
        self.k = int(np.ceil(nbases / self.d2))
        self.n = self.d2 * self.k
        self._init_lenscale(lenscale)
        shape = (self.k, self.d2)
        self.B = self._random.randint(2, size=shape) * 2 - 1
        self.G = self._random.randn(*shape)
        self.PI = np.array([self._random.permutation(self.d2)
                            for _ in range(self.k)])
        chi = np.sqrt(self._random.chisquare(self.d2, size=shape))
        self.S = self.d2 * chi / np.linalg.norm(self.G, axis=1)[:, None]
        self._set_common(regularizer, apply_ind)

    def _freqs(self):
        """Dense (d, n) matrix V^T with VX = X V^T (exact, float64)."""
        if not hasattr(self, "_Wdense"):
            H = _sylvester(self.d2).astype(float) / self.d2
            rows = []
            for b, g, pi, s in zip(self.B, self.G, self.PI, self.S):
                M = H * b[None, :]                 # H diag(B)
                M = M[pi, :] * g[:, None]          # diag(G) Pi (...)
                M = H.dot(M) * (s * math.sqrt(self.d2))[:, None]
                rows.append(M[:, :self.d])
            self._Wdense = np.vstack(rows).T.copy()
        return self._Wdense

    def _blocks(self, d, hypers):
        ls = self._check_dim(self._eff_d(d), hypers[0] if hypers else None)
        return [eng.TrigBlock(self._freqs(), ls, self._cols(d))]

    def transform(self, X, lenscale=None):
        Xv = self._view(_as_2d(X))
        ls = self._check_dim(Xv.shape[1], lenscale)
        Phi = eng.fastfood_features(eng.to_device(Xv / ls), self.B, self.G,
                                    self.PI, self.S)
        return Phi.double().cpu().numpy()

    def grad(self, X, lenscale=None):
        Xv = self._view(_as_2d(X))
        ls = self._check_dim(Xv.shape[1], lenscale)
        out = eng.trig_grad(eng.to_device(Xv), self._freqs(), ls,
                            compat=config.REFERENCE_COMPAT)
        return out.double().cpu().numpy()

    def __getstate__(self):
        state = super(FastFoodRBF, self).__getstate__()
        state.pop("_Wdense", None)
        return state

    def __repr__(self):
        return "{}(nbases={}, Xdim={}, lenscale={}, regularizer={}, " \
            "random_state={})".format(type(self).__name__, self.nbases, self.d,
                                      self.params, self.regularizer,
                                      self.random_state)


class FastFoodGM(FastFoodRBF):
    """One component of a Gaussian spectral-mixture kernel approximation
    (reference :1386-1562): [cos(VX + Xm) | sin(VX + Xm) | cos(VX - Xm) |
    sin(VX - Xm)] / sqrt(2n) with learnable frequency means ``m`` and ARD
    lengthscales (both always of shape (d,)).

    ``transform`` evaluates the two phase families as two trigonometric plan
    blocks with frequency matrices V/l + m and V/l - m; ``grad`` returns
    (d Phi / d mean, d Phi / d lenscale).  As in the reference (Appendix B #6 of
    SURVEY.md) the pair of parameters does not fit the models' single-Parameter
    assumption, so this basis is a feature map only.
    """
    _n_hypers = 2
    _model_ok = False

    def __init__(self, nbases, Xdim, mean=Parameter(norm_dist(), Bound()),
                 lenscale=Parameter(gamma(1.), Positive()), regularizer=None,
                 random_state=None, apply_ind=None):
        # same draw order as the reference: dims, then matrices (:1436-1443)
        super(FastFoodGM, self).__init__(nbases, Xdim, lenscale=Parameter(1., Positive()),
                                         regularizer=regularizer,
                                         random_state=random_state, apply_ind=apply_ind)
        self._params = [self._init_param(mean), self._init_param(lenscale)]

    def _init_param(self, param):
        if param.shape == (self.d,):
            return param
        if param.shape in ((), (1,)):
            # scalar initial value -> the same value for every dimension (:1529-1538)
            if param.dist is not None:
                return Parameter(param.dist, param.bounds, shape=(self.d,))
            return Parameter(np.ones(self.d) * param.value, param.bounds)
        raise ValueError("Parameter dimension doesn't agree with X dimensions!")

    def _check_pair(self, Xdim, mean, lenscale):
        if Xdim != self.d:
            raise ValueError("Dimensions of data inconsistent!")
        out = []
        for v, p in zip((mean, lenscale), self._params):
            v = p.value if v is None else v
            v = np.atleast_1d(np.asarray(v, dtype=float))
            if v.shape != (self.d,):
                raise ValueError("Dimension of input parameter is inconsistent!")
            out.append(v)
        return out

    def _blocks(self, d, hypers):
        hypers = list(hypers) + [None] * (2 - len(hypers))
        mean, ls = self._check_pair(self._eff_d(d), hypers[0], hypers[1])
        V = self._freqs() / ls[:, None]
        amp = 1.0 / math.sqrt(2 * self.n)
        cols = self._cols(d)
        return [eng.TrigBlock(V + mean[:, None], 1.0, cols, amp=amp),
                eng.TrigBlock(V - mean[:, None], 1.0, cols, amp=amp)]

    def get_dim(self, X):
        return 4 * self.n

    def transform(self, X, mean=None, lenscale=None):
        X = _as_2d(X)
        plan = eng.FeaturePlan(self._blocks(X.shape[1], [mean, lenscale]), X.shape[1])
        return eng.features(plan, eng.to_device(X)).double().cpu().numpy()

    def grad(self, X, mean=None, lenscale=None):
        Xv = self._view(_as_2d(X))
        mean, ls = self._check_pair(Xv.shape[1], mean, lenscale)
        dm, dl = eng.gm_grad(eng.to_device(Xv), self._freqs(), mean, ls)
        return dm.double().cpu().numpy(), dl.double().cpu().numpy()

    @property
    def params(self):
        return self._params

    def __repr__(self):
        return "{}(nbases={}, Xdim={}, mean={}, lenscale={}, regularizer={}, " \
            "random_state={})".format(type(self).__name__, self.nbases, self.d,
                                      self.params[0], self.params[1],
                                      self.regularizer, self.random_state)


class BasisCat(object):
    """Column-wise concatenation of bases (reference :1569-1790).

    Positional hyper-parameters are routed to the bases in concatenation
    order, each base taking as many as its ``transform`` accepts.
    """

    def __init__(self, basis_list):
        flat = []
        for b in basis_list:
            flat.extend(b.bases if isinstance(b, BasisCat) else [b])
        self.bases = flat

    # -- hyper-parameter routing -------------------------------------------------
    def _route(self, hypers):
        hypers = list(hypers)
        out = []
        for b in self.bases:
            n = b._n_hypers
            out.append(hypers[:n])
            hypers = hypers[n:]
        return out

    def _blocks(self, d, hypers):
        blocks = []
        for b, h in zip(self.bases, self._route(hypers)):
            blocks.extend(b._blocks(d, h))
        return blocks

    def _plan(self, d, hypers):
        return eng.FeaturePlan(self._blocks(d, hypers), d)

    def _dims(self, X):
        if not hasattr(self, "_dims_cache"):
            self._dims_cache = [b.get_dim(X) for b in self.bases]
        return self._dims_cache

    # -- public protocol ---------------------------------------------------------
    def transform(self, X, *params):
        X = _as_2d(X)
        plan = self._plan(X.shape[1], params)
        return eng.features(plan, eng.to_device(X)).double().cpu().numpy()

    def grad(self, X, *params):
        """Generator of zero-padded gradients, one per hyper-parameterised
        base, (N, D) or (N, D, P) each."""
        X = _as_2d(X)
        N = X.shape[0]
        ends = np.cumsum([0] + list(self._dims(X)))
        D = int(ends[-1])
        for i, (b, h) in enumerate(zip(self.bases, self._route(params))):
            g = b.grad(X, *h)
            gs = g if isinstance(g, (list, tuple)) else (g,)
            for gg in gs:
                if len(gg) == 0:
                    continue
                shape = (N, D) if gg.ndim < 3 else (N, D, gg.shape[2])
                full = np.zeros(shape)
                full[:, ends[i]:ends[i + 1]] = gg
                yield full

    def get_dim(self, X):
        return int(np.sum(self._dims(X)))

    def params_values(self):
        return [v for b in self.bases for v in b.params_values()]

    @property
    def regularizer(self):
        return [b.regularizer for b in self.bases]

    def regularizer_diagonal(self, X, *regularizer):
        regs = list(regularizer) if len(regularizer) else [None] * len(self.bases)
        diag = np.concatenate([b.regularizer_diagonal(X, r)[0]
                               for b, r in zip(self.bases, regs)])
        ends = np.cumsum([0] + list(self._dims(X)))
        slices = [slice(int(a), int(e)) for a, e in zip(ends[:-1], ends[1:])]
        return diag, slices

    @property
    def params(self):
        plist = [b.params for b in self.bases if b.params.has_value]
        if not plist:
            return Parameter()
        return plist if len(plist) > 1 else plist[0]

    def __add__(self, other):
        return BasisCat(self.bases + (other.bases if isinstance(other, BasisCat)
                                      else [other]))

    def __radd__(self, other):
        return self if other == 0 else self.__add__(other)

    def __repr__(self):
        return "{}(basis_list={})".format(type(self).__name__, self.bases)
