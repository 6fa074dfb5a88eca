"""Bayesian generalised linear model (mixture-of-diagonal-Gaussians posterior,
reparameterised SVI) on the GPU.

Drop-in for revrand/glm.py:45-712.  One SVI step (= one ``_elbo`` call,
glm.py:205-322) is split into

  * the data-dependent part -- features of the minibatch, all K_mix * L
    reparameterised latent draws, likelihood derivatives and their
    contractions back onto weights / features / lengthscales -- which runs in
    ``rr_glm_step`` (csrc/rr_glm.cu), and
  * the O(K_mix^2 D) entropy-bound / prior terms (glm.py:222-223, 249-271),
    which are assembled on the host in float64.
"""

from __future__ import annotations

import logging
from itertools import chain

import numpy as np
from scipy.stats.distributions import gamma, norm
from sklearn.base import BaseEstimator, RegressorMixin
from sklearn.utils import check_random_state
from sklearn.utils.validation import check_array, check_is_fitted, check_X_y

from . import _engine as eng
from . import config
from .basis_functions import LinearBasis
from .btypes import Bound, Parameter, Positive
from .likelihoods import Gaussian
from .mathfun.special import logsumexp
from .optimize import Adam, logtrick_sgd, sgd, structured_sgd

log = logging.getLogger(__name__)

WGTRND = norm()                 # initial distribution of mixture means
COVRND = gamma(a=2, scale=0.5)  # initial distribution of mixture variances
LOGITER = 500                   # iterations between ELBO log lines


_DEVICE_LIKELIHOODS = None


def _check_device_likelihood(likelihood, lpars, largs):
    """The SVI step evaluates the likelihood ON THE DEVICE, selected by
    ``_lik_id``, with at most one scalar parameter and one per-row argument.
    A subclass that overrides loglike / df / dp, or a likelihood with more
    parameters, would silently train on the wrong function (the reference calls
    the object's own methods, likelihoods.py:36-44): refuse instead."""
    global _DEVICE_LIKELIHOODS
    if _DEVICE_LIKELIHOODS is None:
        from . import likelihoods as lk
        _DEVICE_LIKELIHOODS = (lk.Gaussian, lk.Bernoulli, lk.Binomial, lk.Poisson)
    if type(likelihood) not in _DEVICE_LIKELIHOODS:
        raise NotImplementedError(
            "GeneralizedLinearModel runs its likelihood on the device and knows "
            "Gaussian, Bernoulli, Binomial and Poisson; %s is not one of them "
            "(subclasses are not dispatched to their overridden methods)"
            % type(likelihood).__name__)
    if sum(1 for p in lpars if np.size(p) > 0) > 1 or any(np.size(p) > 1 for p in lpars):
        raise NotImplementedError("device likelihoods take at most one scalar parameter")
    if len(largs) > 1:
        raise NotImplementedError("device likelihoods take at most one per-row argument")


def _aslist(a):
    return a if isinstance(a, list) else [a]


def _qmatrix(m, C):
    """logq[j, i] = log N(m_i; m_j, diag(C_i + C_j)) (glm.py:697-712),
    vectorised over the K_mix x K_mix pairs."""
    v = C[:, :, None] + C[:, None, :]
    dm = m[:, :, None] - m[:, None, :]
    return -0.5 * (np.log(2 * np.pi * v) + dm ** 2 / v).sum(axis=0)


def _mixture_gradients(m, C, Lam, logNkl, logzk, Edm, EdC, B):
    """d ELBO / d m, d ELBO / d C for all mixture components at once
    (glm.py:249-260: likelihood term + prior term + mixture-entropy term),
    vectorised over the K_mix x K_mix component pairs."""
    K = m.shape[1]
    # alpha[j, k] = N_jk / z_k + N_jk / z_j   (rows j: the other component)
    alpha = np.exp(logNkl - logzk[None, :]) + np.exp(logNkl - logzk[:, None])
    mkmj = m[:, :, None] - m[:, None, :]            # [:, k, j] = m_k - m_j
    iCkCj = 1. / (C[:, :, None] + C[:, None, :])    # [:, k, j]
    aT = alpha.T[None, :, :]                        # [1, k, j] = alpha[j, k]
    dm = (B * Edm - m / Lam[:, None] + (iCkCj * mkmj * aT).sum(axis=2)) / K
    dC = (B * EdC - 1. / Lam[:, None]
          + ((iCkCj - (mkmj * iCkCj) ** 2) * aT).sum(axis=2)) / (2 * K)
    return dm, dC


def _reshape_likelihood_args(likelihood_args, N):
    out = []
    for arg in likelihood_args:
        if np.isscalar(arg):
            arg = arg * np.ones(N)
        if (np.shape(arg)[0] != N) and (len(arg) != 0):
            raise ValueError("Likelihood arguments not a compatible shape!")
        out.append(arg)
    return tuple(out)


class GeneralizedLinearModel(BaseEstimator, RegressorMixin):
    """Bayesian GLM trained by stochastic variational inference.

    Parameters
    ----------
    likelihood : object from :mod:`revrand_b200.likelihoods`
    basis : Basis
    K : int
        number of diagonal Gaussian mixture components of the posterior.
    maxiter : int
        number of SGD iterations.
    batch_size : int
        minibatch size.
    updater : SGDUpdater, optional
        update rule (default ``Adam()``).
    nsamples : int
        reparameterisation draws per mixture component per step.
    nstarts : int
        random candidate starts evaluated before SGD.
    random_state : None, int or RandomState
    """

    def __init__(self, likelihood=Gaussian(), basis=LinearBasis(), K=10,
                 maxiter=3000, batch_size=10, updater=None, nsamples=50,
                 nstarts=500, random_state=None):
        self.likelihood = likelihood
        self.basis = basis
        self.K = K
        self.maxiter = maxiter
        self.batch_size = batch_size
        self.updater = updater
        self.nsamples = nsamples
        self.nstarts = nstarts
        self.random_state = random_state
        self.random_ = check_random_state(self.random_state)

    # -- training ------------------------------------------------------------------
    def fit(self, X, y, likelihood_args=()):
        """Learn variational posterior and hyper-parameters (glm.py:139-203)."""
        X, y = check_X_y(X, y)
        from .basis_functions import require_model_support
        require_model_support(self.basis)
        N, _ = X.shape
        self.B_ = N / self.batch_size
        self.D_ = self.basis.get_dim(X)
        likelihood_args = _reshape_likelihood_args(likelihood_args, N)
        # the training set lives on the device; minibatches are gathered there
        data = (eng.to_device(X), eng.to_device(y)) + tuple(
            eng.to_device(np.asarray(a, dtype=float)) for a in likelihood_args)
        params = [Parameter(WGTRND, Bound(), shape=(self.D_, self.K)),
                  Parameter(COVRND, Positive(), shape=(self.D_, self.K)),
                  self.basis.regularizer, self.likelihood.params,
                  self.basis.params]
        log.info("Optimising parameters...")
        self._it = -self.nstarts
        self._plan_cache = None
        from . import _svi
        if _svi.supported(self, likelihood_args):
            res = self._fit_on_device(params, data)
        else:
            nsgd = structured_sgd(logtrick_sgd(sgd))
            res = nsgd(self._elbo, params, data, eval_obj=True,
                       maxiter=self.maxiter, updater=self.updater,
                       batch_size=self.batch_size, random_state=self.random_,
                       nstarts=self.nstarts)
        (self.weights_, self.covariance_, self.regularizer_,
         self.like_hypers_, self.basis_hypers_) = res.x
        log.info("Finished! reg = {}, likelihood_hypers = {}, "
                 "basis_hypers = {}, message: {}."
                 .format(self.regularizer_, self.like_hypers_,
                         self.basis_hypers_, res.message))
        self._plan_cache = None
        return self

    def _fit_on_device(self, params, data):
        """The optimisation of ``fit`` with the SGD loop resident on the device
        (``_svi.DeviceSVI``): same composition as structured_sgd(logtrick_sgd(sgd)) --
        an initial draw, the random starts (each on the next minibatch, objective
        through ``_elbo``), then ``maxiter`` steps."""
        from . import _svi
        from .optimize.sgd import gen_batch
        from .optimize.structured import _map_params, _random_starts, flatten_values
        x0 = flatten_values(_map_params(lambda p: p.rvs(None), params))
        run = _svi.DeviceSVI(self, params, data, self.maxiter, self.random_, x0=x0)
        if self.nstarts > 0:
            if config.GLM_DEVICE_STARTS:
                run.random_starts(params, self.nstarts)
            else:
                data_gen = gen_batch(data, self.batch_size, random_state=self.random_)
                run.set_x(flatten_values(_random_starts(self._elbo, params, True, (),
                                                        self.nstarts, self.random_, data_gen)))
        return run.run().result()

    def svi_stepper(self, X, y, likelihood_args=(), maxiter=10 ** 9):
        """The loop body of ``fit`` as an object: ``step()`` runs ONE SVI
        iteration (fresh minibatch gathered on the device, ``_elbo``, update)
        through exactly the wrappers ``fit`` composes.  For benchmarks and for
        callers that interleave training with other work."""
        X, y = check_X_y(X, y)
        from .basis_functions import require_model_support
        from .optimize.sgd import SGDRun
        require_model_support(self.basis)
        N, _ = X.shape
        self.B_ = N / self.batch_size
        self.D_ = self.basis.get_dim(X)
        likelihood_args = _reshape_likelihood_args(likelihood_args, N)
        data = (eng.to_device(X), eng.to_device(y)) + tuple(
            eng.to_device(np.asarray(a, dtype=float)) for a in likelihood_args)
        params = [Parameter(WGTRND, Bound(), shape=(self.D_, self.K)),
                  Parameter(COVRND, Positive(), shape=(self.D_, self.K)),
                  self.basis.regularizer, self.likelihood.params,
                  self.basis.params]
        self._it = 1            # > 0: no ELBO logging evaluation on benchmark steps
        self._plan_cache = None
        from . import _svi
        if _svi.supported(self, likelihood_args):
            return _DeviceStepper(_svi.DeviceSVI(self, params, data, maxiter, self.random_),
                                  self, X.shape[1])
        holder = {}

        def capture(fun, x0, data, **kw):
            holder["run"] = SGDRun(fun, x0, data, **kw)
            return holder["run"].result()
        structured_sgd(logtrick_sgd(capture))(
            self._elbo, params, data, eval_obj=True, maxiter=maxiter, updater=self.updater,
            batch_size=self.batch_size, random_state=self.random_, nstarts=0)
        return _SVIStepper(self, holder["run"], X.shape[1])

    def _noise(self, Kmix, L, D, dev):
        """Reparameterisation noise eps (K_mix, L, D) as a device tensor.

        ``config.GLM_HOST_RNG`` draws it from ``self.random_`` in the
        reference's order (glm.py:300, one (L, D) block per component);
        otherwise it comes from the device generator seeded from
        ``self.random_`` once.
        """
        t = eng.torch()
        if config.GLM_HOST_RNG:
            e = np.stack([self.random_.randn(L, D) for _ in range(Kmix)])
            return eng.to_device(e)
        gen = self._device_generator(dev)
        return t.randn((Kmix, L, D), generator=gen, device=dev, dtype=t.float32)

    def _device_generator(self, dev):
        """The device noise generator, seeded from ``self.random_`` on first use (one
        ``randint`` draw, at the model's first objective evaluation)."""
        gen = getattr(self, "_devgen", None)
        if gen is None:
            gen = eng.torch().Generator(device=dev)
            gen.manual_seed(int(self.random_.randint(0, 2 ** 31 - 1)))
            self._devgen = gen
        return gen

    def _get_plan(self, d, bpars):
        plan = getattr(self, "_plan_cache", None)
        if plan is None or plan.d != d:
            plan = self.basis._plan(d, bpars)
            self._plan_cache = plan
        elif plan.trig:
            plan.set_lenscales(bpars)
        return plan

    def _elbo(self, m, C, reg, lpars, bpars, X, y, *largs):
        """(-ELBO, [-dm, -dC, dreg, dlpars, dbpars]) for one minibatch; same
        contract as glm.py:205-294."""
        t = eng.require_cuda()
        D, K = m.shape
        L = self.nsamples
        it = getattr(self, "_it", 0)
        dolog = (it % LOGITER == 0) or (it == self.maxiter - 1)
        calc_ll = dolog or (it < 0)
        Xd = X if isinstance(X, t.Tensor) else eng.to_device(np.asarray(X, float))
        yd = y if isinstance(y, t.Tensor) else eng.to_device(np.asarray(y, float))
        largd = None
        if len(largs):
            a = largs[0]
            largd = a if isinstance(a, t.Tensor) else eng.to_device(np.asarray(a, float))
        hyps = [h for h in _aslist(bpars)]
        if len(hyps) == 1 and np.size(hyps[0]) == 0:
            hyps = []
        plan = self._get_plan(Xd.shape[1], hyps)
        lpl = _aslist(lpars)
        _check_device_likelihood(self.likelihood, lpl, largs)
        has_lpar = len(lpl) > 0 and np.size(lpl[0]) > 0
        lik_param = float(lpl[0]) if has_lpar else 1.0

        eps = self._noise(K, L, D, Xd.device)
        Edm, EdC, R, Ell_d, dlp_d = eng.glm_step(
            plan, Xd, yd, largd, eng.to_device(m), eng.to_device(C), eps,
            self.likelihood._lik_id, lik_param, want_ll=calc_ll, want_R=True)
        parts = [Edm.double().reshape(-1), EdC.double().reshape(-1), dlp_d]
        if calc_ll:
            parts.append(Ell_d)
        if R is not None:
            WR = plan._Wfull_dev * R[:, :plan.ktot]
            g = t.stack([WR[:, ko:ko + b.K].sum(dim=1)
                         for b, ko in zip(plan.trig, plan.freq_offsets)])
            parts.append(g.reshape(-1))
        host = t.cat(parts).cpu().numpy()
        pos = 0

        def take(n, shape=None):
            nonlocal pos
            out = host[pos:pos + n]
            pos += n
            return out if shape is None else out.reshape(shape)
        Edm_h, EdC_h = take(D * K, (D, K)), take(D * K, (D, K))
        dlp_sum = take(1)[0]
        Ell = take(K) if calc_ll else None
        g_h = take(len(plan.trig) * plan.d, (len(plan.trig), plan.d)) \
            if R is not None else None

        # ---- host assembly (float64) -------------------------------------------
        Xprobe = np.zeros((1, Xd.shape[1]))
        Lam, slices = self.basis.regularizer_diagonal(Xprobe, *_aslist(reg))
        iL = 1. / Lam[:, None]
        logNkl = _qmatrix(m, C)
        logzk = logsumexp(logNkl, axis=0)
        dm, dC = _mixture_gradients(m, C, Lam, logNkl, logzk, Edm_h, EdC_h, self.B_)

        def dreg(s):
            return -0.5 * (((m[s] ** 2 + C[s]) * iL[s] ** 2).sum() / K
                           - iL[s].sum())
        dL = [dreg(s) for s in slices] if isinstance(slices, list) else dreg(slices)
        dlpars = [np.float64(-dlp_sum / K)] if has_lpar else []
        if isinstance(lpars, list) and not has_lpar:
            dlpars = []
        # basis hyper-parameter gradients: -(EdPhi * dPhi).sum()  (glm.py:274)
        dbl = []
        for bi, b in enumerate(plan.trig):
            rows = np.arange(plan.d) if b.cols is None else b.cols
            gi = g_h[bi, rows]
            ls = b.lenscale
            if len(ls) > 1:
                dbl.append(gi / ls ** 2)
            elif config.REFERENCE_COMPAT:
                dbl.append(float(gi[0] / ls[0] ** 2))
            else:
                dbl.append(float(gi.sum() / ls[0] ** 2))
        dbpars = dbl if len(dbl) != 1 else dbl[0]

        ELBO = -np.inf
        if calc_ll:
            ELBO = (Ell.sum() * self.B_ - 0.5 * D * K * np.log(2 * np.pi)
                    - 0.5 * K * np.log(Lam).sum()
                    - 0.5 * ((m ** 2 + C) * iL).sum()
                    - logzk.sum() + np.log(K)) / K
        if dolog:
            log.info("{}Iter {}: ELBO = {}, reg = {}, like_hypers = {}, "
                     "basis_hypers = {}".format(
                         "Random starts: " if it < 0 else "", it, ELBO, reg,
                         lpars, bpars))
        self._it = it + 1
        return -ELBO, [-dm, -dC, dL, dlpars, dbpars]

    # -- prediction ------------------------------------------------------------------
    def _draw_weights(self, nsamples):
        D, K = self.weights_.shape
        k = self.random_.randint(0, K, size=(nsamples,))
        return self.weights_[:, k] + self.random_.randn(D, nsamples) \
            * np.sqrt(self.covariance_[:, k])

    def _fitted_plan(self, d):
        hyps = [h for h in _aslist(self.basis_hypers_)]
        if len(hyps) == 1 and np.size(hyps[0]) == 0:
            hyps = []
        return self.basis._plan(d, hyps)

    def _lik_param(self):
        lh = _aslist(self.like_hypers_)
        return float(lh[0]) if len(lh) and np.size(lh[0]) else 1.0

    def predict(self, X, nsamples=200, likelihood_args=()):
        Ey, _ = self.predict_moments(X, nsamples, likelihood_args)
        return Ey

    def predict_moments(self, X, nsamples=200, likelihood_args=()):
        """Monte-Carlo predictive mean and variance (glm.py:349-418); the
        feature map, the latent draws and the link are evaluated on the GPU."""
        check_is_fitted(self, ['weights_', 'covariance_', 'basis_hypers_',
                               'like_hypers_', 'regularizer_'])
        X = check_array(X)
        N = X.shape[0]
        w = self._draw_weights(nsamples)            # (D, S)
        plan = self._fitted_plan(X.shape[1])
        largs = _reshape_likelihood_args(likelihood_args, N)
        largd = eng.to_device(np.asarray(largs[0], float)) if largs else None
        Ey, Ey2 = eng.glm_predict(plan, eng.to_device(X),
                                  eng.to_device(np.ascontiguousarray(w.T)),
                                  self.likelihood._lik_id, self._lik_param(),
                                  largd, want_sq=True)
        Ey = Ey.double().cpu().numpy()
        Vy = np.maximum(Ey2.double().cpu().numpy() - Ey ** 2, 0.0)
        return Ey, Vy

    def _sample_f_dev(self, X, nsamples):
        """Latent function draws f (N, nsamples), float32, on the device."""
        check_is_fitted(self, ['weights_', 'covariance_', 'basis_hypers_',
                               'like_hypers_', 'regularizer_'])
        X = check_array(X)
        w = self._draw_weights(nsamples)
        plan = self._fitted_plan(X.shape[1])
        Phi = eng.features(plan, eng.to_device(X)).double()
        F = Phi @ eng.to_device(w, eng.torch().float64)
        return F.float().contiguous()

    def _sample_f(self, X, nsamples):
        """Latent function draws f (N, nsamples) as a numpy array."""
        return self._sample_f_dev(X, nsamples).double().cpu().numpy()

    def _sample_func(self, X, nsamples, genaxis=1):
        F = self._sample_f(X, nsamples)
        if genaxis == 1:
            return (F[:, s] for s in range(F.shape[1]))
        if genaxis == 0:
            return (F[n] for n in range(F.shape[0]))
        raise ValueError("Invalid axis to generate samples from")

    def _largs_tuple(self, likelihood_args):
        return tuple(chain(_aslist(self.like_hypers_), likelihood_args))

    def predict_logpdf(self, X, y, nsamples=200, likelihood_args=()):
        X, y = check_X_y(X, y)
        F = self._sample_f(X, nsamples)
        ps = self.likelihood.loglike(y[:, None], F, *[
            np.asarray(a)[:, None] if np.ndim(a) else a
            for a in self._largs_tuple(likelihood_args)])
        return ps.mean(axis=1), ps.min(axis=1), ps.max(axis=1)

    def _larg_dev(self, likelihood_args, N):
        largs = _reshape_likelihood_args(likelihood_args, N)
        if len(largs) > 1:
            raise NotImplementedError("device likelihoods take at most one per-row argument")
        return eng.to_device(np.asarray(largs[0], float)) if largs else None

    def predict_cdf(self, X, quantile, nsamples=200, likelihood_args=()):
        """Monte-Carlo predictive CDF P(y* <= quantile) and its min / max over
        the draws (glm.py:468-516); the likelihood CDF runs on the device."""
        _check_device_likelihood(self.likelihood, _aslist(self.like_hypers_),
                                 likelihood_args)
        F = self._sample_f_dev(X, nsamples)
        p, pmin, pmax = eng.glm_cdf(F, self.likelihood._lik_id, self._lik_param(),
                                    quantile, self._larg_dev(likelihood_args, F.shape[0]))
        return tuple(v.double().cpu().numpy() for v in (p, pmin, pmax))

    def predict_interval(self, X, percentile, nsamples=200, likelihood_args=(),
                         multiproc=True):
        """Predictive quantile interval (glm.py:518-570, 669-694): the roots of
        the Monte-Carlo CDF at (1 -+ percentile) / 2 inside the reference's
        bracket, all query rows bisected concurrently on the device (the
        reference farms brentq out to a process pool; ``multiproc`` is accepted
        and ignored)."""
        _check_device_likelihood(self.likelihood, _aslist(self.like_hypers_),
                                 likelihood_args)
        F = self._sample_f_dev(X, nsamples)
        lo_p = (1 - percentile) / 2
        ql, qu = eng.glm_quantiles(F, self.likelihood._lik_id, self._lik_param(), lo_p,
                                   1 - lo_p, self._larg_dev(likelihood_args, F.shape[0]))
        return ql.cpu().numpy(), qu.cpu().numpy()

    def __getstate__(self):
        state = dict(self.__dict__)
        for k in ("_plan_cache", "_devgen"):
            state.pop(k, None)
        return state

    def __repr__(self):
        return "{}(likelihood={}, basis={}, K={}, maxiter={}, batch_size={}," \
            "updater={}, nsamples={}, nstarts={}, random_state={})".format(
                type(self).__name__, self.likelihood, self.basis, self.K,
                self.maxiter, self.batch_size, self.updater, self.nsamples,
                self.nstarts, self.random_state)


class _DeviceStepper(object):
    """``svi_stepper`` on the device-resident loop: nothing crosses the host
    boundary in a step except, amortised, the next chunk of minibatch indices."""

    def __init__(self, run, glm, d):
        self.run, self.glm, self.d = run, glm, d
        self.d2h_bytes = 0

    @property
    def h2d_bytes(self):
        return int(self.run.B * 8)      # this step's minibatch indices (uploaded in chunks)

    def step(self):
        return self.run.step()

    def objective(self):
        """-ELBO estimate of the last step (one 8-byte read, synchronises)."""
        r = self.run
        return float(r.objs[(r.it - 1) % r.ntrace].item())

    def device_ms(self, reps=5):
        return _SVIStepper.device_ms(self, reps)


class _SVIStepper(object):
    """See ``GeneralizedLinearModel.svi_stepper``."""

    def __init__(self, glm, run, d):
        self.glm, self.run, self.d = glm, run, d
        D, K = glm.D_, glm.K
        self.h2d_bytes = int(2 * D * K * 4)            # m, C as float32
        self.d2h_bytes = int(2 * D * K * 8 + 8 * 64)   # Edm, EdC and the scalars, float64

    def step(self):
        glm = self.glm
        if glm._it % LOGITER == 0:   # keep benchmark steps free of the logging ELBO
            glm._it += 1
        return self.run.step()

    def device_ms(self, reps=5):
        """CUDA-event time of the device part of a step (``rr_glm_step`` on one
        minibatch), without the host assembly and the update."""
        t = eng.torch()
        glm = self.glm
        D, K, L = glm.D_, glm.K, glm.nsamples
        rs = np.random.RandomState(0)
        m = eng.to_device(0.1 * rs.randn(D, K))
        C = eng.to_device(0.05 + 0.1 * np.abs(rs.randn(D, K)))
        M = glm.batch_size
        Xb = eng.to_device(rs.randn(M, self.d))
        yb = eng.to_device(rs.poisson(1.0, size=M).astype(float))
        plan = glm._get_plan(Xb.shape[1], glm.basis.params_values())
        eps = glm._noise(K, L, D, Xb.device)
        lik = glm.likelihood._lik_id
        ts = []
        for _ in range(reps + 1):
            a, b = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            a.record()
            eng.glm_step(plan, Xb, yb, None, m, C, eps, lik, 1.0, want_ll=False, want_R=True)
            b.record()
            t.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.mean(ts[1:]))


class GeneralisedLinearModel(GeneralizedLinearModel):
    """Alias with the British spelling."""
    pass
