"""Likelihood objects for the generalised linear model.

Same classes and method protocol as revrand/likelihoods.py:18-545
(``loglike``, ``Ey``, ``df``, ``dp``, ``cdf``, ``params``).  Inside
``GeneralizedLinearModel`` the per-observation work runs in the GPU step
kernel, selected by ``_lik_id``; the numpy methods below serve the public
protocol (e.g. ``predict_logpdf`` on a handful of query points).
"""

from __future__ import annotations

import numpy as np
from scipy.special import expit, gammaln
from scipy.stats import bernoulli, binom, gamma, norm, poisson

from . import _cabi
from .btypes import Parameter, Positive
from .mathfun.special import safesoftplus, softplus


class Bernoulli(object):
    """Bernoulli likelihood with a logistic link."""

    _lik_id = _cabi.RR_LIK_BERNOULLI
    _params = Parameter()

    @property
    def params(self):
        return self._params

    @params.setter
    def params(self, params):
        self._params = params

    def loglike(self, y, f):
        y, f = np.broadcast_arrays(y, f)
        return y * f - np.logaddexp(0.0, f)

    def Ey(self, f):
        return expit(f)

    def df(self, y, f):
        y, f = np.broadcast_arrays(y, f)
        return y - expit(f)

    def dp(self, y, f, *args):
        return []

    def cdf(self, y, f):
        return bernoulli.cdf(y, expit(f))

    def __repr__(self):
        return "{}()".format(type(self).__name__)


class Binomial(Bernoulli):
    """Binomial likelihood; ``n`` (trials) is a per-observation argument."""

    _lik_id = _cabi.RR_LIK_BINOMIAL

    def loglike(self, y, f, n):
        return binom.logpmf(y, n=n, p=expit(f))

    def Ey(self, f, n):
        return expit(f) * n

    def df(self, y, f, n):
        y, f, n = np.broadcast_arrays(y, f, n)
        return y - expit(f) * n

    def cdf(self, y, f, n):
        return binom.cdf(y, n=n, p=expit(f))


class Gaussian(Bernoulli):
    """Gaussian likelihood with a learnable variance."""

    _lik_id = _cabi.RR_LIK_GAUSSIAN

    def __init__(self, var=Parameter(gamma(1., scale=1), Positive())):
        self.params = var

    def _check_param(self, param):
        if param is None:
            return self.params.value
        if not self.params.bounds.check(param):
            raise ValueError("Input parameter is out of bounds!")
        return param

    def loglike(self, y, f, var=None):
        var = self._check_param(var)
        y, f = np.broadcast_arrays(y, f)
        return -0.5 * (np.log(2 * np.pi * var) + (y - f) ** 2 / var)

    def Ey(self, f, var=None):
        self._check_param(var)
        return f

    def df(self, y, f, var=None):
        var = self._check_param(var)
        y, f = np.broadcast_arrays(y, f)
        return (y - f) / var

    def dp(self, y, f, var=None):
        var = self._check_param(var)
        y, f = np.broadcast_arrays(y, f)
        iv = 1. / var
        return 0.5 * (((y - f) * iv) ** 2 - iv)

    def cdf(self, y, f, var=None):
        var = self._check_param(var)
        return norm.cdf(y, loc=f, scale=np.sqrt(var))

    def __repr__(self):
        return "{}(var={})".format(type(self).__name__, self.params)


class Poisson(Bernoulli):
    """Poisson likelihood with an ``exp`` or ``softplus`` link."""

    def __init__(self, tranfcn='exp'):
        if tranfcn not in ('exp', 'softplus'):
            raise ValueError('Invalid transformation function specified!')
        self.tranfcn = tranfcn

    @property
    def _lik_id(self):
        return (_cabi.RR_LIK_POISSON_EXP if self.tranfcn == 'exp'
                else _cabi.RR_LIK_POISSON_SOFTPLUS)

    def _rate(self, f):
        return np.exp(f) if self.tranfcn == 'exp' else softplus(f)

    def loglike(self, y, f):
        y, f = np.broadcast_arrays(y, f)
        g = self._rate(f)
        logg = f if self.tranfcn == 'exp' else np.log(g)
        return y * logg - g - gammaln(y + 1)

    def Ey(self, f):
        return self._rate(f)

    def df(self, y, f):
        y, f = np.broadcast_arrays(y, f)
        if self.tranfcn == 'exp':
            return y - np.exp(f)
        return expit(f) * (y / safesoftplus(f) - 1)

    def cdf(self, y, f):
        return poisson.cdf(y, mu=self._rate(f))

    def __repr__(self):
        return "{}(tranfcn='{}')".format(type(self).__name__, self.tranfcn)
