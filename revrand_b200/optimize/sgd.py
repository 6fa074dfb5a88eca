"""Stochastic gradient descent driver and update rules (host side).

Equivalent of revrand/optimize/sgd.py:16-459: the flat parameter vector is
O(D * K_mix) floats, so the update itself stays on the host while each
objective call runs the minibatch through the GPU step kernel.  Minibatches
are drawn from endless permutations of the row indices in the reference's
RNG order (utils/rand.py:7-31, sgd.py:428-459), vectorised per batch.
"""

from __future__ import annotations

from itertools import chain

import numpy as np
from scipy.optimize import OptimizeResult
from sklearn.utils import check_random_state


class SGDUpdater(object):
    """Plain gradient step x <- x - eta * g."""

    def __init__(self, eta=0.1):
        self.eta = eta

    def __call__(self, x, grad):
        return x - self.eta * grad

    def reset(self):
        pass

    def __repr__(self):
        return "{}(eta={})".format(type(self).__name__, self.eta)


class AdaDelta(SGDUpdater):
    def __init__(self, rho=0.1, epsilon=1e-5):
        if not 0 <= rho <= 1:
            raise ValueError("Decay rate 'rho' must be between 0 and 1!")
        if epsilon <= 0:
            raise ValueError("Constant 'epsilon' must be > 0!")
        self.rho, self.epsilon = rho, epsilon
        self.reset()

    def reset(self):
        self.Eg2 = 0
        self.Edx2 = 0

    def __call__(self, x, grad):
        r = self.rho
        self.Eg2 = r * self.Eg2 + (1 - r) * grad ** 2
        dx = -grad * np.sqrt(self.Edx2 + self.epsilon) / np.sqrt(self.Eg2 + self.epsilon)
        self.Edx2 = r * self.Edx2 + (1 - r) * dx ** 2
        return x + dx

    def __repr__(self):
        return "{}(rho={}, epsilon={})".format(type(self).__name__, self.rho,
                                               self.epsilon)


class AdaGrad(SGDUpdater):
    def __init__(self, eta=1, epsilon=1e-6):
        if eta <= 0:
            raise ValueError("Learning rate 'eta' must be > 0!")
        if epsilon <= 0:
            raise ValueError("Constant 'epsilon' must be > 0!")
        self.eta, self.epsilon = eta, epsilon
        self.reset()

    def reset(self):
        self.g2_hist = 0

    def __call__(self, x, grad):
        self.g2_hist = self.g2_hist + grad ** 2
        return x - self.eta * grad / (self.epsilon + np.sqrt(self.g2_hist))

    def __repr__(self):
        return "{}(eta={}, epsilon={})".format(type(self).__name__, self.eta,
                                               self.epsilon)


class Momentum(SGDUpdater):
    def __init__(self, rho=0.5, eta=0.01):
        if eta <= 0:
            raise ValueError("Learning rate 'eta' must be > 0!")
        if not 0 <= rho <= 1:
            raise ValueError("Decay rate 'rho' must be between 0 and 1!")
        self.rho, self.eta = rho, eta
        self.reset()

    def reset(self):
        self.dx = 0

    def __call__(self, x, grad):
        self.dx = self.rho * self.dx - self.eta * grad
        return x + self.dx

    def __repr__(self):
        return "{}(rho={}, eta={})".format(type(self).__name__, self.rho,
                                           self.eta)


class Adam(SGDUpdater):
    def __init__(self, alpha=0.01, beta1=0.9, beta2=0.99, epsilon=1e-8):
        self.alpha, self.beta1, self.beta2 = alpha, beta1, beta2
        self.epsilon = epsilon
        self.reset()

    def reset(self):
        self.t = 0
        self.m = None
        self.v = None

    def __call__(self, x, grad):
        self.t += 1
        if self.m is None:
            self.m = np.zeros_like(x)
            self.v = np.zeros_like(x)
        b1, b2 = self.beta1, self.beta2
        self.m = b1 * self.m + (1 - b1) * grad
        self.v = b2 * self.v + (1 - b2) * grad ** 2
        mhat = self.m / (1 - b1 ** self.t)
        vhat = self.v / (1 - b2 ** self.t)
        return x - self.alpha * mhat / (np.sqrt(vhat) + self.epsilon)

    def __repr__(self):
        return "{}(alpha={}, beta1={}, beta2={}, epsilon={})".format(
            type(self).__name__, self.alpha, self.beta1, self.beta2,
            self.epsilon)


def _is_seq(x):
    return isinstance(x, (list, tuple))


def _len_data(data):
    if not _is_seq(data):
        return data.shape[0]
    N = len(data[0])
    for dd in data[1:]:
        if dd.shape[0] != N:
            raise ValueError("Not all data is the same length!")
    return N


def endless_permutations(N, random_state=None):
    """Indices from back-to-back random permutations of range(N)."""
    rs = check_random_state(random_state)
    while True:
        for b in rs.permutation(N):
            yield b


def gen_batch(data, batch_size, maxiter=np.inf, random_state=None):
    """Yield minibatches (lists of row-indexed arrays).

    Draw order matches the reference: a fresh permutation of range(N) is
    taken from ``random_state`` whenever the previous one is exhausted, and a
    batch may straddle two permutations.
    """
    N = _len_data(data)
    rs = check_random_state(random_state)
    perm = np.zeros(0, dtype=int)
    it = 0
    while it < maxiter:
        it += 1
        while len(perm) < batch_size:
            perm = np.concatenate((perm, rs.permutation(N)))
        ind, perm = perm[:batch_size], perm[batch_size:]
        if not _is_seq(data):
            yield (data[ind],)
        else:
            yield [dd[ind] for dd in data]


class SGDRun(object):
    """State of one SGD run, advanced one minibatch at a time (``sgd`` below is
    the loop over it; benchmarks call ``step`` themselves)."""

    def __init__(self, fun, x0, data, args=(), bounds=None, batch_size=10,
                 maxiter=5000, updater=None, eval_obj=False, random_state=None):
        self.fun, self.args, self.eval_obj = fun, args, eval_obj
        self.updater = Adam() if updater is None else updater
        self.updater.reset()
        N = _len_data(data)
        self.x = np.array(x0, copy=True, dtype=float)
        self.lower = self.upper = None
        if bounds is not None:
            if len(bounds) != self.x.shape[0]:
                raise ValueError("The dimension of the bounds does not match x0!")
            self.lower = np.array([-np.inf if b[0] is None else b[0] for b in bounds])
            self.upper = np.array([np.inf if b[1] is None else b[1] for b in bounds])
        self.obj, self.objs, self.norms = None, [], []
        self.batches = gen_batch(data, min(batch_size, N), maxiter, random_state)

    def step(self):
        """One minibatch: objective / gradient, bound-aware truncation, update,
        clip (sgd.py:380-415).  Returns False when the batches are exhausted."""
        batch = next(self.batches, None)
        if batch is None:
            return False
        x = self.x
        if self.eval_obj:
            self.obj, grad = self.fun(x, *chain(batch, self.args))
            self.objs.append(self.obj)
        else:
            grad = self.fun(x, *chain(batch, self.args))
        self.norms.append(np.linalg.norm(grad))
        if self.lower is not None:
            at_lo, at_hi = x <= self.lower, x >= self.upper
            grad[at_lo] = np.minimum(grad[at_lo], 0)
            grad[at_hi] = np.maximum(grad[at_hi], 0)
        x = self.updater(x, grad)
        if self.lower is not None:
            x = np.clip(x, self.lower, self.upper)
        self.x = x
        return True

    def result(self):
        return OptimizeResult(x=self.x, norms=self.norms, message='maxiter reached',
                              fun=self.obj, objs=self.objs)


def sgd(fun, x0, data, args=(), bounds=None, batch_size=10, maxiter=5000,
        updater=None, eval_obj=False, random_state=None):
    """Minimise ``fun`` by SGD over minibatches of ``data`` (sgd.py:311-425):
    bound-aware gradient truncation, update, clip; returns an
    ``OptimizeResult`` with ``x``, ``norms``, ``objs``, ``fun``."""
    run = SGDRun(fun, x0, data, args=args, bounds=bounds, batch_size=batch_size,
                 maxiter=maxiter, updater=updater, eval_obj=eval_obj,
                 random_state=random_state)
    while run.step():
        pass
    return run.result()
