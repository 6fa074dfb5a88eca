"""Optimiser glue: nested ``Parameter`` structures <-> flat vectors, log-space
warping of ``Positive`` parameters and random restarts.

Host-side equivalent of revrand/optimize/decorators.py:24-617 (the *caller*
of the hot path): ``structured_minimizer``, ``logtrick_minimizer``,
``structured_sgd``, ``logtrick_sgd`` keep their call signatures so
``structured_minimizer(logtrick_minimizer(minimize))`` composes as before.
All of it is O(#hyper-parameters) bookkeeping around the GPU objective.
"""

from __future__ import annotations

import logging
from functools import wraps
from itertools import chain

import numpy as np

from ..btypes import Bound, Positive
from .sgd import gen_batch

log = logging.getLogger(__name__)

MINPOS = 1e-100
MAXPOS = np.sqrt(np.finfo(float).max)
LOGMINPOS = np.log(MINPOS)
EXPMAX = np.log(MAXPOS)


def _is_seq(x):
    return isinstance(x, (list, tuple)) or hasattr(x, "__next__")


class Layout(object):
    """Shape tree of a nested list of Parameters / values."""

    def __init__(self, tree):
        self.tree = tree  # nested lists of shape tuples

    @classmethod
    def of_parameters(cls, parameters):
        def walk(p):
            if _is_seq(p) and len(p) > 0:
                return [walk(q) for q in p]
            return tuple(p.shape)
        return cls(walk(parameters))

    @staticmethod
    def _size(node):
        if isinstance(node, tuple):
            return int(np.prod(node, dtype=int))
        return int(sum(Layout._size(n) for n in node))

    def unflatten(self, vec):
        """Flat vector -> nested values: python floats for shape (), ``[]``
        for the empty parameter, arrays otherwise (as utils/base.py:280-290)."""
        def build(node, v):
            if isinstance(node, tuple):
                if node == ():
                    return float(np.asarray(v).item())
                if node == (0,):
                    return []
                return np.reshape(v, node)
            out, pos = [], 0
            for n in node:
                sz = Layout._size(n)
                out.append(build(n, v[pos:pos + sz]))
                pos += sz
            return out
        return build(self.tree, np.asarray(vec))


def flatten_values(tree):
    """Nested lists of scalars / arrays / [] -> one flat float vector."""
    if _is_seq(tree):
        tree = list(tree)
        if len(tree) == 0:
            return np.zeros(0)
        return np.concatenate([flatten_values(t) for t in tree])
    return np.ravel(np.asarray(tree, dtype=float))


def _map_params(fn, parameters):
    if _is_seq(parameters):
        return [_map_params(fn, p) for p in parameters]
    return fn(parameters)


def _flat_bounds(parameters):
    out = []

    def walk(p):
        if _is_seq(p):
            for q in p:
                walk(q)
        else:
            out.extend([p.bounds] * int(np.prod(p.shape, dtype=int)))
    walk(parameters)
    return out


def _random_starts(fun, parameters, jac, args, nstarts, random_state,
                   data_gen=None):
    """Best of ``nstarts`` draws from the parameters' distributions
    (decorators.py:541-583); every candidate costs one objective call.  Only
    the objective VALUE of a candidate is used, so an objective may offer a
    cheaper ``fun.value_only(*params)`` (no gradients) for this phase."""
    if nstarts < 1:
        raise ValueError("nstarts has to be greater than or equal to 1")
    flags = flatten_values(_map_params(lambda p: float(p.is_random), parameters))
    if not np.any(flags):
        log.info("No random parameters, not doing any random starts")
        return _map_params(lambda p: p.value, parameters)
    log.info("Evaluating random starts...")
    best_obj, best = None, None
    value_only = getattr(fun, "value_only", None) if data_gen is None else None
    value_batch = getattr(fun, "value_only_batch", None) if data_gen is None else None
    if value_batch is not None:
        # the draws do not depend on the objective: take them all first (same
        # generator order as the loop below), then let the objective pipeline the
        # independent evaluations
        cands = [_map_params(lambda p: p.rvs(random_state), parameters)
                 for _ in range(nstarts)]
        objs = value_batch([tuple(chain(c, args)) for c in cands])
        for cand, obj in zip(cands, objs):
            if best_obj is None or obj < best_obj:
                best_obj, best = obj, cand
        log.info("Best start found with objective = {}".format(best_obj))
        return flatten_values(best)
    for _ in range(nstarts):
        batch = next(data_gen) if data_gen else ()
        cand = _map_params(lambda p: p.rvs(random_state), parameters)
        if value_only is not None:
            obj = value_only(*chain(cand, args))
        else:
            out = fun(*chain(cand, batch, args))
            obj = out[0] if jac is True else out
        if best_obj is None or obj < best_obj:
            best_obj, best = obj, cand
    log.info("Best start found with objective = {}".format(best_obj))
    return flatten_values(best)


def structured_minimizer(minimizer):
    """Let ``minimizer(fun, x0, ...)`` work on nested ``Parameter`` lists."""
    @wraps(minimizer)
    def new_minimizer(fun, parameters, jac=True, args=(), nstarts=0,
                      random_state=None, **minimizer_kwargs):
        layout = Layout.of_parameters(parameters)
        x0 = flatten_values(_map_params(lambda p: p.rvs(random_state),
                                        parameters))
        bounds = _flat_bounds(parameters)
        if nstarts > 0:
            x0 = flatten_values(_random_starts(fun, parameters, jac, args,
                                               nstarts, random_state))

        def flat_fun(x, *a, **kw):
            out = fun(*(tuple(layout.unflatten(x)) + a), **kw)
            if (not callable(jac)) and bool(jac):
                return out[0], flatten_values(out[1])
            return out

        flat_jac = jac
        if callable(jac):
            def flat_jac(x, *a, **kw):
                return flatten_values(jac(*(tuple(layout.unflatten(x)) + a), **kw))

        res = minimizer(flat_fun, x0, jac=flat_jac, args=args, bounds=bounds,
                        **minimizer_kwargs)
        res['x'] = tuple(layout.unflatten(res['x']))
        if bool(jac) and 'jac' in res:
            res['jac'] = tuple(layout.unflatten(res['jac']))
        return res
    return new_minimizer


def structured_sgd(sgd):
    """Let ``sgd(fun, x0, data, ...)`` work on nested ``Parameter`` lists."""
    @wraps(sgd)
    def new_sgd(fun, parameters, data, eval_obj=False, batch_size=10, args=(),
                random_state=None, nstarts=100, **sgd_kwargs):
        layout = Layout.of_parameters(parameters)
        x0 = flatten_values(_map_params(lambda p: p.rvs(None), parameters))
        bounds = _flat_bounds(parameters)
        if eval_obj and nstarts > 0:
            data_gen = gen_batch(data, batch_size, random_state=random_state)
            x0 = flatten_values(_random_starts(fun, parameters, True, args,
                                               nstarts, random_state, data_gen))

        def flat_fun(x, *a, **kw):
            out = fun(*(tuple(layout.unflatten(x)) + a), **kw)
            if bool(eval_obj):
                return out[0], flatten_values(out[1])
            return flatten_values(out)

        res = sgd(flat_fun, x0, data=data, bounds=bounds, args=args,
                  eval_obj=eval_obj, random_state=random_state,
                  batch_size=batch_size, **sgd_kwargs)
        res['x'] = tuple(layout.unflatten(res['x']))
        return res
    return new_sgd


class _LogWarp(object):
    """x -> log x on the coordinates whose bound is ``Positive``."""

    def __init__(self, bounds):
        self.pos = np.array([isinstance(b, Positive) for b in bounds], dtype=bool)
        self.bounds = []
        for b, ispos in zip(bounds, self.pos):
            if ispos:
                up = EXPMAX if b.upper is None else np.log(b.upper)
                self.bounds.append(Bound(lower=LOGMINPOS, upper=up))
            else:
                self.bounds.append(b)

    def fwd(self, x):
        x = np.array(x, dtype=float, copy=True)
        x[self.pos] = np.log(x[self.pos])
        return x

    def inv(self, z):
        z = np.array(z, dtype=float, copy=True)
        z[self.pos] = np.exp(z[self.pos])
        return z

    def grad(self, g, z):
        g = np.array(g, dtype=float, copy=True)
        g[self.pos] = g[self.pos] * np.exp(z[self.pos])
        return g


def logtrick_minimizer(minimizer):
    """Optimise ``Positive``-bounded coordinates in log-space
    (decorators.py:255-326)."""
    @wraps(minimizer)
    def new_minimizer(fun, x0, jac=True, bounds=None, **minimizer_kwargs):
        if bounds is None:
            return minimizer(fun, x0, jac=jac, bounds=bounds, **minimizer_kwargs)
        warp = _LogWarp(bounds)
        with_grad = (not callable(jac)) and bool(jac)

        def new_fun(z, *a, **kw):
            out = fun(warp.inv(z), *a, **kw)
            if with_grad:
                return out[0], warp.grad(out[1], z)
            return out

        new_jac = jac
        if callable(jac):
            def new_jac(z, *a, **kw):
                return warp.grad(jac(warp.inv(z), *a, **kw), z)

        res = minimizer(new_fun, warp.fwd(x0), jac=new_jac, bounds=warp.bounds,
                        **minimizer_kwargs)
        res['x'] = warp.inv(res['x'])
        return res
    return new_minimizer


def logtrick_sgd(sgd):
    """Log-space warping for SGD (decorators.py:329-403)."""
    @wraps(sgd)
    def new_sgd(fun, x0, data, bounds=None, eval_obj=False, **sgd_kwargs):
        if bounds is None:
            return sgd(fun, x0, data, bounds=bounds, eval_obj=eval_obj,
                       **sgd_kwargs)
        warp = _LogWarp(bounds)

        def new_fun(z, *a, **kw):
            out = fun(warp.inv(z), *a, **kw)
            if bool(eval_obj):
                return out[0], warp.grad(out[1], z)
            return warp.grad(out, z)

        res = sgd(new_fun, warp.fwd(x0), data, bounds=warp.bounds,
                  eval_obj=eval_obj, **sgd_kwargs)
        res['x'] = warp.inv(res['x'])
        return res
    return new_sgd
