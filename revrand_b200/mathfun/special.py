"""Stable special functions used around the likelihoods (host side).

Counterparts of revrand/mathfun/special.py:22-142; the GPU step kernel has its
own fp32 versions of softplus / expit (csrc/rr_glm.cu).
"""

import numpy as np

EPS = np.finfo(float).eps
TINY = np.finfo(float).tiny
SMALL = 1e-100
LOGTINY = np.log(TINY)


def logsumexp(X, axis=0):
    """log(sum(exp(X))) along ``axis`` without overflow."""
    X = np.asarray(X, dtype=float)
    mx = X.max(axis=axis, keepdims=True)
    out = np.log(np.exp(X - mx).sum(axis=axis, keepdims=True)) + mx
    return np.squeeze(out, axis=axis)


def softmax(X, axis=0):
    """exp(X) normalised along ``axis`` (2-D arrays)."""
    if axis not in (0, 1):
        raise ValueError("This only works on 2D arrays for now.")
    X = np.asarray(X, dtype=float)
    return np.exp(X - np.expand_dims(logsumexp(X, axis=axis), axis))


def softplus(X):
    """log(1 + exp(X)), stable for large |X|."""
    if np.isscalar(X):
        return float(np.logaddexp(0.0, X))
    X = np.asarray(X, dtype=float)
    if X.ndim > 2:
        raise ValueError("This only works on up to 2D arrays.")
    return np.logaddexp(0.0, X)


def safelog(x, min_x=TINY):
    return np.log(np.maximum(x, min_x))


def safesoftplus(x, min_x=SMALL):
    """softplus floored at ``min_x`` so it can be divided by."""
    return np.maximum(softplus(x), min_x)
