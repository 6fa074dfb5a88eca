"""Dense helpers with the reference's semantics (host side, float64).

``solve_posdef`` here is the small-matrix host version of
revrand/mathfun/linalg.py:84-125; the model posterior solve runs on the GPU in
``_engine.solve_posterior`` with the same Cholesky -> clamped-SVD fallback.
"""

import numpy as np
from scipy.linalg import LinAlgError, cho_solve, cholesky, svd

CHOLTHRESH = 1e-5


def cho_log_det(L):
    """log det A from a Cholesky factor of A."""
    return 2 * np.sum(np.log(L.diagonal()))


def svd_log_det(s):
    return np.sum(np.log(s))


def svd_solve(U, s, V, b, s_tol=1e-15):
    """Solve A x = b from A = U diag(s) V with singular values clamped."""
    inv = 1. / np.maximum(s, s_tol)
    return (U * inv[None, :]).dot(V.dot(b)) if np.ndim(b) == 1 or \
        b.shape[1] < U.shape[0] else (U * inv[None, :]).dot(V).dot(b)


def solve_posdef(A, b):
    """(A^-1 b, log det A) by Cholesky, falling back to an SVD solve when the
    factorisation fails or is numerically unstable."""
    try:
        L = cholesky(A, lower=False)
        if np.any(L.diagonal() < CHOLTHRESH):
            raise LinAlgError("Unstable cholesky factor detected")
        return cho_solve((L, False), b), cho_log_det(L)
    except LinAlgError:
        U, s, V = svd(A)
        return svd_solve(U, s, V, b), svd_log_det(s)


def hadamard(Y, ordering=True):
    """Row-wise fast Walsh-Hadamard transform scaled by 1/n.

    ``ordering=False`` gives natural (Hadamard) order -- what the FastFood
    basis uses and what the GPU butterfly in csrc/rr_features.cu computes;
    ``ordering=True`` returns sequency order.
    """
    Y = np.array(Y, dtype=float)
    nv, n = Y.shape
    if n & (n - 1):
        raise AssertionError("length must be a power of two")
    h = 1
    while h < n:
        Y = Y.reshape(nv, n // (2 * h), 2, h)
        Y = np.stack((Y[:, :, 0] + Y[:, :, 1], Y[:, :, 0] - Y[:, :, 1]), axis=2)
        Y = Y.reshape(nv, n) / 2.
        h *= 2
    if ordering:
        idx = np.arange(n)
        gray = idx ^ (idx >> 1)
        bits = max(n.bit_length() - 1, 0)
        rev = np.zeros(n, dtype=int)
        for bpos in range(bits):
            rev |= ((gray >> bpos) & 1) << (bits - 1 - bpos)
        Y = Y[:, rev]
    return Y
