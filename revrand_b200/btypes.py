"""Bounded / distribution-valued parameter types (host-side metadata).

Mirrors the user-visible constructors of revrand/btypes.py:83-348
(``Bound``, ``Positive``, ``Parameter``) so existing model definitions keep
working; this is O(#hyper-parameters) host bookkeeping, not part of the GPU
hot path.
"""

from __future__ import annotations

import numpy as np
from sklearn.utils import check_random_state


class Bound(tuple):
    """(lower, upper) interval; ``None`` means unbounded on that side.

    As in the reference (btypes.py:39-43) a bound equal to 0 is treated as
    "no bound" by :meth:`check`.
    """

    def __new__(cls, lower=None, upper=None):
        if lower is not None and upper is not None and lower > upper:
            raise ValueError("lower bound cannot be greater than upper bound!")
        return super(Bound, cls).__new__(cls, (lower, upper))

    lower = property(lambda self: self[0])
    upper = property(lambda self: self[1])

    def __getnewargs__(self):
        return (self.lower, self.upper)

    def check(self, value):
        """True when every element of ``value`` lies inside the interval."""
        value = np.asarray(value)
        if self.lower and np.any(value < self.lower):
            return False
        if self.upper and np.any(value > self.upper):
            return False
        return True

    def clip(self, value):
        if not self.lower and not self.upper:
            return value
        return np.clip(value, self.lower, self.upper)

    def __repr__(self):
        return "{}(lower={}, upper={})".format(type(self).__name__, self.lower,
                                               self.upper)


class Positive(Bound):
    """Strictly positive interval (1e-14, upper); marks parameters that the
    optimiser glue moves in log-space."""

    def __new__(cls, upper=None):
        lower = 1e-14
        if upper is not None and lower > upper:
            raise ValueError("Upper bound must be greater than {}".format(lower))
        return tuple.__new__(cls, (lower, upper))

    def __getnewargs__(self):
        return (self.upper,)

    def __repr__(self):
        return "{}(upper={})".format(type(self).__name__, self.upper)


class Parameter(object):
    """A (possibly random) initial value with bounds.

    ``value`` may be a scalar / array, or a frozen ``scipy.stats``
    distribution (anything with ``rvs``), in which case ``shape`` gives the
    shape of the draws and the expected value (clipped to the bounds) is the
    nominal value.  ``Parameter()`` is the "no parameter" marker.
    """

    def __init__(self, value=None, bounds=Bound(), shape=()):
        if value is None:
            value = []
        if hasattr(value, "rvs"):
            self.dist = value
            self.shape = shape
            mean = bounds.clip(value.mean())
            self.value = mean if shape == () else mean * np.ones(shape)
        else:
            if np.any(value) and not bounds.check(value):
                raise ValueError("Value not within bounds!")
            self.dist = None
            self.value = value
            self.shape = np.shape(value)
        self.bounds = bounds

    def rvs(self, random_state=None):
        """A draw clipped to the bounds (the fixed value if not random)."""
        if self.dist is None:
            return self.value
        rs = check_random_state(random_state)
        return self.bounds.clip(self.dist.rvs(size=self.shape, random_state=rs))

    @property
    def has_value(self):
        return self.shape != (0,)

    @property
    def is_random(self):
        return self.dist is not None

    @property
    def is_scalar(self):
        return self.has_value and self.shape == ()

    def __repr__(self):
        return "{}(value={}, bounds={}, shape={})".format(
            type(self).__name__, self.value, self.bounds, self.shape)
