"""Optimiser glue around the GPU objectives (host side)."""

from .sgd import (sgd, gen_batch, endless_permutations, SGDUpdater, AdaDelta,
                  AdaGrad, Momentum, Adam)
from .structured import (structured_minimizer, structured_sgd,
                         logtrick_minimizer, logtrick_sgd, Layout,
                         flatten_values)

__all__ = ['sgd', 'gen_batch', 'endless_permutations', 'SGDUpdater',
           'AdaDelta', 'AdaGrad', 'Momentum', 'Adam', 'structured_minimizer',
           'structured_sgd', 'logtrick_minimizer', 'logtrick_sgd', 'Layout',
           'flatten_values']
