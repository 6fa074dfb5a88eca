"""revrand_b200: a Blackwell-native random-feature Bayesian linear / GLM
engine behind the revrand API.

>>> from revrand_b200 import StandardLinearModel, GeneralizedLinearModel
>>> from revrand_b200.basis_functions import RandomMatern32, LinearBasis

The feature, likelihood and sufficient-statistics passes run in
``lib/librevrand_b200.so`` (hand-written sm_100a CUDA, C-ABI in
``include/revrand_b200.h``); there is no CPU fallback.
"""

from . import basis_functions, btypes, likelihoods, metrics, optimize  # noqa
from .glm import GeneralisedLinearModel, GeneralizedLinearModel
from .slm import StandardLinearModel
from .btypes import Bound, Parameter, Positive

__all__ = ['StandardLinearModel', 'GeneralizedLinearModel',
           'GeneralisedLinearModel', 'Parameter', 'Bound', 'Positive',
           'basis_functions', 'btypes', 'likelihoods', 'metrics', 'optimize']
__version__ = '0.1.0'
