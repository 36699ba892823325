"""jax.numpy -> numpy (float64 throughout, as with jax_enable_x64)."""
from numpy import *  # noqa: F401,F403
from numpy import linalg, ndarray  # noqa: F401
import numpy as _np


def isscalar(x):
    return _np.isscalar(x)
