"""Stand-in for jax.numpy: numpy itself, plus the two keyword spellings that differ."""
import numpy as _np
from numpy import *  # noqa: F401,F403
from numpy import fft  # noqa: F401

ndarray = _np.ndarray
float32 = _np.float32
int32 = _np.int32


def clip(a, a_min=None, a_max=None):
    return _np.clip(a, a_min, a_max)


def array(obj, dtype=None):
    return _np.array(obj, dtype=dtype)
