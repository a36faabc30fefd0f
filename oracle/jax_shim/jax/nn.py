"""jax.nn stand-in (NumPy). TEST INFRASTRUCTURE ONLY."""
import numpy as _np
from .numpy import _wrap


def softmax(x, axis=-1):
    x = _np.asarray(x)
    x_max = _np.max(x, axis=axis, keepdims=True)
    un = _np.exp(x - x_max)
    return _wrap(un / _np.sum(un, axis=axis, keepdims=True))


def sigmoid(x):
    x = _np.asarray(x)
    with _np.errstate(over="ignore"):
        return _wrap(1.0 / (1.0 + _np.exp(-x)))
