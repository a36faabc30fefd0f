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


def softplus(x):
    """log(1 + e^x), the overflow-safe form jax.nn.softplus uses (logaddexp(x, 0))."""
    return _wrap(_np.logaddexp(_np.asarray(x, dtype=_np.float64), 0.0))


def relu(x):
    return _wrap(_np.maximum(_np.asarray(x), 0))
