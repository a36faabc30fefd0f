import numpy as _np
import scipy.special as _sp
from ..numpy import _wrap


def logsumexp(a, axis=None, **k):
    return _wrap(_np.asarray(_sp.logsumexp(_np.asarray(a), axis=axis, **k)))


def i0e(x):
    return _wrap(_np.asarray(_sp.i0e(_np.asarray(x))))


def i1e(x):
    return _wrap(_np.asarray(_sp.i1e(_np.asarray(x))))
