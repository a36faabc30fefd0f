import numpy as _np
import scipy.linalg as _sl
from ..numpy import _wrap


def solve_triangular(a, b, lower=False, trans=0, **_k):
    return _wrap(_sl.solve_triangular(_np.asarray(a), _np.asarray(b), lower=lower, trans=trans))


def cho_factor(a, lower=False):
    c, low = _sl.cho_factor(_np.asarray(a), lower=lower)
    return _wrap(c), low


def cho_solve(c_and_lower, b):
    c, low = c_and_lower
    return _wrap(_sl.cho_solve((_np.asarray(c), low), _np.asarray(b)))
