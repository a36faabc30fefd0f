"""jax.numpy stand-in on NumPy (see ``jax/__init__.py`` of this shim). TEST INFRASTRUCTURE ONLY."""

from __future__ import annotations

import types as _types

import numpy as _np


class _AtIndexer:
    __slots__ = ("_a", "_idx")

    def __init__(self, a, idx):
        self._a = a
        self._idx = idx

    @staticmethod
    def _plain(idx):
        if isinstance(idx, tuple):
            return tuple(_np.asarray(i) if isinstance(i, _np.ndarray) else i for i in idx)
        return _np.asarray(idx) if isinstance(idx, _np.ndarray) else idx

    def _copy(self):
        return _np.array(self._a, copy=True)

    def set(self, v):
        out = self._copy()
        out[self._plain(self._idx)] = _np.asarray(v)
        return _wrap(out)

    def add(self, v):
        out = self._copy()
        _np.add.at(out, self._plain(self._idx), _np.asarray(v))
        return _wrap(out)

    def multiply(self, v):
        out = self._copy()
        _np.multiply.at(out, self._plain(self._idx), _np.asarray(v))
        return _wrap(out)

    def min(self, v):
        out = self._copy()
        _np.minimum.at(out, self._plain(self._idx), _np.asarray(v))
        return _wrap(out)

    def max(self, v):
        out = self._copy()
        _np.maximum.at(out, self._plain(self._idx), _np.asarray(v))
        return _wrap(out)

    def get(self):
        return _wrap(_np.asarray(self._a)[self._plain(self._idx)])


class _At:
    __slots__ = ("_a",)

    def __init__(self, a):
        self._a = a

    def __getitem__(self, idx):
        return _AtIndexer(self._a, idx)


class Array(_np.ndarray):
    """ndarray with the functional ``.at[...]`` update API and ``block_until_ready``."""

    @property
    def at(self):
        return _At(self)

    def block_until_ready(self):
        return self


def _wrap(x):
    if isinstance(x, Array):
        return x
    if isinstance(x, _np.ndarray):
        return x.view(Array)
    if isinstance(x, tuple) and hasattr(x, "_fields"):
        return type(x)(*[_wrap(i) for i in x])
    if isinstance(x, tuple):
        return tuple(_wrap(i) for i in x)
    if isinstance(x, list):
        return [_wrap(i) for i in x]
    return x


def _wrapping(fn):
    def inner(*a, **k):
        return _wrap(fn(*a, **k))

    inner.__name__ = getattr(fn, "__name__", "wrapped")
    return inner


ndarray = _np.ndarray
pi = _np.pi
inf = _np.inf
nan = _np.nan
e = _np.e
newaxis = None
float64 = _np.float64
float32 = _np.float32
int32 = _np.int32
int64 = _np.int64
uint8 = _np.uint8
uint32 = _np.uint32
bool_ = _np.bool_
finfo = _np.finfo
iinfo = _np.iinfo


def asarray(x, dtype=None):
    return _wrap(_np.asarray(x, dtype=dtype))


def array(x, dtype=None, copy=True):
    return _wrap(_np.array(x, dtype=dtype, copy=True))


def where(condition, x=None, y=None, size=None, fill_value=None):
    """jnp.where; the one-argument form takes `size` (fixed-length nonzero: truncate, or pad with fill_value / 0)."""
    if x is not None or y is not None:
        return _wrap(_np.where(_np.asarray(condition), _np.asarray(x), _np.asarray(y)))
    idx = _np.nonzero(_np.asarray(condition))
    if size is not None:
        out = []
        for a in idx:
            a = a[:size]
            if a.shape[0] < size:
                a = _np.concatenate([a, _np.full(size - a.shape[0], 0 if fill_value is None else fill_value, dtype=a.dtype)])
            out.append(a)
        idx = tuple(out)
    return tuple(_wrap(a) for a in idx)


def argsort(a, axis=-1, **_k):
    return _wrap(_np.argsort(_np.asarray(a), axis=axis, kind="stable"))


def sort(a, axis=-1, **_k):
    return _wrap(_np.sort(_np.asarray(a), axis=axis, kind="stable"))


def unique(a, **k):
    return _wrap(_np.unique(_np.asarray(a), **k))


def _ns_wrap(mod, names=None):
    ns = _types.SimpleNamespace()
    for name in dir(mod):
        if name.startswith("_"):
            continue
        obj = getattr(mod, name)
        setattr(ns, name, _wrapping(obj) if callable(obj) and not isinstance(obj, type) else obj)
    return ns


linalg = _ns_wrap(_np.linalg)


def _cholesky(a, symmetrize_input=True, **_k):
    """jnp.linalg.cholesky symmetrises its input by default (symmetrize_input=True); NumPy reads the lower triangle."""
    a = _np.asarray(a)
    if symmetrize_input:
        a = 0.5 * (a + _np.swapaxes(a, -1, -2))
    return _wrap(_np.linalg.cholesky(a))


linalg.cholesky = _cholesky


def __getattr__(name):
    obj = getattr(_np, name)
    if callable(obj) and not isinstance(obj, type):
        return _wrapping(obj)
    return obj
