"""
NumPy-backed stand-in for the small slice of the JAX API that GC-SLAM's LiDAR
evidence operators use.  TEST INFRASTRUCTURE ONLY.

Why this exists: the reference (whabacivch/GC-SLAM) is Python + JAX, and JAX is
not installed in the build container (no network).  With this directory first on
``sys.path`` the reference's *own, unmodified* operator source files import and
run on NumPy float64, so ``tests/golden/make_golden.py`` can generate golden
vectors from the reference's code rather than from our restatement of it.

What is substituted: only the L0 array runtime (XLA -> NumPy/LAPACK).  Semantics
mirrored on purpose: stable ``argsort``/``lax.sort`` (``num_keys=1``),
functional ``.at[].set/add/min/max``, ``vmap`` as a stacked Python loop,
``jit`` as identity, ``nn.softmax`` = exp(x-max)/sum, ``nn.sigmoid`` =
1/(1+exp(-x)), ``linalg.eigh`` ascending, ``linalg.svd`` descending.

Nothing in the product path imports this package.
"""

from __future__ import annotations

import numpy as _np

from . import numpy  # noqa: F401  (jax.numpy)
from . import lax, nn, scipy  # noqa: F401
from .numpy import Array, _wrap

__version__ = "0.0-numpy-shim"


def jit(fun=None, **_kwargs):
    """Identity; supports both ``@jax.jit`` and ``@jax.jit(static_argnames=...)``."""
    if fun is None:
        return lambda f: f
    return fun


def _tree_stack(items):
    first = items[0]
    if isinstance(first, tuple) and hasattr(first, "_fields"):
        return type(first)(*[_tree_stack([it[i] for it in items]) for i in range(len(first))])
    if isinstance(first, (tuple, list)):
        return type(first)(_tree_stack([it[i] for it in items]) for i in range(len(first)))
    return _wrap(_np.stack([_np.asarray(it) for it in items], axis=0))


def vmap(fun, in_axes=0, out_axes=0):
    """Batched map as a plain Python loop over the mapped axis, outputs stacked on axis 0."""
    if out_axes != 0:
        raise NotImplementedError("jax shim: out_axes != 0")

    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _np.shape(a)[ax]
                break
        if n is None:
            raise ValueError("jax shim vmap: no mapped argument")
        outs = []
        for i in range(int(n)):
            call = [
                (_wrap(_np.take(_np.asarray(a), i, axis=ax)) if ax is not None else a)
                for a, ax in zip(args, axes)
            ]
            outs.append(fun(*call))
        return _tree_stack(outs)

    return mapped


def device_get(x):
    return x


def devices(*_a, **_k):
    return ["numpy-shim-cpu"]


class _Config:
    def update(self, *_a, **_k):
        return None


config = _Config()
