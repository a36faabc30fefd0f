"""jax.lax stand-in (NumPy). TEST INFRASTRUCTURE ONLY."""
import numpy as _np
from .numpy import _wrap


def fori_loop(lower, upper, body_fun, init_val):
    val = init_val
    for i in range(int(lower), int(upper)):
        val = body_fun(i, val)
    return val


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def scan(f, init, xs, length=None):
    carry = init
    ys = []
    n = length if xs is None else len(xs[0] if isinstance(xs, (tuple, list)) else xs)
    for i in range(int(n)):
        x = None if xs is None else (
            type(xs)(_wrap(_np.asarray(a)[i]) for a in xs) if isinstance(xs, (tuple, list)) else _wrap(_np.asarray(xs)[i])
        )
        carry, y = f(carry, x)
        ys.append(y)
    from . import _tree_stack
    return carry, (_tree_stack(ys) if ys and ys[0] is not None else None)


def sort(operand, dimension=-1, is_stable=True, num_keys=1):
    """Stable multi-operand sort; only the first ``num_keys`` operands are keys (JAX default 1)."""
    if not isinstance(operand, (tuple, list)):
        return _wrap(_np.sort(_np.asarray(operand), axis=dimension, kind="stable"))
    ops = [_np.asarray(o) for o in operand]
    if num_keys == 1:
        order = _np.argsort(ops[0], axis=dimension, kind="stable")
    else:
        # lexicographic on the first num_keys operands (first is most significant)
        if ops[0].ndim != 2 or dimension not in (1, -1):
            raise NotImplementedError("jax shim: multi-key sort only for 2-D, last axis")
        order = _np.empty(ops[0].shape, dtype=_np.int64)
        for r in range(ops[0].shape[0]):
            order[r] = _np.lexsort(tuple(ops[k][r] for k in reversed(range(num_keys))))
    return tuple(_wrap(_np.take_along_axis(o, order, axis=dimension)) for o in ops)


def _clamp_starts(shape, sizes, starts):
    # XLA clamps start indices so that the slice stays inside the operand
    return [int(min(max(int(s), 0), d - z)) for s, d, z in zip(starts, shape, sizes)]


def dynamic_update_slice(operand, update, start_indices):
    out = _np.array(_np.asarray(operand), copy=True)
    upd = _np.asarray(update)
    st = _clamp_starts(out.shape, upd.shape, start_indices)
    out[tuple(slice(s, s + z) for s, z in zip(st, upd.shape))] = upd
    return _wrap(out)


def dynamic_slice(operand, start_indices, slice_sizes):
    a = _np.asarray(operand)
    st = _clamp_starts(a.shape, slice_sizes, start_indices)
    return _wrap(a[tuple(slice(s, s + int(z)) for s, z in zip(st, slice_sizes))].copy())


def select(pred, on_true, on_false):
    return _wrap(_np.where(_np.asarray(pred), _np.asarray(on_true), _np.asarray(on_false)))


def stop_gradient(x):
    return x


def while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while bool(cond_fun(val)):
        val = body_fun(val)
    return val
