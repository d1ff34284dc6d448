"""NumPy-backed stand-in for the parts of `jax` the reference uses (test infrastructure, see ../README.md)."""
import numpy as _np

from . import numpy  # noqa: F401
from . import random, nn, scipy  # noqa: F401
from .numpy import JArray as Array, _wrap

_LEAF_CONTAINERS = (tuple, list)


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()


def _tree_map(f, t):
    if isinstance(t, _LEAF_CONTAINERS):
        return type(t)(_tree_map(f, v) for v in t)
    return f(t)


def _tree_leaves(t):
    if isinstance(t, _LEAF_CONTAINERS):
        out = []
        for v in t:
            out += _tree_leaves(v)
        return out
    return [t]


def _tree_stack(items):
    first = items[0]
    if isinstance(first, _LEAF_CONTAINERS):
        return type(first)(_tree_stack([it[k] for it in items]) for k in range(len(first)))
    return _wrap(_np.stack([_np.asarray(v) for v in items]))


def jit(fun=None, **kwargs):
    if fun is None:
        return lambda f: f
    return fun


def vmap(fun, in_axes=0, out_axes=0):
    """jax.vmap as a Python loop over the mapped axis; constant outputs are broadcast (stacked), as in JAX."""
    assert out_axes == 0

    def mapped(*args, **kwargs):
        axes = (in_axes,) * len(args) if (in_axes is None or isinstance(in_axes, int)) else tuple(in_axes)
        assert len(axes) == len(args), (len(axes), len(args))
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _np.asarray(_tree_leaves(a)[0]).shape[ax]
                break
        if n is None:
            n = _np.asarray(_tree_leaves(list(kwargs.values()))[0]).shape[0]
        outs = []
        for i in range(n):
            ai = [a if ax is None else _tree_map(lambda x, ax=ax: _wrap(_np.take(_np.asarray(x), i, axis=ax)), a)
                  for a, ax in zip(args, axes)]
            ki = {k: _tree_map(lambda x: _wrap(_np.asarray(x)[i]), v) for k, v in kwargs.items()}
            outs.append(fun(*ai, **ki))
        return _tree_stack(outs)

    return mapped
