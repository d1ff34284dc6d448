"""equinox stand-in (test infrastructure, see ../README.md): `Module` is a plain base class, `filter_jit` calls the
function eagerly after turning NumPy array arguments into jax-style arrays (what jit's tracing does)."""
import numpy as _np
from jax.numpy import _wrap


class Module:
    pass


def _lift_tree(t):
    if isinstance(t, (tuple, list)):
        return type(t)(_lift_tree(v) for v in t)
    if isinstance(t, (_np.ndarray, _np.generic)):
        return _wrap(t)
    return t


def filter_jit(fun=None, **kwargs):
    if fun is None:
        return lambda f: filter_jit(f)

    def call(*a, **k):
        return fun(*[_lift_tree(v) for v in a], **{n: _lift_tree(v) for n, v in k.items()})

    return call
