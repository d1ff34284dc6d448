import numpy as _np
from .numpy import _wrap


def softmax(x, axis=-1):
    """jax.nn.softmax: exp(x - max) / sum(exp(x - max))"""
    x = _np.asarray(x, dtype=_np.float64)
    un = _np.exp(x - _np.max(x, axis=axis, keepdims=True))
    return _wrap(un / _np.sum(un, axis=axis, keepdims=True))
