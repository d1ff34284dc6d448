"""jax.numpy stand-in: NumPy functions returning `JArray`s (ndarray subclass with `.at[...]` updates and
JAX's clamped out-of-range integer gathers)."""
import numpy as _np

pi, inf, nan, e = _np.pi, _np.inf, _np.nan, _np.e
int32, int64, float64, float32, uint32, bool_ = _np.int32, _np.int64, _np.float64, _np.float32, _np.uint32, _np.bool_


class _AtIndex:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, value):
        out = _np.array(self.arr, copy=True)
        out[self.idx] = value
        return out.view(JArray)

    def add(self, value):
        out = _np.array(self.arr, copy=True)
        out[self.idx] += value
        return out.view(JArray)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIndex(self.arr, idx)


class JArray(_np.ndarray):
    @property
    def at(self):
        return _At(self)

    def __getitem__(self, idx):
        # x[int_array]: JAX wraps negative indices and CLAMPS out-of-range ones for retrieval
        if isinstance(idx, _np.ndarray) and idx.dtype.kind in "iu" and self.ndim >= 1:
            n = self.shape[0]
            idx = _np.asarray(idx)
            idx = _np.clip(_np.where(idx < 0, idx + n, idx), 0, n - 1)
        out = _np.ndarray.__getitem__(self, idx)
        return out


def _wrap(x):
    if isinstance(x, (tuple, list)):
        return type(x)(_wrap(v) for v in x)
    if isinstance(x, _np.ndarray):
        return x.view(JArray)
    if isinstance(x, (_np.generic,)):
        return _np.asarray(x).view(JArray)
    return x


def _lift(f):
    def g(*a, **k):
        return _wrap(f(*a, **k))
    g.__name__ = getattr(f, "__name__", "f")
    return g


def array(x, dtype=None, **k):
    return _np.array(x, dtype=dtype).view(JArray)


def asarray(x, dtype=None, **k):
    return _np.asarray(x, dtype=dtype).view(JArray)


def searchsorted(a, v, side="left", **k):
    return _wrap(_np.searchsorted(_np.asarray(a), _np.asarray(v), side=side))


def clip(x, min=None, max=None, a_min=None, a_max=None):
    lo = min if min is not None else a_min
    hi = max if max is not None else a_max
    return _wrap(_np.clip(_np.asarray(x), lo, hi))


class _Linalg:
    def __getattr__(self, name):
        return _lift(getattr(_np.linalg, name))


linalg = _Linalg()


def __getattr__(name):
    f = getattr(_np, name)
    return _lift(f) if callable(f) and not isinstance(f, type) else f
