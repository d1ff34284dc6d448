"""jax.random stand-in.  Keys are uint64 scalars; `split` derives children with a splitmix64 hash; every draw
comes from a NumPy Philox generator keyed by the key AND is appended to TAPE, so that the run can be replayed
with injected variates.  (JAX's threefry stream is not reproduced.)"""
import numpy as _np
from .numpy import _wrap

TAPE = []
_MASK = (1 << 64) - 1


def tape_reset():
    TAPE.clear()


def _mix(z):
    z = (z + 0x9E3779B97F4A7C15) & _MASK
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
    return z ^ (z >> 31)


def key(seed):
    return _wrap(_np.uint64(_mix(int(seed) & _MASK)))


PRNGKey = key


def split(k, num=2):
    base = int(_np.asarray(k))
    return _wrap(_np.array([_mix(base ^ _mix(i + 1)) for i in range(int(num))], dtype=_np.uint64))


def _rng(k):
    return _np.random.Generator(_np.random.Philox(key=int(_np.asarray(k))))


def _shape(shape):
    if shape is None:
        return ()
    if isinstance(shape, (int, _np.integer)):
        return (int(shape),)
    return tuple(int(s) for s in shape)


def _log(kind, v):
    TAPE.append((kind, _np.array(v, dtype=_np.float64, copy=True)))
    return _wrap(_np.asarray(v, dtype=_np.float64))


def uniform(k, shape=(), dtype=None, minval=0.0, maxval=1.0):
    assert minval == 0.0 and maxval == 1.0
    return _log("uniform", _rng(k).random(_shape(shape)))


def normal(k, shape=(), dtype=None):
    return _log("normal", _rng(k).standard_normal(_shape(shape)))


def multivariate_normal(k, mean, cov, shape=None, dtype=None, method="cholesky"):
    """mean + chol(cov) z with z = normal(key, shape + (n,)) (JAX's default 'cholesky' method)"""
    mean, cov = _np.asarray(mean, dtype=_np.float64), _np.asarray(cov, dtype=_np.float64)
    z = _np.asarray(normal(k, _shape(shape) + mean.shape[-1:]))
    return _wrap(mean + _np.einsum("ij,...j->...i", _np.linalg.cholesky(cov), z))


def chisquare(k, df, shape=None, dtype=None):
    df = _np.asarray(df, dtype=_np.float64)
    return _log("chisquare", _rng(k).chisquare(df, size=_shape(shape) if shape is not None else df.shape))


def t(k, df, shape=(), dtype=None):
    """Student-t variate (JAX: normal * sqrt(df/2 / gamma(df/2)); the tape records the t variate itself)"""
    return _log("t", _rng(k).standard_t(float(_np.asarray(df)), size=_shape(shape)))
