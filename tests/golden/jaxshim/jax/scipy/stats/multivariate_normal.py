import numpy as _np
import scipy.linalg as _sla
from ...numpy import _wrap


def logpdf(x, mean, cov, allow_singular=None):
    """jax.scipy.stats.multivariate_normal.logpdf (JAX 0.4.38): through the Cholesky factor of cov"""
    x, mean, cov = (_np.asarray(v, dtype=_np.float64) for v in (x, mean, cov))
    if not mean.shape:
        return _wrap(-0.5 * _np.square(x - mean) / cov - 0.5 * (_np.log(2 * _np.pi) + _np.log(cov)))
    n = mean.shape[-1]
    if not cov.shape:
        y = x - mean
        return _wrap(-0.5 * _np.einsum("...i,...i->...", y, y) / cov - n / 2 * (_np.log(2 * _np.pi) + _np.log(cov)))
    assert cov.ndim == 2 and cov.shape == (n, n)
    L = _np.linalg.cholesky(cov)
    d = _np.atleast_1d(x - mean)
    y = _sla.solve_triangular(L, d.reshape(-1, n).T, lower=True).T.reshape(d.shape)
    return _wrap(_np.asarray(-0.5 * _np.einsum("...i,...i->...", y, y) - n / 2 * _np.log(2 * _np.pi)
                             - _np.log(_np.diagonal(L)).sum(-1)))
