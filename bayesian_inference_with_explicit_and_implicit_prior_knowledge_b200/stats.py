"""Stand-in for `jax.scipy.stats` as the reference's `likelihood_fcn` lambdas use it
(src/EMPS.py:250-252, src/Toy_Example.py:142-144):

    likelihood_fcn=lambda obs, state, input: np.squeeze(stats.multivariate_normal.logpdf(obs, mean=f_y(state), cov=R))

On concrete numbers `logpdf` evaluates the Gaussian log-density (NumPy, set-up time only).  With a mean that is not affine in the
state (or any other expression of observation, state and input) the callable is traced a second time into an expression program
for the log-density itself (models.ProgramLikelihood).  When the sampler traces the
callable (models.resolve_likelihood: symbolic state, a marker in place of the observation) it returns a descriptor of the
Gaussian observation model — output map and covariance — from which the device model is built; the kernels never call Python.
"""
import numpy as np

from . import models as _models


class multivariate_normal:
    @staticmethod
    def logpdf(x, mean, cov):
        if isinstance(x, _models.Sym) or (isinstance(mean, _models.Sym) and not isinstance(x, _models.ObservationMarker)):
            # expression tracing (model plug-in): the Gaussian log-density written out over symbolic observation / mean
            cov = np.atleast_2d(np.asarray(cov, dtype=np.float64))
            n = cov.shape[0]
            L = np.linalg.cholesky(cov)
            W = np.linalg.inv(L)
            d = x - mean
            if len(d) != n:
                raise TypeError(f"multivariate_normal.logpdf: {len(d)} residuals for a {n} x {n} covariance")
            q = None
            for r in range(n):
                e = None
                for c in range(r + 1):
                    term = d[c] * float(W[r, c])
                    e = term if e is None else e + term
                q = e * e if q is None else q + e * e
            return q * (-0.5) + float(-0.5 * n * np.log(2 * np.pi) - np.sum(np.log(np.diag(L))))
        if isinstance(x, _models.ObservationMarker) or isinstance(mean, (_models.Affine, _models.Sym)):
            if not isinstance(x, _models.ObservationMarker):
                raise TypeError("likelihood_fcn: the density must be evaluated at the observation itself")
            if isinstance(mean, _models.Sym):
                raise TypeError("likelihood_fcn: the output map must be affine in the state (Gaussian observation of an affine map)")
            if not isinstance(mean, _models.Affine):          # a constant mean: no dependence on the state
                raise TypeError("likelihood_fcn: the mean does not depend on the state")
            return _models.GaussianLogpdfTrace(mean, cov)
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        mean = np.atleast_1d(np.asarray(mean, dtype=np.float64))
        cov = np.atleast_2d(np.asarray(cov, dtype=np.float64))
        L = np.linalg.cholesky(cov)
        e = np.linalg.solve(L, (x - mean).reshape(-1))
        return -0.5 * float(e @ e) - 0.5 * cov.shape[0] * np.log(2 * np.pi) - float(np.sum(np.log(np.diag(L))))
