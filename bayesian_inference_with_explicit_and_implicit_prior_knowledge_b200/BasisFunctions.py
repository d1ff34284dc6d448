"""Hilbert-space Gaussian-process basis (reference src/BasisFunctions.py).

`generate_Hilbert_BasisFunction` keeps the reference signature and returns
`(basis, spectral_density)`; `basis` is a callable like the reference's jitted closure, but it
evaluates on the GPU (pgas_hgp_eval_f64) and, when applied to a traced (state, input) value,
yields the symbolic descriptor the sweep kernels consume (models.BasisExpr).
"""
import ctypes as C
import heapq

import numpy as np

from . import _lib
from . import models as _models
from . import tracing as _tracing


def _lattice_search(num_fcn, size, idx_start, idx_step):
    """The num_fcn lattice points with the smallest Laplacian eigenvalue, in the reference's
    order (src/BasisFunctions.py:24-57): Dijkstra-like expansion from the corner where a child's
    cost is its PARENT's accumulated float cost plus the single-coordinate increment, and equal
    costs are ordered by the position tuple."""
    D = size.shape[0]
    lattice = np.arange(idx_start, num_fcn * idx_step + 1 + idx_start, idx_step)
    per_dim = (np.pi / size) ** 2
    sq = lattice ** 2
    root = (0,) * D
    heap = [(float(np.sum(per_dim * sq[0])), root)]
    pushed = {root}
    order = []
    while heap and len(order) < num_fcn:
        cost, node = heapq.heappop(heap)
        order.append(node)
        for d in range(D):
            k = node[d] + 1
            if k >= len(lattice):
                continue
            child = node[:d] + (k,) + node[d + 1:]
            if child not in pushed:
                pushed.add(child)
                heapq.heappush(heap, (cost + float(per_dim[d] * (sq[k] - sq[k - 1])), child))
    return lattice[np.array(order, dtype=np.int64)].reshape(len(order), D)


def _spectral_density_Gaussian(freq, magnitude, lengthscale):
    """Spectral density of the squared-exponential kernel (src/BasisFunctions.py:83-105)."""
    freq = np.atleast_1d(np.asarray(freq, dtype=np.float64))
    ls = np.broadcast_to(np.asarray(lengthscale, dtype=np.float64), freq.shape)
    D = freq.shape[0]
    return magnitude * (2 * np.pi) ** (D / 2) * np.prod(ls) * np.exp(-0.5 * np.sum(ls ** 2 * freq ** 2))


class HilbertBasis:
    """phi_m(x) = prod_d L_d^-1/2 sin(pi S[m,d] (x_d - c_d + L_d) / (2 L_d))  (src/BasisFunctions.py:60-80)."""

    def __init__(self, freq, domain_boundary, idx_start, idx_step):
        bounds = np.atleast_2d(np.asarray(domain_boundary, dtype=np.float64))
        self.freq = np.ascontiguousarray(freq, dtype=np.int32)
        self.M, self.D = self.freq.shape
        self.center = (bounds[:, 0] + bounds[:, 1]) / 2
        self.half_width = (bounds[:, 1] - bounds[:, 0]) / 2
        self.idx_start, self.idx_step = int(idx_start), int(idx_step)
        self.eigen_val = (np.pi * self.freq.astype(np.float64) / (2 * self.half_width)) ** 2
        self._eval_model = None

    def _model(self):
        if self._eval_model is None:
            lik = _models.GaussianLikelihood(np.eye(1, self.D), np.zeros(1), np.eye(1))
            ident = _models.BasisExpr(self, np.eye(self.D), np.zeros(self.D))
            self._eval_model = _models.DeviceModel(np.zeros((2, 1)), np.zeros((2, 0)), np.zeros(self.D),
                                                   np.eye(self.D), lik, ident)
        return self._eval_model

    def __call__(self, x):
        if isinstance(x, _tracing.Expr):                 # marginalised filters (Algorithm1/2/3): per-step tracer
            return _tracing.BasisCall(self, x)
        if isinstance(x, _models.Sym):                   # non-affine GP-input map: expression program (model plug-in)
            return _models.ProgramBasis(self, x)
        if isinstance(x, _models.Affine):
            if len(x) != self.D:
                raise ValueError(f"basis expects {self.D} inputs, traced value has {len(x)}")
            return _models.BasisExpr(self, x.A, x.b)
        torch = _lib.require_cuda()
        xt = torch.as_tensor(np.asarray(x, dtype=np.float64) if not torch.is_tensor(x) else x, dtype=torch.float64)
        single = xt.ndim == 0 or (xt.ndim == 1 and self.D > 1) or (xt.ndim == 1 and self.D == 1 and xt.numel() == 1)
        pts = xt.reshape(-1, self.D).contiguous().cuda()
        out = torch.empty((pts.shape[0], self.M), dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().pgas_hgp_eval_f64(self._model().handle, _lib.ptr(pts), C.c_void_p(0), 0, pts.shape[0],
                                                _lib.ptr(out), _lib.stream_ptr()))
        if torch.is_tensor(x) and x.is_cuda:
            return out[0] if single else out
        res = out.cpu().numpy()
        return res[0] if single else res


def generate_Hilbert_BasisFunction(num_fcn, domain_boundary, lengthscale, scale, idx_start=1, idx_step=1):
    """Reference signature (src/BasisFunctions.py:8-10) -> (basis callable, spectral_density (M,))."""
    bounds = np.atleast_2d(np.asarray(domain_boundary, dtype=np.float64))
    if idx_start < 1:
        idx_start = 1
    size = bounds[:, 1] - bounds[:, 0]
    freq = _lattice_search(num_fcn, size, idx_start, idx_step)
    basis = HilbertBasis(freq, bounds, idx_start, idx_step)
    sd = np.array([_spectral_density_Gaussian(np.sqrt(ev), scale, lengthscale) for ev in basis.eigen_val])
    return basis, sd
