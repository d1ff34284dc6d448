"""Oracle (test infrastructure): restatement of reference src/BasisFunctions.py.

Hilbert-space GP basis on a box domain: eigenfunctions of the Dirichlet
Laplacian, phi_m(x) = prod_d L_d^{-1/2} sin(sqrt(lambda_{m,d}) (x_d - c_d + L_d)),
lambda_{m,d} = (pi s_{m,d} / size_d)^2  (src/BasisFunctions.py:60,63-66,77-80).
"""
import heapq

import numpy as np


def select_indices(num_fcn, domain_boundary, idx_start=1, idx_step=1):
    """The index-lattice search of src/BasisFunctions.py:12-57.

    Returns the (num_fcn, D) integer array of per-dimension frequencies, in the
    reference's selection order: a best-first walk from the lattice corner with
    the cost sum_d (pi/size_d)^2 j_d^2 accumulated *incrementally in floating
    point* along the walk (:50-52) and ties broken by the position tuple
    (heapq on (cost, tuple), :31,53).
    """
    bounds = np.atleast_2d(np.asarray(domain_boundary, dtype=np.float64))
    D = bounds.shape[0]
    if idx_start < 1:                                         # :19-20
        idx_start = 1
    size = bounds[:, 1] - bounds[:, 0]                        # :23
    stop = num_fcn * idx_step + 1 + idx_start                 # :24
    freq = np.arange(idx_start, stop, idx_step)               # :25
    wgt = (np.pi / size) ** 2                                 # :29
    fsq = freq ** 2                                           # :30

    origin = (0,) * D
    frontier = [(float(np.sum(wgt * fsq[0])), origin)]        # :33-35
    seen = {origin}
    chosen = []
    while len(chosen) < num_fcn and frontier:                 # :39
        cost, pos = heapq.heappop(frontier)
        chosen.append(freq[np.array(pos, dtype=int)])         # :41-42
        for d in range(D):                                    # :45
            if pos[d] + 1 >= len(freq):
                continue
            nxt = pos[:d] + (pos[d] + 1,) + pos[d + 1:]
            if nxt in seen:
                continue
            step_cost = cost + float(wgt[d] * (fsq[nxt[d]] - fsq[pos[d]]))  # :51-53
            heapq.heappush(frontier, (step_cost, nxt))
            seen.add(nxt)
    return np.array(chosen, dtype=np.int64)


def eigen_fnc(x, eigen_val, L):
    """src/BasisFunctions.py:77-80.  x (D,), eigen_val (M,D), L (D,) -> (M,)."""
    return np.prod(np.sqrt(1.0 / L) * np.sin(np.sqrt(eigen_val) * (x + L)), axis=1)


def spectral_density_gaussian(freq, magnitude, lengthscale):
    """src/BasisFunctions.py:83-105 for one frequency row (D,)."""
    freq = np.asarray(freq, dtype=np.float64)
    D = len(freq)
    ls = np.broadcast_to(np.asarray(lengthscale, dtype=np.float64), freq.shape)
    return (magnitude * (2.0 * np.pi) ** (D / 2.0) * np.prod(ls)
            * np.exp(-0.5 * np.sum(ls ** 2 * freq ** 2)))


class HilbertBasis:
    """Callable returned by generate_Hilbert_BasisFunction (the reference returns a
    jitted closure; this keeps the pieces inspectable for the tests)."""

    def __init__(self, indices, domain_boundary):
        bounds = np.atleast_2d(np.asarray(domain_boundary, dtype=np.float64))
        self.indices = indices
        self.center = (bounds[:, 0] + bounds[:, 1]) / 2.0          # :15
        self.size = bounds[:, 1] - bounds[:, 0]
        self.L = self.size / 2.0
        self.eig_val = (np.pi * indices.astype(np.float64) / self.size) ** 2  # :60

    def __call__(self, x):
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        return eigen_fnc(x - self.center, self.eig_val, self.L)    # :63-66

    def batch(self, X):
        """vmap(basis)(X): X (n,D) or (n,) for D=1 -> (n,M)."""
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[:, None]
        arg = np.sqrt(self.eig_val)[None] * ((X - self.center) + self.L)[:, None, :]
        return np.prod(np.sqrt(1.0 / self.L) * np.sin(arg), axis=2)


def generate_Hilbert_BasisFunction(num_fcn, domain_boundary, lengthscale, scale,
                                   idx_start=1, idx_step=1):
    """src/BasisFunctions.py:8-74 -> (basis callable, spectral_density (M,))."""
    idx = select_indices(num_fcn, domain_boundary, idx_start, idx_step)
    basis = HilbertBasis(idx, domain_boundary)
    sd = np.array([spectral_density_gaussian(np.sqrt(ev), scale, lengthscale)
                   for ev in basis.eig_val])                       # :69-72
    return basis, sd
