"""Resampling and path reconstruction (reference src/Filtering.py) on the GPU."""
import ctypes as C

import numpy as np

from . import _lib
from . import random as _random


def _uniform_of(key_or_u):
    if isinstance(key_or_u, _random.PhiloxKey):
        return _random.uniform(key_or_u)
    return float(key_or_u)


def systematic_SISR(key, w):
    """Reference signature `systematic_SISR(key, w)` (src/Filtering.py:6).  `key` is a PhiloxKey or, for
    injected-variates testing, the uniform itself.  w: array (N,) or (n_sets, N) -> int32 indices."""
    torch = _lib.require_cuda()
    wt = torch.as_tensor(w, dtype=torch.float64)
    batched = wt.ndim == 2
    w2 = wt.reshape(-1, wt.shape[-1]).contiguous().cuda()
    n_sets, N = w2.shape
    if batched and not isinstance(key, _random.PhiloxKey) and np.ndim(key) == 1:
        u = torch.as_tensor(np.asarray(key, dtype=np.float64)).cuda()
    else:
        u = torch.full((n_sets,), _uniform_of(key), dtype=torch.float64, device="cuda")
    idx = torch.empty((n_sets, N), dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().pgas_resample_f64(_lib.ptr(w2), N, n_sets, _lib.ptr(u), _lib.ptr(idx), _lib.stream_ptr()))
    if torch.is_tensor(w) and w.is_cuda:
        return idx if batched else idx[0]
    out = idx.cpu().numpy()
    return out if batched else out[0]


def reconstruct_trajectory(Particles, ancestry, idx):
    """Reference signature (src/Filtering.py:40).  Particles (T,N[,n]), ancestry (>=T-1,N), idx -> (T[,n])."""
    torch = _lib.require_cuda()
    on_dev = torch.is_tensor(Particles) and Particles.is_cuda
    P = torch.as_tensor(Particles, dtype=torch.float64)
    if P.ndim == 2:
        P = P.unsqueeze(-1)
    T, N, n = P.shape
    P = P.contiguous().cuda()
    anc = torch.as_tensor(np.asarray(ancestry) if not torch.is_tensor(ancestry) else ancestry)
    anc = anc[: T - 1].to(torch.int32).contiguous().cuda()
    if T == 1:
        anc = torch.zeros((1, N), dtype=torch.int32, device="cuda")
    ix = torch.tensor([int(idx)], dtype=torch.int32, device="cuda")
    out = torch.empty((T, n), dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().pgas_reconstruct_trajectory_f64(_lib.ptr(P), _lib.ptr(anc), _lib.ptr(ix), 1, T, N, n,
                                                          _lib.ptr(out), _lib.stream_ptr()))
    res = out.squeeze()
    return res if on_dev else res.cpu().numpy()
