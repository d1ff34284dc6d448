"""Oracle (test infrastructure): restatement of reference src/Filtering.py.

Injected-variates form: the scalar uniform the reference draws from its key
(`src/Filtering.py:19`) is an argument.
"""
import numpy as np


def softmax(x):
    """jax.nn.softmax semantics: exp(x - max) / sum(exp(x - max))
    (used at src/PGAS.py:102,118,224)."""
    x = np.asarray(x, dtype=np.float64)
    e = np.exp(x - np.max(x))
    return e / np.sum(e)


def systematic_SISR(u, w):
    """src/Filtering.py:6-37 with `u` = jax.random.uniform(key) (line 19)."""
    w = np.asarray(w, dtype=np.float64)
    N = len(w)
    w = np.clip(w, 0.0, np.inf)                      # :23
    w_sum = np.sum(w)                                # :24
    # :25  (NaN sum -> comparison False -> uniform weights)
    w = w / w_sum if w_sum > 0 else np.ones_like(w) / N
    U = (u + np.arange(N)) / N                       # :28
    W = np.cumsum(w)                                 # :29
    W = np.clip(W, 0.0, 1.0)                         # :30-32
    indices = np.searchsorted(W, U, side="left")     # :34
    indices = np.clip(indices, 0, N - 1)             # :35
    return indices.astype(np.int64)


def categorical_searchsorted(weights, u):
    """idx = searchsorted(cumsum(w), u)  (src/PGAS.py:122-124, :225) — NOT
    clipped, can return N when cumsum(w)[-1] < u."""
    return int(np.searchsorted(np.cumsum(weights), u, side="left"))


def reconstruct_trajectory(Particles, ancestry, idx):
    """src/Filtering.py:40-55 (backward walk through the ancestor table)."""
    Particles = np.atleast_3d(Particles)
    n_steps = Particles.shape[0]
    n_dim = Particles.shape[-1]
    traj = np.zeros((n_steps, n_dim))
    ancestor_idx = np.zeros((n_steps,))
    ancestor_idx[-1] = idx
    traj[-1] = Particles[-1, int(idx)]
    for i in range(n_steps - 2, -1, -1):
        ancestor_idx[i] = ancestry[i, int(ancestor_idx[i + 1])]
        traj[i] = Particles[i, int(ancestor_idx[i])]
    return np.squeeze(traj)
