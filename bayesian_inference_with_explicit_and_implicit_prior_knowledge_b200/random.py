"""Counter-based keys (Philox-4x32-10) standing in for jax.random keys.

The reference threads `jax.random.key(12345678)` through `jax.random.split`
(src/PGAS.py:184,203,356,365,377).  JAX's threefry stream cannot be reproduced without JAX, so
the library defines its own stream (include/pgas_b200.h: pgas_rng): a key is a 64-bit Philox key;
`split` derives children by hashing (key, index) with Philox itself.  The device kernels index
the stream by (chain, iteration, time, particle), so a whole PGAS run needs ONE key.
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Philox-4x32-10 on uint32 numpy arrays (broadcasting); returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3))
    k0 = np.asarray(k0, dtype=np.uint64) & _MASK
    k1 = np.asarray(k1, dtype=np.uint64) & _MASK
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & _MASK
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & _MASK
        c1, c3, c0, c2 = p1 & _MASK, p0 & _MASK, n0, n2
        k0 = (k0 + np.uint64(_W0)) & _MASK
        k1 = (k1 + np.uint64(_W1)) & _MASK
    return tuple(np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))


def u53(hi, lo):
    k = ((np.asarray(hi, dtype=np.uint64) >> np.uint64(5)) << np.uint64(26)) | (np.asarray(lo, dtype=np.uint64) >> np.uint64(6))
    return k.astype(np.float64) * (1.0 / 9007199254740992.0)


class PhiloxKey:
    """64-bit Philox key; accepted wherever the reference takes a jax.random key."""
    __slots__ = ("seed",)

    def __init__(self, seed):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF

    def __repr__(self):
        return f"PhiloxKey(0x{self.seed:016x})"


def key(seed):
    return PhiloxKey(seed)


PRNGKey = key


def split(k, num=2):
    """jax.random.split stand-in: `num` child keys."""
    idx = np.arange(num, dtype=np.uint64)
    o = philox4x32(idx, 0, 0, 0xFFFFFFFF, k.seed & 0xFFFFFFFF, k.seed >> 32)
    return [PhiloxKey((int(o[1][i]) << 32) | int(o[0][i])) for i in range(num)]


def uniform(k, shape=()):
    """jax.random.uniform stand-in for the drivers' scalar draws (host side)."""
    n = int(np.prod(shape)) if shape != () else 1
    o = philox4x32(np.arange(n, dtype=np.uint64), 0, 0, 0xFFFFFFFE, k.seed & 0xFFFFFFFF, k.seed >> 32)
    u = u53(o[0], o[1])
    return float(u[0]) if shape == () else u.reshape(shape)


def as_key(k):
    if isinstance(k, PhiloxKey):
        return k
    if isinstance(k, (int, np.integer)):
        return PhiloxKey(k)
    raise TypeError(f"expected a PhiloxKey (see {__name__}.key), got {type(k).__name__}")


def normal(k, shape=()):
    """jax.random.normal stand-in for the example modules' data synthesis (host side, Box-Muller)."""
    n = int(np.prod(shape)) if shape != () else 1
    o = philox4x32(np.arange(n, dtype=np.uint64), 0, 0, 0xFFFFFFFD, k.seed & 0xFFFFFFFF, k.seed >> 32)
    u1, u2 = u53(o[0], o[1]) + 1.0 / 9007199254740992.0, u53(o[2], o[3])
    z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    return float(z[0]) if shape == () else z.reshape(shape)
