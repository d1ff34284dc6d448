"""Oracle (test infrastructure): NumPy restatement of the library's Philox-4x32-10 stream and
its uniform / normal / chi-square mappings (include/pgas_b200.h: pgas_rng; csrc/common.cuh).
This is the library's OWN random-number contract (the reference uses JAX threefry, which cannot
be reproduced without JAX), restated so that the Philox mode of the kernels can be checked by
feeding the same variates to the injected-variates oracle."""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
PURPOSE_STATE, PURPOSE_STEP_U, PURPOSE_DRAW_G, PURPOSE_DRAW_N, PURPOSE_DRAW_CHI = 0, 1, 2, 3, 4
_TINY = 1.0 / 9007199254740992.0


def philox(c0, c1, c2, c3, seed):
    mask = np.uint64(0xFFFFFFFF)
    c = [np.asarray(x, dtype=np.uint64) & mask for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64(seed >> 32)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c[0], np.uint64(M1) * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k0) & mask, p1 & mask, ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & mask, p0 & mask]
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return c


def u53(hi, lo):
    return (((hi >> np.uint64(5)) << np.uint64(26)) | (lo >> np.uint64(6))).astype(np.float64) * _TINY


def uniform2(seed, purpose, chain, it, t, i):
    o = philox(i, t, it, (purpose << 24) | (chain & 0xFFFFFF), seed)
    return u53(o[0], o[1]), u53(o[2], o[3])


def normal2(seed, purpose, chain, it, t, i):
    ua, ub = uniform2(seed, purpose, chain, it, t, i)
    r = np.sqrt(-2.0 * np.log(ua + _TINY))
    return r * np.cos(2 * np.pi * ub), r * np.sin(2 * np.pi * ub)


def sweep_variates(seed, chain, it, T, N, n_x):
    """Z (T,N,n_x), U (T,2) of sweep `it` of chain `chain`."""
    Z = np.zeros((T, N, n_x))
    tt, ii = np.meshgrid(np.arange(T), np.arange(N), indexing="ij")
    for k in range(0, n_x, 2):
        za, zb = normal2(seed, PURPOSE_STATE, chain, it, tt | ((k >> 1) << 28), ii)
        Z[:, :, k] = za
        if k + 1 < n_x:
            Z[:, :, k + 1] = zb
    ua, ub = uniform2(seed, PURPOSE_STEP_U, chain, it, np.arange(T), 0)
    return Z, np.stack([ua, ub], axis=1)


def chisquare(seed, chain, it, idx, nu):
    """Marsaglia-Tsang with Philox-indexed attempts (csrc/mniw_draw.cu: philox_chisquare)."""
    a, boost = 0.5 * nu, 1.0
    if a < 1.0:
        u, _ = uniform2(seed, PURPOSE_DRAW_CHI, chain, it, 0x40000000, idx)
        boost = float(u + _TINY) ** (1.0 / a)
        a += 1.0
    d = a - 1.0 / 3.0
    c = 1.0 / np.sqrt(9.0 * d)
    attempt = 0
    while True:
        x, _ = normal2(seed, PURPOSE_DRAW_CHI, chain, it, 2 * attempt, idx)
        u, _ = uniform2(seed, PURPOSE_DRAW_CHI, chain, it, 2 * attempt + 1, idx)
        x, u = float(x), float(u)
        v = 1.0 + c * x
        attempt += 1
        if v <= 0.0:
            continue
        v = v ** 3
        if np.log(u + _TINY) < 0.5 * x * x + d - d * v + d * np.log(v) or attempt > 1001:
            return 2.0 * d * v * boost


def draw_variates(seed, chain, it, n_x, M, df):
    """chi2 (n_x,), G (n_x,n_x), Nrm (n_x,M) of draw `it` of chain `chain`; df = eta3."""
    chi2 = np.array([chisquare(seed, chain, it, i, df - i) for i in range(n_x)])
    flat = np.arange(n_x * n_x)
    za, zb = normal2(seed, PURPOSE_DRAW_G, chain, it, 0, flat >> 1)
    G = np.where(flat & 1, zb, za).reshape(n_x, n_x)
    flat = np.arange(n_x * M)
    za, zb = normal2(seed, PURPOSE_DRAW_N, chain, it, 0, flat >> 1)
    Nrm = np.where(flat & 1, zb, za).reshape(n_x, M)
    return chi2, G, Nrm
