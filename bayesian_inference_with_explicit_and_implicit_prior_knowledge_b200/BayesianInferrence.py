"""Matrix-normal-inverse-Wishart conjugate-prior algebra (reference src/BayesianInferrence.py).

Hot-path pieces run on the GPU:
  * batched sufficient statistics of a trajectory  -> `trajectory_statistics` (pgas_suffstats_f64)
  * natural -> standard conversion + posterior draw -> `mniw_posterior_draw` (pgas_mniw_draw_f64)
The small set-up / post-processing conversions the drivers call once per run
(`prior_mniw_2naturalPara` when a prior is built at import time, `prior_mniw_2naturalPara_inv`
/ `prior_mniw_Predictive` on the final averaged statistics) are host float64 LAPACK calls; they
are outside the per-iteration path.
"""
import ctypes as C

import numpy as np
import scipy.linalg as _sla

from . import _lib


def _solve_spd(A, B):
    """A^-1 B through a Cholesky factorisation (src/BayesianInferrence.py:11-13)."""
    return _sla.cho_solve(_sla.cho_factor(np.asarray(A, dtype=np.float64), lower=True), B)


def prior_mniw_2naturalPara(mean, col_cov, row_scale, df):
    """(M, V, Psi, nu) -> (eta0 = V^-1 M^T, eta1 = V^-1, eta2 = M eta0 + Psi, eta3 = nu)
    (src/BayesianInferrence.py:18-32)."""
    mean = np.atleast_2d(np.asarray(mean, dtype=np.float64))
    row_scale = np.atleast_2d(np.asarray(row_scale, dtype=np.float64))
    col_cov = np.asarray(col_cov, dtype=np.float64)
    m = col_cov.shape[0]
    sol = _solve_spd(col_cov, np.concatenate([mean.T, np.eye(m)], axis=1))
    eta_0, eta_1 = sol[:, : mean.shape[0]], sol[:, mean.shape[0]:]
    return eta_0, eta_1, mean @ eta_0 + row_scale, df


def prior_mniw_2naturalPara_inv(eta_0, eta_1, eta_2, eta_3):
    """natural -> (mean (n,M), col_cov (M,M), row_scale (n,n), df) (src/BayesianInferrence.py:35-45)."""
    eta_0 = np.asarray(eta_0, dtype=np.float64)
    eta_1 = np.asarray(eta_1, dtype=np.float64)
    n = eta_0.shape[1]
    sol = _solve_spd(eta_1, np.concatenate([eta_0, np.eye(eta_1.shape[0])], axis=1))
    mean = sol[:, :n].T
    return np.atleast_2d(mean), sol[:, n:], np.atleast_2d(np.asarray(eta_2) - mean @ eta_0), eta_3


def prior_mniw_mean(eta_0, eta_1):
    """(eta1^-1 eta0)^T with eta1 symmetrised (src/BayesianInferrence.py:48-50)."""
    eta_1 = np.asarray(eta_1, dtype=np.float64)
    return _solve_spd(0.5 * (eta_1 + eta_1.T), np.asarray(eta_0, dtype=np.float64)).T


def prior_mniw_calcStatistics(y, basis):
    """Statistics of ONE (basis, y) pair (src/BayesianInferrence.py:53-61); for a whole trajectory use
    `trajectory_statistics`, which runs on the GPU."""
    y = np.atleast_1d(np.asarray(y, dtype=np.float64))
    b = np.atleast_1d(np.asarray(basis, dtype=np.float64))
    return np.outer(b, y), np.outer(b, b), np.outer(y, y), 1


def prior_mniw_Predictive(mean, col_cov, row_scale, df, basis):
    """Student-t predictive parameters at `basis` (src/BayesianInferrence.py:64-89)."""
    basis = np.atleast_2d(np.asarray(basis, dtype=np.float64))
    col_cov = np.atleast_2d(col_cov)
    row_scale = np.atleast_2d(row_scale)
    df_p = df + 1 - row_scale.shape[0]
    return (np.squeeze(basis @ np.asarray(mean).T), basis @ col_cov @ basis.T + np.eye(basis.shape[0]),
            row_scale / df_p, df_p)


def prior_mniw_log_base_measure(T_0, T_1, T_2, T_3):
    """log base measure of the MNIW family (src/BayesianInferrence.py:111-124)."""
    from scipy.special import multigammaln
    T_0, T_1, T_2 = (np.asarray(a, dtype=np.float64) for a in (T_0, T_1, T_2))
    n, m = T_2.shape[0], T_1.shape[0]
    Psi = T_2 - T_0.T @ _solve_spd(T_1, T_0)
    return (-0.5 * n * m * np.log(2 * np.pi) + 0.5 * n * np.linalg.slogdet(T_1)[1] - 0.5 * T_3 * n * np.log(2)
            - multigammaln(T_3 / 2, n) + 0.5 * T_3 * np.linalg.slogdet(Psi)[1])


# ------------------------------------------------------------------------------- GPU path
def trajectory_statistics(model, traj):
    """sum_t prior_mniw_calcStatistics(x_{t+1}, basis(x_t, u_t)) for `traj` (n_chains, T, n_x) CUDA tensor
    (src/PGAS.py:294-303) -> (T0 (n_chains,M,n_x), T1 (n_chains,M,M), T2 (n_chains,n_x,n_x), T3 = T-1)."""
    torch = _lib.require_cuda()
    traj = traj.reshape(-1, model.T, model.n_x).contiguous()
    nc = traj.shape[0]
    T0 = torch.empty((nc, model.M, model.n_x), dtype=torch.float64, device="cuda")
    T1 = torch.empty((nc, model.M, model.M), dtype=torch.float64, device="cuda")
    T2 = torch.empty((nc, model.n_x, model.n_x), dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().pgas_suffstats_f64(model.handle, _lib.ptr(traj), nc, _lib.ptr(T0), _lib.ptr(T1), _lib.ptr(T2),
                                             _lib.stream_ptr()))
    return T0, T1, T2, float(model.T - 1)


def mniw_posterior_draw(eta0, eta1, eta2, eta3, rng, flags=0):
    """(A, Sigma) ~ MNIW(eta) for n_chains stacked natural parameters (src/PGAS.py:306-343).
    Returns (A (n_chains,n_x,M), S (n_chains,n_x,n_x), status (n_chains) int32) CUDA tensors."""
    torch = _lib.require_cuda()
    eta0, eta1, eta2 = (t.contiguous() for t in (eta0, eta1, eta2))
    nc, M, nx = eta0.shape
    A = torch.empty((nc, nx, M), dtype=torch.float64, device="cuda")
    S = torch.empty((nc, nx, nx), dtype=torch.float64, device="cuda")
    status = torch.zeros((nc,), dtype=torch.int32, device="cuda")
    nbytes = _lib.lib().pgas_mniw_draw_workspace_bytes(M, nx, nc)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device="cuda")
    _lib.check(_lib.lib().pgas_mniw_draw_f64(_lib.ptr(eta0), _lib.ptr(eta1), _lib.ptr(eta2), float(eta3), M, nx, nc,
                                             C.byref(rng), int(flags), _lib.ptr(A), _lib.ptr(S), _lib.ptr(status),
                                             _lib.ptr(ws), nbytes, _lib.stream_ptr()))
    return A, S, status


def posterior_predictive_batch(eta0, eta1, eta2, eta3, basis=None):
    """Batched post-processing of traced statistics on the GPU: `jax.vmap(prior_mniw_2naturalPara_inv)`
    over K statistics sets followed by `prior_mniw_Predictive` on a grid, as the reference's figure scripts do
    (SingleMassOscillator_Figures.py:58-89, :131-140).  eta0 (K,M,n), eta1 (K,M,M), eta2 (K,n,n), eta3 (K,): natural parameters
    INCLUDING the prior (array-likes or CUDA tensors); basis (G,M) or None.
    Returns a dict of CUDA tensors: mean (K,n,M), row_scale (K,n,n), df (K,), status (K,) and, with a grid, pred_mean (K,G,n),
    pred_col_scale (K,G) = diag(col_scale) of prior_mniw_Predictive, pred_row_scale (K,n,n), pred_df (K,)."""
    torch = _lib.require_cuda()
    f64 = dict(dtype=torch.float64, device="cuda")
    eta0, eta1, eta2 = (torch.as_tensor(np.asarray(t) if not torch.is_tensor(t) else t, **f64).contiguous() for t in (eta0, eta1, eta2))
    K, M, n = eta0.shape
    eta3 = torch.as_tensor(np.broadcast_to(np.asarray(eta3.cpu() if torch.is_tensor(eta3) else eta3, dtype=np.float64), (K,)).copy(), **f64)
    G = 0
    if basis is not None:
        basis = torch.as_tensor(np.asarray(basis) if not torch.is_tensor(basis) else basis, **f64).contiguous()
        G = basis.shape[0]
    out = dict(mean=torch.empty((K, n, M), **f64), row_scale=torch.empty((K, n, n), **f64), df=torch.empty((K,), **f64),
               status=torch.zeros((K,), dtype=torch.int32, device="cuda"))
    pm = torch.empty((K, max(G, 1), n), **f64)
    pc = torch.empty((K, max(G, 1)), **f64)
    L = _lib.lib()
    nbytes = L.pgas_mniw_posterior_batch_workspace_bytes(M, n, K)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device="cuda")
    _lib.check(L.pgas_mniw_posterior_batch_f64(M, n, K, _lib.ptr(eta0), _lib.ptr(eta1), _lib.ptr(eta2), _lib.ptr(eta3),
                                               _lib.ptr(basis) if G else None, G, _lib.ptr(out["mean"]), _lib.ptr(out["row_scale"]),
                                               _lib.ptr(out["df"]), _lib.ptr(pm), _lib.ptr(pc), _lib.ptr(out["status"]), _lib.ptr(ws),
                                               nbytes, _lib.stream_ptr()))
    if G:
        pdf = out["df"] + 1 - n
        out.update(pred_mean=pm, pred_col_scale=pc, pred_row_scale=out["row_scale"] / pdf[:, None, None], pred_df=pdf)
    return out
