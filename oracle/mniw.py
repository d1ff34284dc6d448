"""Oracle (test infrastructure): restatement of reference src/BayesianInferrence.py
(matrix-normal-inverse-Wishart conjugate-prior algebra), float64 NumPy/SciPy.
"""
import numpy as np
import scipy.linalg as sla
from scipy.special import multigammaln


def _solve_spd(A, B):
    """src/BayesianInferrence.py:11-13 — Cholesky + cho_solve."""
    L = np.linalg.cholesky(A)
    return sla.cho_solve((L, True), B)


def prior_mniw_2naturalPara(mean, col_cov, row_scale, df):
    """src/BayesianInferrence.py:18-32."""
    mean = np.atleast_2d(np.asarray(mean, dtype=np.float64))
    row_scale = np.atleast_2d(np.asarray(row_scale, dtype=np.float64))
    col_cov = np.asarray(col_cov, dtype=np.float64)
    temp = _solve_spd(col_cov, np.hstack([mean.T, np.eye(col_cov.shape[0])]))
    eta_0 = temp[:, : mean.shape[0]]
    eta_1 = temp[:, mean.shape[0]:]
    eta_2 = mean @ eta_0 + row_scale
    eta_3 = df
    return eta_0, eta_1, eta_2, eta_3


def prior_mniw_2naturalPara_inv(eta_0, eta_1, eta_2, eta_3):
    """src/BayesianInferrence.py:35-45."""
    eta_0 = np.asarray(eta_0, dtype=np.float64)
    eta_1 = np.asarray(eta_1, dtype=np.float64)
    temp = _solve_spd(eta_1, np.hstack([eta_0, np.eye(eta_1.shape[0])]))
    mean = temp[:, : eta_0.shape[1]].T
    col_cov = temp[:, eta_0.shape[1]:]
    row_scale = eta_2 - mean @ eta_0
    df = eta_3
    return np.atleast_2d(mean), col_cov, np.atleast_2d(row_scale), df


def prior_mniw_mean(eta_0, eta_1):
    """src/BayesianInferrence.py:48-50."""
    eta_1_sym = 0.5 * (eta_1 + eta_1.T)
    return _solve_spd(eta_1_sym, eta_0).T


def prior_mniw_calcStatistics(y, basis):
    """src/BayesianInferrence.py:53-61."""
    y = np.atleast_1d(y)
    basis = np.atleast_1d(basis)
    return np.outer(basis, y), np.outer(basis, basis), np.outer(y, y), 1


def prior_mniw_Predictive(mean, col_cov, row_scale, df, basis):
    """src/BayesianInferrence.py:64-89."""
    basis = np.atleast_2d(basis)
    col_cov = np.atleast_2d(col_cov)
    row_scale = np.atleast_2d(row_scale)
    n_b = basis.shape[0]
    df = df + 1 - row_scale.shape[0]
    mean = np.squeeze(basis @ mean.T)
    col_scale = basis @ col_cov @ basis.T + np.eye(n_b)
    row_scale = row_scale / df
    return mean, col_scale, row_scale, df


def prior_mniw_drawPred(t_samples, mean, col_scale, row_scale, df):
    """src/BayesianInferrence.py:92-108 with `t_samples` = jax.random.t(key, df,
    shape=(n,)) injected (jax.random.t = normal / sqrt(gamma(df/2) / (df/2)))."""
    chol_col = np.linalg.cholesky(np.atleast_2d(col_scale))
    chol_row = np.linalg.cholesky(np.atleast_2d(row_scale))
    t_samples = np.atleast_1d(t_samples)
    return mean + np.squeeze(np.einsum("ij,j,jk->ik", chol_row, t_samples, chol_col.T))


def prior_mniw_log_base_measure(T_0, T_1, T_2, T_3):
    """src/BayesianInferrence.py:111-124."""
    n = T_2.shape[0]
    m = T_1.shape[0]
    Psi = T_2 - T_0.T @ _solve_spd(T_1, T_0)
    nu = T_3
    temp_1 = -0.5 * n * m * np.log(2 * np.pi)
    temp_2 = 0.5 * n * np.log(np.linalg.det(T_1))
    temp_3 = -0.5 * nu * n * np.log(2)
    temp_4 = -multigammaln(nu / 2, n)
    temp_5 = np.log(np.linalg.det(Psi)) * nu / 2
    return temp_1 + temp_2 + temp_3 + temp_4 + temp_5
