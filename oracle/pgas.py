"""Oracle (test infrastructure): restatement of reference src/PGAS.py — the
Theta-conditioned conditional SMC sweep with ancestor sampling and the MNIW
parameter draw — in float64 NumPy with INJECTED VARIATES in place of JAX keys.

Variates contract (identical to the CUDA library's
"injected" mode, include/pgas_b200.h):
  sweep:  Z (T,N,n_x) standard normals, Z[0] initial particles, Z[t] process
          noise of step t;  U (T,2) uniforms, U[t,0]=u_res(t), U[t,1]=u_anc(t)
          for t>=1, U[0,0]=u_idx (final trajectory pick), U[0,1] unused.
  draw:   chi2 (n_x,) chi-square(nu - i) variates, G (n_x,n_x) normals (strict
          lower triangle used), Nrm (n_x,M) normals.

The three reference quirks are reproduced by default and can
be switched off with flags, mirroring the library's flags.
"""
import numpy as np
import scipy.linalg as sla

from . import filtering as F
from . import mniw


# --------------------------------------------------------------------------- model
def mvn_logpdf(x, mean, cov):
    """jax.scipy.stats.multivariate_normal.logpdf semantics (Cholesky whitening),
    vectorised over the leading axis of `mean` (src/PGAS.py:109-116,
    src/StateSpaceModel.py:83-87)."""
    mean = np.asarray(mean, dtype=np.float64)          # (batch, n)
    x = np.atleast_1d(np.asarray(x, dtype=np.float64)).ravel()
    cov = np.atleast_2d(np.asarray(cov, dtype=np.float64))
    n = cov.shape[0]
    L = np.linalg.cholesky(cov)
    y = sla.solve_triangular(L, (x[None, :] - mean).T, lower=True)   # (n, batch)
    return (-0.5 * np.sum(y * y, axis=0) - 0.5 * n * np.log(2.0 * np.pi)
            - np.sum(np.log(np.diag(L))))


class ThetaModel:
    """Data + the two user callables of condSequentialMonteCarlo
    (src/PGAS.py:24-43), batched over particles (the reference vmaps them).

    basis(states (n,n_x), input) -> (n,M);  loglik(obs, states (n,n_x), input) -> (n,)
    """

    def __init__(self, observations, inputs, m0, P0, basis, loglik):
        self.observations = np.asarray(observations, dtype=np.float64)
        self.inputs = np.asarray(inputs, dtype=np.float64)
        self.m0 = np.atleast_1d(np.asarray(m0, dtype=np.float64))
        self.P0 = np.atleast_2d(np.asarray(P0, dtype=np.float64))
        self.basis = basis
        self.loglik = loglik
        self.T = self.observations.shape[0]
        self.n_x = self.m0.shape[0]


def gaussian_loglik(H, h0, R):
    """likelihood_fcn(obs, state, input) = N(obs; H state + h0, R) as the shipped
    callables are (src/EMPS.py:250-252, src/Toy_Example.py:142-144)."""
    H = np.atleast_2d(np.asarray(H, dtype=np.float64))
    h0 = np.atleast_1d(np.asarray(h0, dtype=np.float64))

    def loglik(obs, states, inp):
        return mvn_logpdf(obs, states @ H.T + h0, R)
    return loglik


def affine_hgp_basis(hgp, A, b):
    """basis_fcn(state, input) = hgp(A [state; input] + b)  (src/EMPS.py:110-113,
    src/Toy_Example.py:146, src/SingleMassOscillator.py:151)."""
    A = np.atleast_2d(np.asarray(A, dtype=np.float64))
    b = np.atleast_1d(np.asarray(b, dtype=np.float64))

    def basis(states, inp):
        states = np.atleast_2d(states)
        inp = np.atleast_1d(np.asarray(inp, dtype=np.float64)).ravel()
        full = np.concatenate([states, np.broadcast_to(inp, (states.shape[0], inp.size))], axis=1)
        return hgp.batch(full @ A.T + b)
    return basis


def vehicle_slip_basis(hgp, l_f, l_r):
    """2-D tensor-product basis over the slip angles (alpha_f, alpha_r) of
    src/Vehicle.py:50-57 (BASELINE config 5's GP input map)."""
    def basis(states, inp):
        states = np.atleast_2d(states)
        vy_f = states[:, 1] + states[:, 0] * l_f
        vy_r = states[:, 1] - states[:, 0] * l_r
        a_f = inp[0] - np.arctan(vy_f / inp[1])
        a_r = -np.arctan(vy_r / inp[1])
        return hgp.batch(np.stack([a_f, a_r], axis=1))
    return basis


# --------------------------------------------------------------------------- sweep
def csmc_step(model, t, logw, state, Theta, Sigma, ref_t, u_res, u_anc, z,
              ancestor_gather=False, input_offset=0):
    """condSequentialMonteCarlo.step, src/PGAS.py:79-153.

    Returns (new_logw (N,), new_state (N,n_x), a_indices (N,), extras)."""
    u_t = model.inputs[t + input_offset]
    y_t = model.observations[t]
    # :89 -> :45-57   Phi(x_{t-1}, u_t) Theta^T
    aux = model.basis(state, u_t) @ Theta.T
    # :92-102
    ll_aux = model.loglik(y_t, aux, u_t)
    lw_aux = ll_aux + logw
    w_aux = F.softmax(lw_aux)
    # :105-106
    a = F.systematic_SISR(u_res, w_aux)
    # :109-124
    h = mvn_logpdf(ref_t, aux, Sigma)
    w_anc = F.softmax(lw_aux + h)
    ref_idx = F.categorical_searchsorted(w_anc, u_anc)
    a[-1] = ref_idx                                                # :127
    # :130-134 -> :59-77  (quirk i: propagates particle i from particle i)
    src = state[np.clip(a, 0, len(a) - 1)] if ancestor_gather else state
    mean = model.basis(src, u_t) @ Theta.T
    new_state = mean + z @ np.linalg.cholesky(np.atleast_2d(Sigma)).T
    new_state[-1] = ref_t
    # :137-147  (JAX gather clamps an out-of-range ref_idx)
    new_logw = model.loglik(y_t, new_state, u_t) - ll_aux[np.clip(a, 0, len(a) - 1)]
    return new_logw, new_state, a, dict(aux=aux, ll_aux=ll_aux, w_aux=w_aux, w_anc=w_anc)


def csmc_sweep(model, N, ref, Theta, Sigma, Z, U, ancestor_gather=False,
               input_offset=0, keep_weights=False):
    """condSequentialMonteCarlo.__call__, src/PGAS.py:176-228 (+ :155-174)."""
    T, n_x = model.T, model.n_x
    ref = np.asarray(ref, dtype=np.float64).reshape(T, n_x)
    Theta = np.atleast_2d(np.asarray(Theta, dtype=np.float64))
    Sigma = np.atleast_2d(np.asarray(Sigma, dtype=np.float64))
    state_trace = np.zeros((T, N, n_x))
    logw_trace = np.zeros((T, N))
    anc_trace = np.zeros((T, N), dtype=np.int64)
    # :167-172   x_0 ~ N(m0, P0)  == m0 + chol(P0) z
    state_trace[0] = model.m0 + Z[0] @ np.linalg.cholesky(model.P0).T
    state_trace[0, -1] = ref[0]                                    # :194
    cdfs = [] if keep_weights else None
    for t in range(1, T):                                          # :199
        lw, xs, a, ex = csmc_step(model, t, logw_trace[t - 1], state_trace[t - 1],
                                  Theta, Sigma, ref[t], U[t, 0], U[t, 1], Z[t],
                                  ancestor_gather, input_offset)
        state_trace[t] = xs
        logw_trace[t] = lw
        anc_trace[t - 1] = a                                       # :219-221
        if keep_weights:
            cdfs.append((ex["w_aux"], ex["w_anc"]))
    w = F.softmax(logw_trace[-1])                                  # :224
    idx = F.categorical_searchsorted(w, U[0, 0])                   # :225
    idx_c = min(idx, N - 1)
    traj = F.reconstruct_trajectory(state_trace, np.clip(anc_trace, 0, N - 1), idx_c)
    return dict(traj=traj.reshape(T, n_x), state_trace=state_trace, logw_trace=logw_trace,
                anc_trace=anc_trace[: T - 1], idx=idx, w_final=w, cdfs=cdfs)


# --------------------------------------------------------------------------- params
def suff_stats(model, traj, input_offset_stats=0):
    """First half of PGAS.sample_params, src/PGAS.py:294-303 (statistics only,
    prior not yet added): T0 = Phi^T Y, T1 = Phi^T Phi, T2 = Y^T Y, T3 = T-1."""
    traj = np.asarray(traj, dtype=np.float64).reshape(model.T, model.n_x)
    X, Y = traj[:-1], traj[1:]
    U = model.inputs[:-1] if input_offset_stats == 0 else model.inputs[1:]
    Phi = np.stack([model.basis(X[t:t + 1], U[t])[0] for t in range(model.T - 1)])
    return Phi.T @ Y, Phi.T @ Phi, Y.T @ Y, float(model.T - 1)


def mniw_draw(eta, chi2, G, Nrm, transpose_fix=False):
    """Second half of PGAS.sample_params, src/PGAS.py:306-343.
    eta = (eta0 (M,n), eta1 (M,M), eta2 (n,n), eta3)."""
    mean, col_cov, row_scale, df = mniw.prior_mniw_2naturalPara_inv(*eta)   # :306-308
    p = row_scale.shape[0]
    chol_row = np.linalg.cholesky(row_scale)                               # :317
    L = sla.solve_triangular(chol_row, np.eye(p), lower=True)              # :319
    Tm = np.tril(np.asarray(G, dtype=np.float64), k=-1) + np.diag(np.sqrt(chi2))  # :327-329
    C = L @ Tm                                                             # :332
    S_chol = sla.solve_triangular(C.T, np.eye(p), lower=False)             # :334
    S = S_chol @ S_chol.T                                                  # :335
    V_chol = np.linalg.cholesky(col_cov)                                   # :339
    Vf = V_chol.T if transpose_fix else V_chol
    A = mean + S_chol @ np.asarray(Nrm, dtype=np.float64) @ Vf             # :341 (quirk iii)
    return A, S, dict(mean=mean, col_cov=col_cov, row_scale=row_scale, df=df, V_chol=V_chol)


def sample_params(model, prior, traj, chi2, G, Nrm, transpose_fix=False):
    """PGAS.sample_params, src/PGAS.py:288-343."""
    T0, T1, T2, T3 = suff_stats(model, traj)
    eta = (prior[0] + T0, prior[1] + T1, prior[2] + T2, prior[3] + T3)
    A, S, ex = mniw_draw(eta, chi2, G, Nrm, transpose_fix)
    ex["stats"] = (T0, T1, T2, T3)
    return A, S, ex


def pgas_run(model, N, K, prior, init_ref, variates, **flags):
    """PGAS.__call__, src/PGAS.py:345-397.  `variates(k)` returns a dict with keys
    chi2, G, Nrm (draw that FOLLOWS trajectory k) and, for k>=1, Z, U (sweep k).
    Returns (state_trace (T,K,n_x), log_likelihood (T,K), A_trace, S_trace)."""
    T, n_x = model.T, model.n_x
    state_trace = np.zeros((K, T, n_x))
    state_trace[0] = np.asarray(init_ref, dtype=np.float64).reshape(T, n_x)   # :274
    v = variates(0)
    A, S, _ = sample_params(model, prior, state_trace[0], v["chi2"], v["G"], v["Nrm"])
    A_tr, S_tr = [A], [S]
    for k in range(1, K):                                                      # :361
        v = variates(k)
        sw = csmc_sweep(model, N, state_trace[k - 1], A, S, v["Z"], v["U"], **flags)
        state_trace[k] = sw["traj"]
        A, S, _ = sample_params(model, prior, state_trace[k], v["chi2"], v["G"], v["Nrm"])
        A_tr.append(A)
        S_tr.append(S)
    state_trace = np.swapaxes(state_trace, 0, 1)                               # :380
    ll = np.stack([model.loglik(model.observations[t], state_trace[t], model.inputs[t])
                   for t in range(T)])                                         # :383-392
    return state_trace, ll, np.stack(A_tr), np.stack(S_tr)
