"""Oracle (test infrastructure): restatement of the reference's marginalised particle filters —
Algorithm1 (online marginalised auxiliary PF, src/Algorithm1.py), Algorithm3 (marginalised
conditional SMC with ancestor sampling, src/Algorithm3.py) and Algorithm2 (the PGAS outer loop,
src/Algorithm2.py) — in float64 NumPy with INJECTED VARIATES in place of JAX keys.

Every particle carries the MNIW sufficient statistics of each GP (T0 (M,n), T1 (M,M), T2 (n,n), T3);
the GP output ("interface variable" xi) is drawn from the Student-t predictive.

Variates contract (identical to the CUDA library's injected mode, csrc/marginal.cu):
  init:  Z0 (N,n_x) normals, ZXI0[g] (N,n_xi) normals
  step t=1..T-1: u_res, [u_anc (Algorithm3)], Z[t] (N,n_x) normals, TS[g][t] (N,n_xi) Student-t(df') variates
                 (jax.random.t(key, df) draws, src/BayesianInferrence.py:103)
  final (Algorithm3): u_idx
"""
import numpy as np
from scipy.special import gammaln

from . import filtering as F
from . import mniw
from .pgas import mvn_logpdf


class SSM:
    """StateSpaceModel (src/StateSpaceModel.py:8-87), batched over particles.
    transition(x (n,n_x), u, *xi (n,n_xi)) -> (n,n_x); output(x, u, *xi) -> (n,n_y)."""

    def __init__(self, Q, R, transition, output):
        self.Q = np.atleast_2d(np.asarray(Q, dtype=np.float64))
        self.R = np.atleast_2d(np.asarray(R, dtype=np.float64))
        self.transition, self.output = transition, output
        self.is_deterministic = bool(np.all(self.Q == 0))                 # :30

    def draw_state(self, z, x, u, *xi):
        new = self.transition(x, u, *xi)                                   # :64
        if self.is_deterministic:
            return new
        return new + z @ np.linalg.cholesky(self.Q).T                      # :67-73

    def log_likelihood(self, y, x, u, *xi):
        out = np.atleast_2d(self.output(x, u, *xi).T).T                    # (n,n_y)
        return mvn_logpdf(y, out, self.R)                                  # :82-87


class MargModel:
    """data + callables of Algorithm1.__init__ (src/Algorithm1.py:27-66)"""

    def __init__(self, observations, inputs, ssm, m0, P0, xi_mean, xi_cov, priors, bases):
        self.obs = np.asarray(observations, dtype=np.float64)
        self.inputs = np.asarray(inputs, dtype=np.float64)
        self.ssm = ssm
        self.m0 = np.atleast_1d(np.asarray(m0, dtype=np.float64))
        self.P0 = np.atleast_2d(np.asarray(P0, dtype=np.float64))
        self.xi_mean = [np.atleast_1d(np.asarray(m, dtype=np.float64)) for m in xi_mean]
        self.xi_cov = [np.atleast_2d(np.asarray(c, dtype=np.float64)) for c in xi_cov]
        self.priors = [tuple(np.asarray(p, dtype=np.float64) for p in pr) for pr in priors]
        self.bases = bases                      # bases[g](x (n,n_x), u) -> (n,M_g)
        self.T = self.obs.shape[0]
        self.G = len(bases)


def _stats(xi, phi):
    """vmap(prior_mniw_calcStatistics)(xi (n,n_xi), phi (n,M)) (src/BayesianInferrence.py:53-61)"""
    return (np.einsum("nm,nk->nmk", phi, xi), np.einsum("nm,nl->nml", phi, phi), np.einsum("nk,nl->nkl", xi, xi),
            np.ones(phi.shape[0]))


def _gp_mean(model, g, st):
    """vmap(prior_mniw_mean)(prior0 + T0, prior1 + T1) (src/Algorithm1.py:211-217)"""
    p = model.priors[g]
    return np.stack([mniw.prior_mniw_mean(p[0] + st[0][i], p[1] + st[1][i]) for i in range(st[0].shape[0])])


def _draw_xi(model, g, st, phi, tvar):
    """_draw_int_vars for one GP (src/Algorithm1.py:235-274): natural -> standard -> predictive -> draw"""
    p = model.priors[g]
    n = phi.shape[0]
    out = np.zeros((n, st[2].shape[1]))
    for i in range(n):
        std = mniw.prior_mniw_2naturalPara_inv(p[0] + st[0][i], p[1] + st[1][i], p[2] + st[2][i], p[3] + st[3][i])
        pm, pc, pr, pdf = mniw.prior_mniw_Predictive(*std, phi[i])
        out[i] = mniw.prior_mniw_drawPred(tvar[i], pm, pc, pr, pdf)
    return out


def init_particles(model, N, Z0, ZXI0):
    """Algorithm1._init_algorithm (src/Algorithm1.py:100-177)"""
    x = model.m0 + Z0 @ np.linalg.cholesky(model.P0).T
    xi = [model.xi_mean[g] + ZXI0[g] @ np.linalg.cholesky(model.xi_cov[g]).T for g in range(model.G)]
    st = [_stats(xi[g], model.bases[g](x, model.inputs[0])) for g in range(model.G)]
    return x, xi, st


def alg1_step(model, t, logw, x, xi, st, lam, u_res, z, tvar):
    """Algorithm1.step (src/Algorithm1.py:298-397)"""
    G = model.G
    st = [tuple(s * lam for s in st[g]) for g in range(G)]                                   # :315-318
    aux_x = model.ssm.transition(x, model.inputs[t - 1], *xi)                                 # :206-208
    aux_xi = []
    for g in range(G):
        A = _gp_mean(model, g, st[g])                                                          # (n, n_xi, M)
        phi = model.bases[g](aux_x, model.inputs[t])
        aux_xi.append(np.einsum("ikj,ij->ik", A, phi))                                         # :228-231
    ll_aux = model.ssm.log_likelihood(model.obs[t], aux_x, model.inputs[t], *aux_xi)          # :326-343
    a = F.systematic_SISR(u_res, F.softmax(ll_aux + logw))                                    # :344-349
    new_x = model.ssm.draw_state(z, x[a], model.inputs[t - 1], *[v[a] for v in xi])           # :276-296
    st_a = [tuple(s[a] for s in st[g]) for g in range(G)]                                     # :356-359
    new_xi, new_st = [], []
    for g in range(G):
        phi = model.bases[g](new_x, model.inputs[t])
        v = _draw_xi(model, g, st_a[g], phi, tvar[g])
        new_xi.append(v)
        Tn = _stats(v, phi)                                                                    # :368-375
        new_st.append(tuple(st_a[g][j] + Tn[j] for j in range(4)))
    new_logw = model.ssm.log_likelihood(model.obs[t], new_x, model.inputs[t], *new_xi) - ll_aux[a]   # :378-389
    return new_logw, new_x, new_xi, new_st, a


def alg1_run(model, N, lam, V):
    """Algorithm1.__call__ (src/Algorithm1.py:399-492).  V: dict Z (T,N,n_x), ZXI0 [g](N,n_xi), U (T,), TS [g](T,N,n_xi)."""
    T, G = model.T, model.G
    x, xi, st = init_particles(model, N, V["Z"][0], V["ZXI0"])
    xs, xis = np.zeros((T, N, x.shape[1])), [np.zeros((T, N, xi[g].shape[1])) for g in range(G)]
    lws, anc = np.zeros((T, N)), np.zeros((T - 1, N), dtype=np.int64)
    sst = [[np.zeros((T,) + model.priors[g][j].shape) if j < 3 else np.zeros(T) for j in range(4)] for g in range(G)]
    xs[0] = x
    w = F.softmax(lws[0])
    for g in range(G):
        xis[g][0] = xi[g]
        for j in range(4):
            sst[g][j][0] = np.einsum("n...,n->...", st[g][j], w)                              # :165-169
    for t in range(1, T):
        u_res = V["U"][t, 0] if np.ndim(V["U"]) == 2 else V["U"][t]
        lw, x, xi, st, a = alg1_step(model, t, lws[t - 1], xs[t - 1], [v[t - 1] for v in xis], st, lam, u_res,
                                     V["Z"][t], [V["TS"][g][t] for g in range(G)])
        xs[t], lws[t], anc[t - 1] = x, lw, a
        w = F.softmax(lw)
        for g in range(G):
            xis[g][t] = xi[g]
            for j in range(4):
                sst[g][j][t] = np.einsum("n...,n->...", st[g][j], w)                          # :445-457
    weights = np.stack([F.softmax(l) for l in lws])                                           # :460
    return dict(state_trace=xs, int_var_trace=xis, suff_stats_trace=sst, weights_trace=weights, ancestor_trace=anc,
                suff_stats=st, logw_trace=lws)


def log_base_measure_batch(T0, T1, T2, T3):
    """vmap(prior_mniw_log_base_measure) (src/BayesianInferrence.py:111-124); T3 scalar or (n,)"""
    n_b = T0.shape[0]
    T3 = np.broadcast_to(np.asarray(T3, dtype=np.float64), (n_b,))
    return np.array([mniw.prior_mniw_log_base_measure(T0[i], T1[i], T2[i], T3[i]) for i in range(n_b)])


def alg3_step(model, t, logw, x, xi, st, ref_x, ref_xi, ref_st, u_res, u_anc, z, tvar):
    """Algorithm3.step (src/Algorithm3.py:43-197)"""
    G, N = model.G, x.shape[0]
    aux_x = model.ssm.transition(x, model.inputs[t - 1], *xi)
    aux_xi = []
    for g in range(G):
        A = _gp_mean(model, g, st[g])
        phi = model.bases[g](aux_x, model.inputs[t])
        aux_xi.append(np.einsum("ikj,ij->ik", A, phi))
    ll_aux = model.ssm.log_likelihood(model.obs[t], aux_x, model.inputs[t], *aux_xi)
    lw_aux = ll_aux + logw
    a = F.systematic_SISR(u_res, F.softmax(lw_aux))                                           # :88-89
    g_T, g_t = np.zeros(N), np.zeros(N)
    for g in range(G):                                                                         # :94-106
        p = model.priors[g]
        g_T += log_base_measure_batch(p[0] + ref_st[g][0] + st[g][0], p[1] + ref_st[g][1] + st[g][1],
                                      p[2] + ref_st[g][2] + st[g][2], p[3] + ref_st[g][3] + st[g][3])
        g_t += log_base_measure_batch(p[0] + st[g][0], p[1] + st[g][1], p[2] + st[g][2], p[3] + st[g][3])
    h_x = mvn_logpdf(ref_x, aux_x, model.ssm.Q)                                               # :107-114
    w_anc = F.softmax(lw_aux + g_t - g_T + h_x)
    ref_idx = F.categorical_searchsorted(w_anc, u_anc)                                        # :119-121
    a[-1] = ref_idx
    ac = np.clip(a, 0, N - 1)
    new_x = model.ssm.draw_state(z, x[ac], model.inputs[t - 1], *[v[ac] for v in xi])
    new_x[-1] = ref_x                                                                          # :132
    st_a = [tuple(s[ac] for s in st[g]) for g in range(G)]
    new_xi, new_st, new_ref = [], [], []
    for g in range(G):
        phi = model.bases[g](new_x, model.inputs[t])
        v = _draw_xi(model, g, st_a[g], phi, tvar[g])
        v[-1] = np.atleast_1d(ref_xi[g])                                                       # :147-150
        new_xi.append(v)
        Tn = _stats(v, phi)
        new_st.append(tuple(st_a[g][j] + Tn[j] for j in range(4)))
        rphi = model.bases[g](ref_x[None], model.inputs[t])                                    # :163-174
        rT = _stats(np.atleast_2d(ref_xi[g]), rphi)
        new_ref.append(tuple(ref_st[g][j] - rT[j][0] for j in range(4)))
    new_logw = model.ssm.log_likelihood(model.obs[t], new_x, model.inputs[t], *new_xi) - ll_aux[ac]
    return new_logw, new_x, new_xi, new_st, a, new_ref, dict(w_aux=F.softmax(lw_aux), w_anc=w_anc)


def alg3_run(model, N, ref_x, ref_xi, ref_st, V, keep_weights=False):
    """Algorithm3.__call__ (src/Algorithm3.py:199-303).  ref_x (T,n_x), ref_xi [g](T,n_xi), ref_st [g] 4-tuple.
    V: Z (T,N,n_x), ZXI0, U (T,2) [U[t]=(u_res,u_anc), U[0,0]=u_idx], TS [g](T,N,n_xi)."""
    T, G = model.T, model.G
    x, xi, st = init_particles(model, N, V["Z"][0], V["ZXI0"])
    x[-1] = ref_x[0]                                                                           # :221
    st = [list(s) for s in st]
    ref_st = [tuple(np.asarray(s, dtype=np.float64) for s in r) for r in ref_st]
    new_ref = []
    for g in range(G):
        xi[g][-1] = ref_xi[g][0]
        T0 = _stats(np.atleast_2d(ref_xi[g][0]), model.bases[g](ref_x[0][None], model.inputs[0]))
        for j in range(4):
            st[g][j] = st[g][j].copy()
            st[g][j][-1] = T0[j][0]                                                            # :226-231
        new_ref.append(tuple(ref_st[g][j] - T0[j][0] for j in range(4)))                       # :235-246
    st = [tuple(s) for s in st]
    ref_st = new_ref
    xs, xis = np.zeros((T, N, x.shape[1])), [np.zeros((T, N, xi[g].shape[1])) for g in range(G)]
    lws, anc = np.zeros((T, N)), np.zeros((T - 1, N), dtype=np.int64)
    xs[0] = x
    for g in range(G):
        xis[g][0] = xi[g]
    cdfs = []
    for t in range(1, T):
        lw, x, xi, st, a, ref_st, ex = alg3_step(model, t, lws[t - 1], xs[t - 1], [v[t - 1] for v in xis], st, ref_x[t],
                                                 [ref_xi[g][t] for g in range(G)], ref_st, V["U"][t, 0], V["U"][t, 1], V["Z"][t],
                                                 [V["TS"][g][t] for g in range(G)])
        xs[t], lws[t], anc[t - 1] = x, lw, a
        for g in range(G):
            xis[g][t] = xi[g]
        if keep_weights:
            cdfs.append((ex["w_aux"], ex["w_anc"]))
    w = F.softmax(lws[-1])
    idx = F.categorical_searchsorted(w, V["U"][0, 0])                                          # :292-293
    ic = min(idx, N - 1)
    ancc = np.clip(anc, 0, N - 1)
    traj = F.reconstruct_trajectory(xs, ancc, ic).reshape(T, -1)
    xi_traj = [F.reconstruct_trajectory(xis[g], ancc, ic).reshape(T, -1) for g in range(G)]
    return dict(traj=traj, xi_traj=xi_traj, state_trace=xs, int_var_trace=xis, logw_trace=lws, anc_trace=anc, idx=idx, cdfs=cdfs,
                w_final=w, ref_st_end=ref_st)


def reference_stats(model, x_traj, xi_traj):
    """sum over ALL T steps of calcStatistics(xi_t, basis(x_t, u_t)) (src/Algorithm2.py:83-96, :139-152)"""
    out = []
    for g in range(model.G):
        phi = np.stack([model.bases[g](x_traj[t][None], model.inputs[t])[0] for t in range(model.T)])
        Tn = _stats(np.atleast_2d(xi_traj[g].reshape(model.T, -1)), phi)
        out.append(tuple(np.sum(Tn[j], axis=0) for j in range(4)))
    return out


def alg2_run(model, N, K, init_x, init_xi, variates):
    """Algorithm2.__call__ (src/Algorithm2.py:106-187).  variates(k) -> V of alg3_run for sweep k >= 1."""
    T, G = model.T, model.G
    xs = np.zeros((K, T, init_x.shape[1]))
    xis = [np.zeros((K, T, np.atleast_2d(init_xi[g].reshape(T, -1)).shape[1])) for g in range(G)]
    xs[0] = init_x
    for g in range(G):
        xis[g][0] = init_xi[g].reshape(T, -1)
    sst = [reference_stats(model, xs[0], [xis[g][0] for g in range(G)])]
    for k in range(1, K):
        r = alg3_run(model, N, xs[k - 1], [xis[g][k - 1] for g in range(G)], sst[k - 1], variates(k))
        xs[k] = r["traj"]
        for g in range(G):
            xis[g][k] = r["xi_traj"][g]
        sst.append(reference_stats(model, xs[k], [xis[g][k] for g in range(G)]))
    return dict(state_trace=np.swapaxes(xs, 0, 1), int_var_trace=[np.swapaxes(v, 0, 1) for v in xis], suff_stats_trace=sst)


# ------------------------------------------------------------------------------- shipped models (batched NumPy)
def smo_ssm(dt=0.02, m=0.2, Q=None, R=None):
    """single-mass oscillator, src/SingleMassOscillator.py:17-48, 85-107"""
    def dx(x, F, Fsd):
        return np.stack([x[:, 1], (-Fsd + F) / m], axis=1)

    def f(x, F, Fsd):
        Fsd = np.asarray(Fsd, dtype=np.float64).reshape(-1)
        k1 = dx(x, F, Fsd); k2 = dx(x + dt / 2.0 * k1, F, Fsd); k3 = dx(x + dt / 2.0 * k2, F, Fsd); k4 = dx(x + dt * k3, F, Fsd)
        return x + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    return SSM(np.diag([5e-8, 5e-9]) if Q is None else Q, np.array([[1e-3]]) if R is None else R, f,
               lambda x, u, *xi: x[:, 0])


def emps_ssm(dt, M=95.11, Q=None, R=None):
    """EMPS, src/EMPS.py:157-209"""
    def dx(x, tau, Fr):
        return np.stack([x[:, 1], (tau - Fr) / M], axis=1)

    def f(x, tau, Fr):
        Fr = np.asarray(Fr, dtype=np.float64).reshape(-1)
        k1 = dx(x, tau, Fr); k2 = dx(x + dt * k1 / 2, tau, Fr); k3 = dx(x + dt * k2 / 2, tau, Fr); k4 = dx(x + dt * k3, tau, Fr)
        return x + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    return SSM(np.diag([1e-6, 1e-7]) if Q is None else Q, np.array([[1e-4]]) if R is None else R, f, lambda x, u, *xi: x[:, 0])


VEH = dict(m=1720.0, I_zz=1827.5, l_f=1.16, l_r=1.47, g=9.81, mu_x=0.9)


def vehicle_ssm(dt=0.02, Q=None, R=None):
    """vehicle lateral dynamics, src/Vehicle.py:29-128, 195-220"""
    p = VEH
    lt = p["l_f"] + p["l_r"]
    Fzf, Fzr = p["m"] * p["g"] * p["l_r"] / lt, p["m"] * p["g"] * p["l_f"] / lt

    def dvy(x, u, mf, mr):
        return 1 / p["m"] * (Fzf * mf * np.cos(u[0]) + Fzr * mr + Fzf * p["mu_x"] * np.sin(u[0])) - u[1] * x[:, 0]

    def dx(x, u, mf, mr):
        ddpsi = 1 / p["I_zz"] * (p["l_f"] * Fzf * mf * np.cos(u[0]) - p["l_r"] * Fzr * mr + p["l_f"] * Fzf * p["mu_x"] * np.sin(u[0]))
        return np.stack([ddpsi, dvy(x, u, mf, mr)], axis=1)

    def f(x, u, mf, mr):
        mf, mr = np.asarray(mf).reshape(-1), np.asarray(mr).reshape(-1)
        k1 = dx(x, u, mf, mr); k2 = dx(x + dt * k1 / 2.0, u, mf, mr); k3 = dx(x + dt * k2 / 2.0, u, mf, mr); k4 = dx(x + dt * k3, u, mf, mr)
        return x + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)

    def out(x, u, mf, mr):
        mf, mr = np.asarray(mf).reshape(-1), np.asarray(mr).reshape(-1)
        return np.tanh(np.stack([x[:, 0], dvy(x, u, mf, mr)], axis=1))
    return SSM(np.diag([1e-8, 1e-8]) if Q is None else Q, np.diag([0.001 / 180 * np.pi, 1e-3]) if R is None else R, f, out)


def slip_basis(hgp, which, l_f=1.16, l_r=1.47):
    """basis_fcn_f / basis_fcn_r of src/Vehicle.py:146-153 (1-D basis of one slip angle)"""
    def basis(x, u):
        x = np.atleast_2d(x)
        if which == 0:
            al = u[0] - np.arctan((x[:, 1] + x[:, 0] * l_f) / u[1])
        else:
            al = -np.arctan((x[:, 1] - x[:, 0] * l_r) / u[1])
        return hgp.batch(al)
    return basis
