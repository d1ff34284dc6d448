"""Test helpers for the marginalised filters (SURVEY.md 8a group B): build the SAME small problem for
the CPU oracle (oracle/marginal.py) and for the CUDA library (Algorithm1/2/3 host classes), feed both the
same injected variates and compare with the two-tier rule of BASELINE.json (helpers.TIE_TOL / REL_TOL).
"""
import numpy as np

from helpers import PKG, REL_TOL, pkg  # noqa: F401


def make_marg_problem(kind, T=24, N=32, M=None, seed=0):
    """kinds: smo (G=1, D=2, identity GP input, y = x[0]); emps (G=1, D=1 on x[1]); vehicle (G=2, D=1 each,
    arctan-linked slip angles, tanh output, n_y=2).  Returns dict(prod=..., oracle=..., N, T, lam)."""
    from oracle import basis as OB, marginal as OMg, mniw as OM
    BF, SSMm = pkg("BasisFunctions"), pkg("StateSpaceModel")
    BI = pkg("BayesianInferrence")
    rng = np.random.default_rng(seed)
    p = dict(kind=kind, T=T, N=N, lam=0.999)
    if kind == "smo":
        S = pkg("SingleMassOscillator")
        M = M or 12
        dom = np.array([[-7.5, 7.5], [-7.5, 7.5]])
        hgp, sd = BF.generate_Hilbert_BasisFunction(M, dom, 15.0 / M, 100.0)
        ohgp, osd = OB.generate_Hilbert_BasisFunction(M, dom, 15.0 / M, 100.0)
        dt = 0.02
        Q, R = np.diag([5e-8, 5e-9]), np.array([[1e-3]])
        inputs = 1.962 * np.sign(np.sin(np.arange(T) / 5.0 + 0.3))
        ssm = SSMm.StateSpaceModel(Q, R, lambda s, u, *xi: S.f_x(s, u, xi[0], dt), lambda s, u, *xi: S.f_y(s))
        ossm = OMg.smo_ssm(dt=dt, Q=Q, R=R)
        m0, P0 = np.zeros(2), np.diag([1e-4, 1e-4])
        xi_mean, xi_cov = [np.zeros(1)], [np.diag([1e-12])]
        priors = [BI.prior_mniw_2naturalPara(np.zeros((1, M)), np.diag(sd), np.eye(1), 3)]
        opriors = [OM.prior_mniw_2naturalPara(np.zeros((1, M)), np.diag(osd), np.eye(1), 3)]
        bases = [lambda s, u: hgp(s)]
        obases = [lambda x, u: ohgp.batch(x)]
        # data: a noisy roll-out of the true plant
        x = np.zeros((T, 2))
        for t in range(1, T):
            f = S.F_spring(x[t - 1, 0]) + S.F_damper(x[t - 1, 1])
            x[t] = S.f_x(x[t - 1], inputs[t - 1], f, dt)
        Y = x[:, 0] + np.sqrt(R[0, 0]) * rng.normal(size=T)
    elif kind == "emps":
        S = pkg("EMPS")
        M = M or 9
        dom = np.array([-0.2, 0.2])
        hgp, sd = BF.generate_Hilbert_BasisFunction(M, dom, 0.4 / M, 20.0)
        ohgp, osd = OB.generate_Hilbert_BasisFunction(M, dom, 0.4 / M, 20.0)
        dt = 0.01
        Q, R = np.diag([1e-6, 1e-7]), np.array([[1e-4]])
        inputs = 40.0 * np.sin(np.arange(T) / 6.0)
        ssm = SSMm.StateSpaceModel(Q, R, lambda s, u, *xi: S.f_x(s, u, xi[0], dt), lambda s, u, *xi: S.f_y(s))
        ossm = OMg.emps_ssm(dt, Q=Q, R=R)
        m0, P0 = np.array([0.1, 0.0]), np.diag([1e-5, 1e-6])
        xi_mean, xi_cov = [np.zeros(1)], [np.diag([1e-12])]
        priors = [BI.prior_mniw_2naturalPara(np.zeros((1, M)), np.diag(sd), np.eye(1) * 4, 2)]
        opriors = [OM.prior_mniw_2naturalPara(np.zeros((1, M)), np.diag(osd), np.eye(1) * 4, 2)]
        bases = [lambda s, u: hgp(s[1])]
        obases = [lambda x, u: ohgp.batch(np.atleast_2d(x)[:, 1])]
        x = np.zeros((T, 2))
        x[0] = m0
        for t in range(1, T):
            x[t] = S.f_x_linModel(x[t - 1], inputs[t - 1], dt)
        Y = x[:, 0] + np.sqrt(R[0, 0]) * rng.normal(size=T)
    elif kind == "vehicle":
        S = pkg("Vehicle")
        M = M or 8
        dom = np.array([-30 / 180 * np.pi, 30 / 180 * np.pi])
        args = (M, dom, 2 / 180 * np.pi, 50.0, 2, 2)
        hgp, sd = BF.generate_Hilbert_BasisFunction(*args)
        ohgp, osd = OB.generate_Hilbert_BasisFunction(*args)
        dt = 0.02
        Q, R = np.diag([1e-8, 1e-8]), np.diag([0.001 / 180 * np.pi, 1e-3])
        tt = np.arange(T) * dt
        inputs = np.stack([8 / 180 * np.pi * np.sin(2 * np.pi * tt / 0.4), 11.0 * np.ones(T)], axis=1)
        ssm = SSMm.StateSpaceModel(Q, R, lambda s, u, *xi: S.f_x(s, u, xi[0], xi[1], dt), lambda s, u, *xi: S.f_y(s, u, xi[0], xi[1]))
        ossm = OMg.vehicle_ssm(dt=dt, Q=Q, R=R)
        m0, P0 = np.zeros(2), np.diag([1e-4, 1e-4])
        xi_mean, xi_cov = [np.zeros(1), np.zeros(1)], [np.diag([1e-4]), np.diag([1e-4])]
        # df = 0 as shipped (src/Vehicle.py:157-174) makes the very first predictive a Cauchy; keep it
        priors = [list(BI.prior_mniw_2naturalPara(np.zeros((1, M)), np.diag(sd), np.eye(1), 0)) for _ in range(2)]
        opriors = [OM.prior_mniw_2naturalPara(np.zeros((1, M)), np.diag(osd), np.eye(1), 0) for _ in range(2)]
        bases = [lambda s, u: hgp(S.f_alpha(s, u)[0]), lambda s, u: hgp(S.f_alpha(s, u)[1])]
        obases = [OMg.slip_basis(ohgp, 0), OMg.slip_basis(ohgp, 1)]
        x = np.zeros((T, 2))
        Y = np.zeros((T, 2))
        for t in range(T):
            af, ar = S.f_alpha(x[t], inputs[t])
            mf, mr = S.mu_y(af), S.mu_y(ar)
            Y[t] = S.f_y(x[t], inputs[t], mf, mr) + np.sqrt(np.diag(R)) * rng.normal(size=2)
            if t + 1 < T:
                x[t + 1] = S.f_x(x[t], inputs[t], mf, mr, dt)
    elif kind == "plugin":
        # model plug-in: a driven pendulum with a GP friction torque — NOTHING of it is in the coefficient-table
        # families: sin / cos / products of the state in the transition, a non-affine output without a tanh link, a GP input that
        # is a sine of the state.  The product gets the per-particle callables below, the oracle their batched twins.
        M = M or 10
        dom = np.array([-6.0, 6.0])
        hgp, sd = BF.generate_Hilbert_BasisFunction(M, dom, 12.0 / M, 10.0)
        ohgp, osd = OB.generate_Hilbert_BasisFunction(M, dom, 12.0 / M, 10.0)
        dt = 0.02
        Q, R = np.diag([1e-6, 1e-5]), np.array([[1e-3]])
        inputs = 3.0 * np.sin(np.arange(T) / 7.0 + 0.2)
        ssm = SSMm.StateSpaceModel(Q, R, lambda s, u, *xi: plugin_f(s, u, xi[0], dt), lambda s, u, *xi: plugin_g(s, u))
        ossm = OMg.SSM(Q, R, lambda x, u, *xi: np.stack([x[:, 0] + dt * x[:, 1], x[:, 1] + dt * (-9.81 * np.sin(x[:, 0]) - np.reshape(xi[0], -1)
                                                                                                  + u * np.cos(x[:, 0]))], axis=1),
                       lambda x, u, *xi: 1.5 * np.sin(x[:, 0]) + 0.05 * x[:, 1] * x[:, 1])
        m0, P0 = np.array([0.3, 0.0]), np.diag([1e-4, 1e-4])
        xi_mean, xi_cov = [np.zeros(1)], [np.diag([1e-6])]
        priors = [BI.prior_mniw_2naturalPara(np.zeros((1, M)), np.diag(sd), np.eye(1), 3)]
        opriors = [OM.prior_mniw_2naturalPara(np.zeros((1, M)), np.diag(osd), np.eye(1), 3)]
        bases = [lambda s, u: hgp(plugin_z(s))]
        obases = [lambda x, u: ohgp.batch(2.0 * np.sin(np.atleast_2d(x)[:, 0]) + np.atleast_2d(x)[:, 1])]
        x = np.zeros((T, 2))
        x[0] = m0
        Y = np.zeros(T)
        for t in range(T):
            Y[t] = plugin_g(x[t], inputs[t]) + np.sqrt(R[0, 0]) * rng.normal()
            if t + 1 < T:
                x[t + 1] = plugin_f(x[t], inputs[t], np.array([0.4 * x[t, 1] + 0.2 * np.tanh(5.0 * x[t, 1])]), dt)
    else:
        raise ValueError(kind)
    p["prod_kwargs"] = dict(N_samples=N, observations=Y, inputs=inputs, SSM=ssm, init_state_mean=m0, init_state_cov=P0,
                            init_int_var_mean=xi_mean, init_int_var_cov=xi_cov, GP_prior=priors, basis_fcn=bases)
    p["oracle"] = OMg.MargModel(Y, inputs, ossm, m0, P0, xi_mean, xi_cov, opriors, obases)
    p["G"], p["n_x"], p["M"] = len(bases), 2, M
    p["prior_df"] = [float(pr[3]) for pr in priors]
    return p


def plugin_f(s, u, xi, dt):
    return np.hstack([s[0] + dt * s[1], s[1] + dt * (-9.81 * np.sin(s[0]) - xi + u * np.cos(s[0]))])


def plugin_g(s, u):
    return 1.5 * np.sin(s[0]) + 0.05 * s[1] * s[1]


def plugin_z(s):
    return 2.0 * np.sin(s[0]) + s[1]


def predictive_df(prob, lam):
    """degrees of freedom of the Student-t predictive at each step (same for all particles): df_t = prior_df +
    lam * T3_{t-1} + 1 - n_xi with T3_0 = 1, T3_t = lam T3_{t-1} + 1"""
    T, G = prob["T"], prob["G"]
    df = np.zeros((G, T))
    for g in range(G):
        T3 = 1.0
        for t in range(1, T):
            df[g, t] = prob["prior_df"][g] + lam * T3 + 1.0 - 1.0
            T3 = lam * T3 + 1.0
    return df


def make_variates(prob, lam, seed=1, conditional=False, K=None):
    """injected variates; with K a leading iteration axis is added (block 0 unused)"""
    rng = np.random.default_rng(seed)
    T, N, G, nx = prob["T"], prob["N"], prob["G"], prob["n_x"]
    df = predictive_df(prob, lam)
    lead = () if K is None else (K,)
    V = dict(Z=rng.normal(size=lead + (T, N, nx)), ZXI0=rng.normal(size=lead + (G, N)),
             U=rng.uniform(size=lead + (T, 2)))
    TS = np.zeros(lead + (G, T, N))
    for g in range(G):
        for t in range(1, T):
            TS[..., g, t, :] = rng.standard_t(max(df[g, t], 0.5), size=lead + (N,))
    V["TS"] = TS
    return V


def oracle_variates(V, k=None):
    """the dict layout oracle/marginal.py expects (per sweep)"""
    pick = (lambda a: a) if k is None else (lambda a: a[k])
    Z, ZX, U, TS = pick(V["Z"]), pick(V["ZXI0"]), pick(V["U"]), pick(V["TS"])
    G = ZX.shape[0]
    return dict(Z=Z, ZXI0=[ZX[g][:, None] for g in range(G)], U=U, TS=[TS[g][:, :, None] for g in range(G)])


def device_variates(V, n_chains_axis=True):
    import torch
    out = {}
    for k, a in V.items():
        a = np.ascontiguousarray(a)
        if n_chains_axis:
            # (..) -> add the chain axis after an optional leading K axis
            a = a[None] if a.ndim == {"Z": 3, "ZXI0": 2, "U": 2, "TS": 3}[k] else a[:, None]
        out[k] = torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
    return out


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)
