"""Shared test helpers: build the same small problem for the CPU oracle and the CUDA library, and
compare them under injected variates with the two-tier rule of BASELINE.json:
  * ancestor / resample indices exact, except at CDF ties within TIE_TOL of the oracle's CDF;
  * float64 states, log-weights, statistics and draws within REL_TOL (relative, normwise).
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "bayesian_inference_with_explicit_and_implicit_prior_knowledge_b200"
TIE_TOL = 1e-12      # BASELINE.json north_star: "except at CDF ties within 1e-12"
REL_TOL = 1e-9       # BASELINE.json north_star: "within 1e-9 relative"


def pkg(name=None):
    return importlib.import_module(PKG if name is None else PKG + "." + name)


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# ------------------------------------------------------------------------------ problems
def make_problem(kind, T=30, N=64, M=None, seed=0, flags=0):
    """Returns a dict with the oracle model, the constructor kwargs of the product classes and a
    synthetic reference trajectory / prior.  kinds: smo (D=2,n_x=2,n_y=1), emps (D=3 with input),
    toy (D=1,n_x=1), vehicle (slip map, D=2, n_y=2, lattice 2,4,..), smo1d (n_x=2, D=1 on x[1])."""
    from oracle import basis as OB, mniw as OM, pgas as OP
    rng = np.random.default_rng(seed)
    p = dict(kind=kind, T=T, N=N, flags=flags)
    if kind == "smo":
        M = M or 41
        dom = np.array([[-7.5, 7.5], [-7.5, 7.5]])
        args = (M, dom, 15.0 / M, 100.0)
        n_x, n_u = 2, 1
        A, b = np.array([[1.0, 0, 0], [0, 1.0, 0]]), np.zeros(2)
        H, h0, R = np.array([[1.0, 0.0]]), np.zeros(1), np.array([[1e-3]])
        m0, P0 = np.zeros(2), np.diag([1e-2, 1e-2])
        inputs = 0.5 * rng.normal(size=(T, 1))
        ref = np.cumsum(0.05 * rng.normal(size=(T, 2)), axis=0)
        df = 3
    elif kind == "emps":
        M = M or 64
        dom = np.array([[-1.0, 1.0]] * 3)
        args = (M, dom, 0.5 / 9, 20.0)
        n_x, n_u = 2, 1
        A, b = np.diag(1.0 / np.array([0.4, 0.4, 160.0])), np.zeros(3)
        H, h0, R = np.array([[1.0, 0.0]]), np.zeros(1), np.array([[1e-4]])
        m0, P0 = np.array([0.05, 0.0]), np.diag([1e-5, 1e-6])
        inputs = 60.0 * np.sin(np.arange(T) / 7.0)[:, None] + 5.0 * rng.normal(size=(T, 1))
        ref = np.stack([0.05 + 0.1 * np.sin(np.arange(T) / 9.0), 0.1 * np.cos(np.arange(T) / 9.0)], axis=1)
        df = 2
    elif kind == "toy":
        M = M or 40
        dom = np.array([-30.0, 30.0])
        args = (M, dom, 3.0, 50.0)
        n_x, n_u = 1, 0
        A, b = np.array([[1.0]]), np.zeros(1)
        H, h0, R = np.array([[1.0]]), np.zeros(1), np.array([[4.0]])
        m0, P0 = np.zeros(1), np.diag([1e-4])
        inputs = np.zeros((T, 0))
        ref = 5.0 * rng.normal(size=(T, 1))
        df = 10
    elif kind == "vehicle":
        M = M or 36
        dom = np.array([[-0.5, 0.5], [-0.5, 0.5]])
        args = (M, dom, 2.0 / 180 * np.pi, 50.0, 2, 2)
        n_x, n_u = 2, 2
        A, b = None, None
        H, h0 = np.array([[1.0, 0.0], [0.3, 1.0]]), np.array([0.0, 0.01])
        R = np.array([[1e-3, 2e-4], [2e-4, 2e-3]])
        m0, P0 = np.zeros(2), np.diag([1e-4, 1e-4])
        tt = np.arange(T) * 0.02
        inputs = np.stack([10 / 180 * np.pi * np.sin(2 * np.pi * tt / 5), 11.0 + 0 * tt], axis=1)
        ref = np.stack([0.2 * np.sin(tt), 0.5 * np.cos(tt)], axis=1)
        df = 3
    elif kind == "plugin":
        # a GP-input map outside the compiled-in families (model plug-in): sines, tanh, a rational term
        M = M or 41
        dom = np.array([[-7.5, 7.5], [-7.5, 7.5]])
        args = (M, dom, 15.0 / M, 100.0)
        n_x, n_u = 2, 1
        A, b = None, None
        H, h0, R = np.array([[1.0, 0.0]]), np.zeros(1), np.array([[1e-3]])
        m0, P0 = np.zeros(2), np.diag([1e-2, 1e-2])
        inputs = 0.5 * rng.normal(size=(T, 1))
        ref = np.cumsum(0.05 * rng.normal(size=(T, 2)), axis=0)
        df = 3
    elif kind == "pluginlik":
        # a likelihood_fcn outside the Gaussian-of-an-affine-map family (model plug-in): Student-t log-density (4 degrees of freedom) of
        # the observation around a non-affine output map that also reads the input; the basis is the single-mass oscillator's
        M = M or 41
        dom = np.array([[-7.5, 7.5], [-7.5, 7.5]])
        args = (M, dom, 15.0 / M, 100.0)
        n_x, n_u = 2, 1
        A, b = np.array([[1.0, 0, 0], [0, 1.0, 0]]), np.zeros(2)
        H, h0, R = np.array([[1.0, 0.0]]), np.zeros(1), np.array([[1e-2]])
        m0, P0 = np.zeros(2), np.diag([1e-2, 1e-2])
        inputs = 0.5 * rng.normal(size=(T, 1))
        ref = np.cumsum(0.05 * rng.normal(size=(T, 2)), axis=0)
        df = 3
    else:
        raise ValueError(kind)
    hgp, sd = OB.generate_Hilbert_BasisFunction(*args)
    M = hgp.indices.shape[0]
    obs = ref @ H.T + h0 + rng.normal(size=(T, H.shape[0])) * np.sqrt(np.diag(R))
    if kind == "vehicle":
        obasis = OP.vehicle_slip_basis(hgp, 1.16, 1.47)
    elif kind == "plugin":
        def obasis(states, inp):                 # the user callable evaluated directly with numpy, particle by column
            st = np.atleast_2d(states)
            z = plugin_map([st[:, 0], st[:, 1]], [np.full(st.shape[0], float(np.ravel(inp)[0]))])
            return hgp.batch(np.stack(z, axis=1))
    else:
        obasis = OP.affine_hgp_basis(hgp, A, b)
    Sg = rng.normal(size=(n_x, n_x))
    ologlik = OP.gaussian_loglik(H, h0, R)
    if kind == "pluginlik":
        def ologlik(y, states, inp):             # the user callable evaluated directly with numpy, particle by column
            st = np.atleast_2d(states)
            return plugin_loglik(y, [st[:, 0], st[:, 1]], np.ravel(inp))
    omodel = OP.ThetaModel(obs, inputs, m0, P0, obasis, ologlik)
    prior = OM.prior_mniw_2naturalPara(np.zeros((n_x, M)), np.diag(sd), np.eye(n_x), df)
    Theta = 0.3 * rng.normal(size=(n_x, M)) / np.sqrt(M)
    # Under reference quirk (i) every particle free-runs x_t = Theta phi(x_{t-1}) + noise for all T steps, so
    # a map with Lipschitz constant > 1 amplifies last-bit differences exponentially and no two float64
    # implementations agree after a few dozen steps.  Keep the synthetic dynamics contractive (|J| <= 0.5).
    probe = ref[rng.integers(0, T, size=64)] + 0.05 * rng.normal(size=(64, n_x))
    h = 1e-6
    lip = 0.0
    for k in range(n_x):
        dphi = (obasis(probe + h * np.eye(n_x)[k], inputs[1]) - obasis(probe - h * np.eye(n_x)[k], inputs[1])) / (2 * h)
        lip = max(lip, float(np.max(np.linalg.norm(dphi @ Theta.T, axis=1))))
    Theta *= min(1.0, 0.5 / (lip * np.sqrt(n_x) + 1e-30))
    Sigma = 0.02 * (Sg @ Sg.T + n_x * np.eye(n_x))
    p.update(omodel=omodel, prior=prior, Theta=Theta, Sigma=Sigma, ref=ref, obs=obs, inputs=inputs, m0=m0, P0=P0,
             H=H, h0=h0, R=R, A=A, b=b, hgp_args=args, n_x=n_x, n_u=n_u, M=M, ohgp=hgp, sd=sd, df=df)
    return p


def plugin_map(state, inp):
    """GP-input map of the "plugin" problems, written the way a user writes a basis_fcn: plain numpy on state / input components"""
    z0 = 2.0 * np.sin(state[0]) + 0.3 * state[1] + 0.5 * inp[0]
    z1 = 3.0 * np.tanh(state[1]) - state[0] ** 2 / (1.0 + np.abs(state[0]))
    return [z0, z1]


def plugin_loglik(obs, state, inp):
    """likelihood_fcn of the "pluginlik" problems, written the way a user writes it: plain numpy on observation / state / input"""
    e = obs[0] - (state[0] + 0.2 * np.sin(state[1]) + 0.1 * inp[0])
    return -2.5 * np.log(1.0 + e * e / (4.0 * 1e-2)) - 1.9


def _product_likelihood(p, MD):
    if p["kind"] == "pluginlik":
        return plugin_loglik
    return MD.GaussianLikelihood(p["H"], p["h0"], p["R"])


def _product_basis(hgp, p, MD):
    if p["kind"] == "vehicle":
        return MD.VehicleSlipBasis(hgp, 1.16, 1.47)
    if p["kind"] == "plugin":
        return lambda state, inp: hgp(MD.hstack(plugin_map(state, inp)))
    return _BasisThunk(hgp, p, MD)


def product_csmc(p, cluster_size=0):
    """condSequentialMonteCarlo of the product for problem p."""
    BF, MD, PG = pkg("BasisFunctions"), pkg("models"), pkg("PGAS")
    hgp, _ = BF.generate_Hilbert_BasisFunction(*p["hgp_args"])
    basis = _product_basis(hgp, p, MD)
    lik = _product_likelihood(p, MD)
    return PG.condSequentialMonteCarlo(N_samples=p["N"], observations=p["obs"], inputs=p["inputs"],
                                       init_state_mean=p["m0"], init_state_cov=p["P0"], likelihood_fcn=lik,
                                       basis_fcn=basis, flags=p["flags"], cluster_size=cluster_size)


def _affine_apply(MD, state, inp, A, b):
    v = MD.hstack([state, inp]) if len(inp) else state
    rows = [sum(v[k] * A[d, k] for k in range(A.shape[1])) + b[d] for d in range(A.shape[0])]
    return MD.hstack(rows)


class _BasisThunk:
    """a user-style basis_fcn(state, input) built from ordinary arithmetic on the traced values"""

    def __init__(self, hgp, p, MD):
        self.hgp, self.p, self.MD = hgp, p, MD

    def __call__(self, state, inp):
        return self.hgp(_affine_apply(self.MD, state, inp, self.p["A"], self.p["b"]))


def product_pgas(p, K, cluster_size=0):
    BF, MD, PG = pkg("BasisFunctions"), pkg("models"), pkg("PGAS")
    hgp, _ = BF.generate_Hilbert_BasisFunction(*p["hgp_args"])
    basis = _product_basis(hgp, p, MD)
    lik = _product_likelihood(p, MD)
    return PG.PGAS(N_samples=p["N"], N_iterations=K, observations=p["obs"], inputs=p["inputs"], init_state_mean=p["m0"],
                   init_state_cov=p["P0"], likelihood_fcn=lik, GP_prior=p["prior"], basis_fcn=basis, flags=p["flags"],
                   cluster_size=cluster_size)


def sweep_variates(p, seed=1):
    rng = np.random.default_rng(seed)
    return rng.normal(size=(p["T"], p["N"], p["n_x"])), rng.uniform(size=(p["T"], 2))


def oracle_flags(p):
    return dict(ancestor_gather=bool(p["flags"] & 1), input_offset=-1 if p["flags"] & 2 else 0)


# ------------------------------------------------------------------------------ comparisons
def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def resample_cdf(w):
    """the CDF systematic_SISR searches (src/Filtering.py:23-32)"""
    w = np.clip(w, 0, np.inf)
    return np.clip(np.cumsum(w / np.sum(w)), 0.0, 1.0)


def index_mismatch_is_tie(idx_gpu, idx_ref, cdf, points, tol=TIE_TOL):
    """every disagreement must sit within `tol` of a CDF value between the two candidate indices"""
    bad = np.nonzero(np.asarray(idx_gpu) != np.asarray(idx_ref))[0]
    worst = 0.0
    for j in bad:
        lo, hi = sorted((int(idx_gpu[j]), int(idx_ref[j])))
        hi = min(hi, len(cdf) - 1)
        gap = np.min(np.abs(cdf[lo:hi + 1] - points[j])) if hi >= lo else np.inf
        worst = max(worst, gap)
        if gap > tol:
            return False, len(bad), worst
    return True, len(bad), worst


def compare_step(p, o_step, g_step, u_res, u_anc):
    """o_step = oracle (logw, state, a, extras); g_step = product (logw, state, a) numpy."""
    N = p["N"]
    lw_o, x_o, a_o, ex = o_step
    lw_g, x_g, a_g = g_step
    cdf = resample_cdf(ex["w_aux"])
    pts = (u_res + np.arange(N)) / N
    ok, nbad, worst = index_mismatch_is_tie(a_g[:-1], a_o[:-1], cdf, pts[:-1])
    cdf_r = np.cumsum(ex["w_anc"])
    ok_r, nbad_r, worst_r = index_mismatch_is_tie(a_g[-1:], a_o[-1:], cdf_r, np.array([u_anc]))
    same = np.asarray(a_g) == np.asarray(a_o)
    e_x = rel_err(x_g, x_o)
    e_lw = float(np.max(np.abs(lw_g[same] - lw_o[same]) / np.maximum(1.0, np.abs(lw_o[same])))) if same.any() else 0.0
    return dict(ok=bool(ok and ok_r and e_x < REL_TOL and e_lw < REL_TOL), idx_mismatch=nbad + nbad_r,
                tie_gap=max(worst, worst_r), state_err=e_x, logw_err=e_lw)


def run_step_parity(p, n_steps=10, cluster_size=0, seed=5):
    """teacher-forced single steps: oracle state in, one step on both sides"""
    import torch
    from oracle import pgas as OP
    cs = product_csmc(p, cluster_size)
    rng = np.random.default_rng(seed)
    N, n_x, T = p["N"], p["n_x"], p["T"]
    state = p["m0"] + rng.normal(size=(N, n_x)) @ np.linalg.cholesky(p["P0"]).T
    state[-1] = p["ref"][0]
    logw = np.zeros(N)
    worst = dict(ok=True, idx_mismatch=0, tie_gap=0.0, state_err=0.0, logw_err=0.0)
    for s in range(n_steps):
        t = 1 + (s % (T - 1))
        u = rng.uniform(size=2)
        z = rng.normal(size=(N, n_x))
        o = OP.csmc_step(p["omodel"], t, logw, state, p["Theta"], p["Sigma"], p["ref"][t], u[0], u[1], z, **oracle_flags(p))
        g = cs.step(t, logw, state, p["Theta"], p["Sigma"], p["ref"][t], u, z)
        g = tuple(v.cpu().numpy() for v in g)
        r = compare_step(p, o, g, u[0], u[1])
        worst["ok"] &= r["ok"]
        worst["idx_mismatch"] += r["idx_mismatch"]
        for k in ("tie_gap", "state_err", "logw_err"):
            worst[k] = max(worst[k], r[k])
        logw, state = o[0], o[1]
    return worst


def run_sweep_parity(p, cluster_size=0, seed=1, philox_seed=None):
    """full sweep under injected variates (or the library's Philox stream fed to the oracle);
    compared row by row up to the first index disagreement, which must be a CDF tie."""
    import torch
    from oracle import pgas as OP, philox as OPH
    cs = product_csmc(p, cluster_size)
    T, N, n_x = p["T"], p["N"], p["n_x"]
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    if philox_seed is None:
        Z, U = sweep_variates(p, seed)
        g = cs.sweep(dev(p["ref"]), dev(p["Theta"]), dev(p["Sigma"]), variates=dict(Z=dev(Z[None]), U=dev(U[None])))
    else:
        Z, U = OPH.sweep_variates(philox_seed, 2, 3, T, N, n_x)
        g = cs.sweep(dev(p["ref"]), dev(p["Theta"]), dev(p["Sigma"]), key=pkg("random").key(philox_seed), chain_base=2,
                     iteration=3)
    o = OP.csmc_sweep(p["omodel"], N, p["ref"], p["Theta"], p["Sigma"], Z, U, keep_weights=True, **oracle_flags(p))
    st_g = g["state_trace"][0].cpu().numpy()
    an_g = g["anc_trace"][0].cpu().numpy()
    res = dict(ok=True, rows_compared=0, idx_mismatch=0, tie_gap=0.0, state_err=0.0, diverged_at=None)
    res["state_err"] = rel_err(st_g[0], o["state_trace"][0])
    for t in range(1, T):
        a_o, a_g = o["anc_trace"][t - 1], an_g[t - 1]
        w_aux, w_anc = o["cdfs"][t - 1]
        pts = (U[t, 0] + np.arange(N)) / N
        ok1, n1, g1 = index_mismatch_is_tie(a_g[:-1], a_o[:-1], resample_cdf(w_aux), pts[:-1])
        ok2, n2, g2 = index_mismatch_is_tie(a_g[-1:], a_o[-1:], np.cumsum(w_anc), np.array([U[t, 1]]))
        res["idx_mismatch"] += n1 + n2
        res["tie_gap"] = max(res["tie_gap"], g1, g2)
        res["state_err"] = max(res["state_err"], rel_err(st_g[t], o["state_trace"][t]))
        res["rows_compared"] = t
        if not (ok1 and ok2):
            res["ok"] = False
            break
        if n1 + n2:                    # a legitimate tie flipped: later rows follow different weights
            res["diverged_at"] = t
            break
    if res["state_err"] > REL_TOL:
        res["ok"] = False
    if res["diverged_at"] is None and res["ok"]:
        lw_err = float(np.max(np.abs(g["logw_last"][0].cpu().numpy() - o["logw_trace"][-1])
                              / np.maximum(1.0, np.abs(o["logw_trace"][-1]))))
        res["logw_err"] = lw_err
        idx_g = int(g["idx"][0])
        cdf_f = np.cumsum(o["w_final"])
        okf, nf, gf = index_mismatch_is_tie(np.array([idx_g]), np.array([o["idx"]]), cdf_f, np.array([U[0, 0]]))
        res["ok"] &= bool(okf and lw_err < REL_TOL)
        if nf == 0:
            res["traj_err"] = rel_err(g["traj"][0].cpu().numpy(), o["traj"])
            res["ok"] &= res["traj_err"] < REL_TOL
    return res


def run_draw_parity(p, seed=2, flags=None):
    """sufficient statistics + MNIW draw of the reference trajectory under injected variates"""
    import torch
    from oracle import pgas as OP
    pg = product_pgas(p, K=2)
    rng = np.random.default_rng(seed)
    n_x, M, T = p["n_x"], p["M"], p["T"]
    df = p["prior"][3] + T - 1
    chi2 = rng.chisquare(df - np.arange(n_x))
    G = rng.normal(size=(n_x, n_x))
    Nrm = rng.normal(size=(n_x, M))
    A_o, S_o, ex = OP.sample_params(p["omodel"], p["prior"], p["ref"], chi2, G, Nrm, transpose_fix=bool(p["flags"] & 4))
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    T0, T1, T2, T3 = pkg("BayesianInferrence").trajectory_statistics(pg.cSMC.model, dev(p["ref"][None]))
    A_g, S_g = pg.sample_params(None, dev(p["ref"][None]), variates=dict(chi2=dev(chi2[None]), G=dev(G[None]), Nrm=dev(Nrm[None])))
    res = dict(T0_err=rel_err(T0[0].cpu().numpy(), ex["stats"][0]), T1_err=rel_err(T1[0].cpu().numpy(), ex["stats"][1]),
               T2_err=rel_err(T2[0].cpu().numpy(), ex["stats"][2]), A_err=rel_err(A_g[0].cpu().numpy(), A_o),
               S_err=rel_err(S_g[0].cpu().numpy(), S_o))
    res["ok"] = all(v < REL_TOL for v in res.values())
    return res
