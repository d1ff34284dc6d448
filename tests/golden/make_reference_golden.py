"""Regenerates tests/golden/reference_golden.npz by EXECUTING THE REFERENCE'S OWN SOURCE.

The reference (/root/reference, pure Python) needs jax 0.4.38 / equinox 0.12.2, which are not installable
offline.  tests/golden/jaxshim/ is a NumPy/SciPy-backed stand-in for the ~60 jax / equinox entry points its
src/*.py call (see jaxshim/README.md); with it on sys.path the UNMODIFIED modules src/Filtering.py,
src/BasisFunctions.py, src/BayesianInferrence.py, src/StateSpaceModel.py, src/PGAS.py, src/Algorithm1/2/3.py,
src/SingleMassOscillator.py and src/Vehicle.py are imported from /root/reference and run on small seeded
problems.  `jax.random` of the stand-in is a tape recorder: every variate the reference consumes is logged,
sliced into the injected-variates layout of SURVEY.md section 8c and stored next to the reference's outputs.

Only this script needs /root/reference (it exists in the build container only); the tests read the .npz.
Run:  python tests/golden/make_reference_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402
import helpers_marginal as HM  # noqa: E402

REF = os.environ.get("PGAS_REFERENCE_DIR", "/root/reference")
sys.path.insert(0, REF)                                  # `import src` must resolve to the reference, not the repo's src/
sys.path.insert(0, os.path.join(HERE, "jaxshim"))
for name in [n for n in sys.modules if n == "src" or n.startswith("src.")]:
    del sys.modules[name]

import jax  # noqa: E402  (the stand-in)
import jax.numpy as jnp  # noqa: E402
import jax.scipy as jsp  # noqa: E402
import src  # noqa: E402
assert os.path.realpath(os.path.dirname(src.__file__)).startswith(os.path.realpath(REF)), src.__file__
import src.Filtering as RF  # noqa: E402
import src.BasisFunctions as RB  # noqa: E402
import src.BayesianInferrence as RBI  # noqa: E402
import src.PGAS as RP  # noqa: E402
from src.StateSpaceModel import StateSpaceModel as RSSM  # noqa: E402
from src.Algorithm1 import Algorithm1 as RA1  # noqa: E402
from src.Algorithm2 import Algorithm2 as RA2  # noqa: E402
from src.Algorithm3 import Algorithm3 as RA3  # noqa: E402
import src.SingleMassOscillator as RSMO  # noqa: E402
import src.Vehicle as RVEH  # noqa: E402

TAPE = jax.random.TAPE
OUT = {}


def put(name, value):
    OUT[name] = np.asarray(value)


class Reader:
    """sequential reader of the variate tape"""

    def __init__(self):
        self.pos = 0

    def take(self, kind, shape=None):
        k, v = TAPE[self.pos]
        assert k == kind, (self.pos, k, kind)
        if shape is not None:
            assert v.shape == tuple(shape), (self.pos, kind, v.shape, shape)
        self.pos += 1
        return v

    def take_n(self, kind, n, shape):
        return np.stack([self.take(kind, shape) for _ in range(n)])

    def done(self):
        assert self.pos == len(TAPE), (self.pos, len(TAPE))


# ------------------------------------------------------------------------------------ A1: systematic_SISR
def gen_sisr():
    rng = np.random.default_rng(101)
    cases = [rng.uniform(size=16), np.ones(9), np.zeros(7), np.array([0.0, 0.0, 5.0, 0.0]),
             np.array([1.0, np.nan, 2.0, 1.0, 0.5]), np.exp(-8 * rng.uniform(size=64)), np.array([-1.0, 2.0, 1.0])]
    for i, w in enumerate(cases):
        jax.random.tape_reset()
        idx = RF.systematic_SISR(jax.random.key(500 + i), jnp.array(w))
        put(f"sisr/{i}/w", w)
        put(f"sisr/{i}/u", TAPE[0][1])
        put(f"sisr/{i}/idx", np.asarray(idx, dtype=np.int64))
    put("sisr/n", len(cases))
    # reconstruct_trajectory
    P = rng.normal(size=(6, 5, 2))
    anc = rng.integers(0, 5, size=(6, 5)).astype(np.float64)
    put("recon/P", P)
    put("recon/anc", anc)
    put("recon/traj", RF.reconstruct_trajectory(P, anc, 3))
    put("recon/traj_1d", RF.reconstruct_trajectory(P[:, :, 0], anc, 1))


# ------------------------------------------------------------------------------------ A3-A5: basis
BASIS_CFGS = {
    "smo41": dict(num_fcn=41, domain_boundary=np.array([[-7.5, 7.5], [-7.5, 7.5]]), lengthscale=15.0 / 41, scale=100),
    "vehicle20": dict(num_fcn=20, domain_boundary=np.array([-30 / 180 * np.pi, 30 / 180 * np.pi]), lengthscale=2 / 180 * np.pi,
                      scale=50, idx_start=2, idx_step=2),
    "emps9": dict(num_fcn=9, domain_boundary=np.array([-0.2, 0.2]), lengthscale=0.4 / 9, scale=20),
    "emps729": dict(num_fcn=729, domain_boundary=np.array([[-1, 1], [-1, 1], [-1, 1]]), lengthscale=0.5 / 729, scale=20),
    "smo256": dict(num_fcn=256, domain_boundary=np.array([[-7.5, 7.5], [-7.5, 7.5]]), lengthscale=15.0 / 256, scale=100),
    "aniso30": dict(num_fcn=30, domain_boundary=np.array([[-1.0, 2.0], [-4.0, 1.0]]), lengthscale=np.array([0.3, 0.7]), scale=3.0),
}


def gen_basis():
    rng = np.random.default_rng(102)
    for name, kw in BASIS_CFGS.items():
        fn, sd = RB.generate_Hilbert_BasisFunction(**kw)
        dom = np.atleast_2d(kw["domain_boundary"]).astype(float)
        D = dom.shape[0]
        size = dom[:, 1] - dom[:, 0]
        pts = dom[:, 0] + size * rng.uniform(0.05, 0.95, size=(4, D))
        phi = np.stack([np.asarray(fn(jnp.array(x if D > 1 else x[0]))) for x in pts])
        put(f"basis/{name}/pts", pts)
        put(f"basis/{name}/phi", phi)
        put(f"basis/{name}/sd", np.asarray(sd))
    put("basis/names", np.array(list(BASIS_CFGS)))


# ------------------------------------------------------------------------------------ A15-A17, B2-B5: MNIW algebra
def gen_mniw():
    rng = np.random.default_rng(103)
    M, n = 7, 2
    mean = rng.normal(size=(n, M))
    Vh = rng.normal(size=(M, M))
    V = Vh @ Vh.T + M * np.eye(M)
    Ph = rng.normal(size=(n, n))
    Psi = Ph @ Ph.T + n * np.eye(n)
    eta = RBI.prior_mniw_2naturalPara(jnp.array(mean), jnp.array(V), jnp.array(Psi), 5)
    put("mniw/mean", mean); put("mniw/V", V); put("mniw/Psi", Psi)
    for j in range(3):
        put(f"mniw/eta{j}", eta[j])
    back = RBI.prior_mniw_2naturalPara_inv(*eta)
    for j in range(3):
        put(f"mniw/inv{j}", back[j])
    put("mniw/gp_mean", RBI.prior_mniw_mean(eta[0], eta[1]))
    y, phi = rng.normal(size=n), rng.normal(size=M)
    st = RBI.prior_mniw_calcStatistics(jnp.array(y), jnp.array(phi))
    put("mniw/y", y); put("mniw/phi", phi)
    for j in range(3):
        put(f"mniw/stat{j}", st[j])
    pred = RBI.prior_mniw_Predictive(back[0], back[1], back[2], back[3], jnp.array(phi))
    for j in range(4):
        put(f"mniw/pred{j}", pred[j])
    jax.random.tape_reset()
    draw = RBI.prior_mniw_drawPred(jax.random.key(7), *pred)
    put("mniw/t", TAPE[0][1]); put("mniw/draw", draw)
    T1 = eta[1] + np.outer(phi, phi)
    T0 = eta[0] + np.outer(phi, y)
    T2 = eta[2] + np.outer(y, y) + 3 * np.eye(n)
    put("mniw/lbm_T0", T0); put("mniw/lbm_T1", T1); put("mniw/lbm_T2", T2)
    put("mniw/lbm", RBI.prior_mniw_log_base_measure(jnp.array(T0), jnp.array(T1), jnp.array(T2), 9.0))


# ------------------------------------------------------------------------------------ group A: cSMC / PGAS
CSMC_CASES = {"smo": dict(T=14, N=24, seed=21), "emps": dict(T=12, N=20, seed=22), "toy": dict(T=16, N=16, seed=23),
              "vehicle": dict(T=12, N=32, seed=24)}


def reference_callables(p):
    """user callables in the style of the shipped modules (src/EMPS.py:110-113,250-252; src/Toy_Example.py:142-146;
    src/Vehicle.py:50-57) on the REFERENCE's basis"""
    fn, sd = RB.generate_Hilbert_BasisFunction(*p["hgp_args"])
    H, h0, R = jnp.array(p["H"]), jnp.array(p["h0"]), jnp.array(p["R"])
    if p["kind"] == "vehicle":
        basis = lambda state, input: fn(jnp.hstack(RVEH.f_alpha(state, input, l_f=1.16, l_r=1.47)))  # noqa: E731
    elif p["kind"] == "toy":
        basis = lambda state, input: fn(state)  # noqa: E731
    else:
        A, b = jnp.array(p["A"]), jnp.array(p["b"])
        basis = lambda state, input: fn(A @ jnp.hstack([state, input]) + b)  # noqa: E731
    lik = lambda obs, state, input: jnp.squeeze(jsp.stats.multivariate_normal.logpdf(obs, mean=H @ state + h0, cov=R))  # noqa: E731
    return basis, lik, sd


def read_sweep_variates(rd, T, N, n_x):
    Z = np.zeros((T, N, n_x))
    U = np.zeros((T, 2))
    Z[0] = rd.take("normal", (N, n_x))
    for t in range(1, T):
        U[t, 0] = rd.take("uniform", ())
        U[t, 1] = rd.take("uniform", ())
        Z[t] = rd.take_n("normal", N, (n_x,))
    U[0, 0] = rd.take("uniform", ())
    return Z, U


def read_param_variates(rd, n_x, M):
    return rd.take("chisquare", (n_x,)), rd.take("normal", (n_x, n_x)), rd.take("normal", (n_x, M))


def gen_group_a():
    for kind, g in CSMC_CASES.items():
        p = helpers.make_problem(kind, T=g["T"], N=g["N"], seed=g["seed"])
        basis, lik, sd = reference_callables(p)
        T, N, n_x, M = p["T"], p["N"], p["n_x"], p["M"]
        obs = p["obs"] if p["obs"].shape[1] > 1 else p["obs"]
        inputs = p["inputs"]
        csmc = RP.condSequentialMonteCarlo(N_samples=N, observations=jnp.array(obs), inputs=jnp.array(inputs),
                                           init_state_mean=jnp.array(p["m0"]), init_state_cov=jnp.array(p["P0"]),
                                           likelihood_fcn=lik, basis_fcn=basis)
        assert csmc.dim_basis == M
        put(f"csmc/{kind}/sd", np.asarray(sd))
        # whole sweep: condSequentialMonteCarlo.__call__
        jax.random.tape_reset()
        traj = csmc(jax.random.key(g["seed"]), jnp.array(p["ref"]), jnp.array(p["Theta"]), jnp.array(p["Sigma"]))
        rd = Reader()
        Z, U = read_sweep_variates(rd, T, N, n_x)
        rd.done()
        put(f"csmc/{kind}/Z", Z); put(f"csmc/{kind}/U", U); put(f"csmc/{kind}/traj", np.asarray(traj).reshape(T, n_x))
        # one step from a generic particle set: condSequentialMonteCarlo.step
        rng = np.random.default_rng(g["seed"] + 1000)
        state = p["ref"][4] + 0.05 * rng.normal(size=(N, n_x))
        logw = rng.normal(size=N)
        jax.random.tape_reset()
        nlw, nst, a = csmc.step(jax.random.key(g["seed"] + 1), jnp.array(5), jnp.array(logw), jnp.array(state),
                                jnp.array(p["Theta"]), jnp.array(p["Sigma"]), jnp.array(p["ref"][5]))
        rd = Reader()
        u_res, u_anc = rd.take("uniform", ()), rd.take("uniform", ())
        z = rd.take_n("normal", N, (n_x,))
        rd.done()
        put(f"step/{kind}/state", state); put(f"step/{kind}/logw", logw); put(f"step/{kind}/u", np.array([u_res, u_anc]))
        put(f"step/{kind}/z", z); put(f"step/{kind}/new_logw", nlw); put(f"step/{kind}/new_state", nst)
        put(f"step/{kind}/a", np.asarray(a, dtype=np.int64))
        # PGAS.sample_params and a short PGAS.__call__
        K = 3
        pg = RP.PGAS(N_samples=N, N_iterations=K, observations=jnp.array(obs), inputs=jnp.array(inputs),
                     init_state_mean=jnp.array(p["m0"]), init_state_cov=jnp.array(p["P0"]), likelihood_fcn=lik,
                     GP_prior=tuple(jnp.array(v) for v in p["prior"]), basis_fcn=basis)
        jax.random.tape_reset()
        A, S = pg.sample_params(jax.random.key(g["seed"] + 2), jnp.array(p["ref"]))
        rd = Reader()
        chi2, G, Nrm = read_param_variates(rd, n_x, M)
        rd.done()
        put(f"params/{kind}/chi2", chi2); put(f"params/{kind}/G", G); put(f"params/{kind}/Nrm", Nrm)
        put(f"params/{kind}/A", A); put(f"params/{kind}/S", S)
        jax.random.tape_reset()
        st_tr, ll = pg(jax.random.key(g["seed"] + 3), jnp.array(p["ref"]))
        rd = Reader()
        chi2s, Gs, Nrms, Zs, Us = [], [], [], [np.zeros((T, N, n_x))], [np.zeros((T, 2))]
        c, G_, Nr = read_param_variates(rd, n_x, M)
        chi2s.append(c); Gs.append(G_); Nrms.append(Nr)
        for k in range(1, K):
            Zk, Uk = read_sweep_variates(rd, T, N, n_x)
            Zs.append(Zk); Us.append(Uk)
            c, G_, Nr = read_param_variates(rd, n_x, M)
            chi2s.append(c); Gs.append(G_); Nrms.append(Nr)
        rd.done()
        put(f"pgas/{kind}/chi2", np.stack(chi2s)); put(f"pgas/{kind}/G", np.stack(Gs)); put(f"pgas/{kind}/Nrm", np.stack(Nrms))
        put(f"pgas/{kind}/Z", np.stack(Zs)); put(f"pgas/{kind}/U", np.stack(Us))
        put(f"pgas/{kind}/state_trace", np.asarray(st_tr).reshape(T, K, n_x)); put(f"pgas/{kind}/loglik", ll)
    put("csmc/kinds", np.array(list(CSMC_CASES)))


# ------------------------------------------------------------------------------------ group B: Algorithm1/2/3
MARG_CASES = {"smo": dict(T=10, N=16, M=8, seed=31), "vehicle": dict(T=9, N=16, M=6, seed=32)}


def reference_marginal(kind, mp):
    """the reference's classes on the reference's own model functions (src/SingleMassOscillator.py:17-48,101-107;
    src/Vehicle.py:17-128,146-153,215-220) and the data / noise levels of helpers_marginal.make_marg_problem"""
    kw = mp["prod_kwargs"]
    if kind == "smo":
        fn, sd = RB.generate_Hilbert_BasisFunction(mp["M"], np.array([[-7.5, 7.5], [-7.5, 7.5]]), 15.0 / mp["M"], 100.0)
        dt = 0.02
        ssm = RSSM(process_noise=np.diag([5e-8, 5e-9]), output_noise=np.array([[1e-3]]),
                   transition_model=lambda state, input, *iv: RSMO.f_x(state, input, iv[0], dt),
                   output_model=lambda state, input, *iv: RSMO.f_y(state))
        priors = [RBI.prior_mniw_2naturalPara(np.zeros((1, mp["M"])), np.diag(sd), np.eye(1), 3)]
        bases = [lambda state, input: fn(state)]
    else:
        fn, sd = RB.generate_Hilbert_BasisFunction(mp["M"], np.array([-30 / 180 * np.pi, 30 / 180 * np.pi]), 2 / 180 * np.pi, 50.0,
                                                   idx_start=2, idx_step=2)
        dt = 0.02
        ssm = RSSM(process_noise=np.diag([1e-8, 1e-8]), output_noise=np.diag([0.001 / 180 * np.pi, 1e-3]),
                   transition_model=lambda state, input, *iv: RVEH.f_x(state, input, iv[0], iv[1], dt),
                   output_model=lambda state, input, *iv: RVEH.f_y(state, input, iv[0], iv[1]))
        priors = [list(RBI.prior_mniw_2naturalPara(np.zeros((1, mp["M"])), np.diag(sd), np.eye(1), 0)) for _ in range(2)]
        bases = [lambda state, input: fn(RVEH.f_alpha(state, input)[0]), lambda state, input: fn(RVEH.f_alpha(state, input)[1])]
    common = dict(N_samples=kw["N_samples"], observations=kw["observations"], inputs=kw["inputs"], SSM=ssm,
                  init_state_mean=kw["init_state_mean"], init_state_cov=kw["init_state_cov"],
                  init_int_var_mean=[jnp.array(v) for v in kw["init_int_var_mean"]], init_int_var_cov=kw["init_int_var_cov"],
                  GP_prior=priors, basis_fcn=bases)
    return common


def read_marg_variates(rd, T, N, n_x, G, conditional):
    Z = np.zeros((T, N, n_x)); U = np.zeros((T, 2)); TS = np.zeros((G, T, N)); ZX = np.zeros((G, N))
    Z[0] = rd.take("normal", (N, n_x))
    for g in range(G):
        ZX[g] = rd.take("normal", (N, 1))[:, 0]
    for t in range(1, T):
        U[t, 0] = rd.take("uniform", ())
        if conditional:
            U[t, 1] = rd.take("uniform", ())
        Z[t] = rd.take_n("normal", N, (n_x,))
        for g in range(G):
            TS[g, t] = rd.take_n("t", N, (1,))[:, 0]
    if conditional:
        U[0, 0] = rd.take("uniform", ())
    return dict(Z=Z, ZXI0=ZX, U=U, TS=TS)


def gen_group_b():
    for kind, g in MARG_CASES.items():
        mp = HM.make_marg_problem(kind, T=g["T"], N=g["N"], M=g["M"], seed=g["seed"])
        T, N, G, n_x = mp["T"], mp["N"], mp["G"], mp["n_x"]
        common = reference_marginal(kind, mp)
        # Algorithm1.__call__
        a1 = RA1(forgetting_factor=mp["lam"], **common)
        jax.random.tape_reset()
        r = a1(jax.random.key(g["seed"]))
        rd = Reader()
        V = read_marg_variates(rd, T, N, n_x, G, conditional=False)
        rd.done()
        for k, v in V.items():
            put(f"alg1/{kind}/V_{k}", v)
        put(f"alg1/{kind}/state_trace", r[0])
        for i in range(G):
            put(f"alg1/{kind}/int_var_trace{i}", r[1][i])
            for j in range(4):
                put(f"alg1/{kind}/sst{i}_{j}", r[2][i][j])
                put(f"alg1/{kind}/final{i}_{j}", r[5][i][j])
        put(f"alg1/{kind}/weights_trace", r[3]); put(f"alg1/{kind}/ancestor_trace", np.asarray(r[4], dtype=np.int64))
        put(f"alg1/{kind}/obs_trace", r[6]); put(f"alg1/{kind}/loglik", r[7])
        # Algorithm2.__call__ (K - 1 conditional sweeps of Algorithm3 + reference statistics)
        K = 3
        a2 = RA2(N_iterations=K, **common)
        ref_x = np.asarray(r[0])[:, 0]
        ref_xi = [np.asarray(r[1][i])[:, 0, 0] for i in range(G)]
        jax.random.tape_reset()
        r2 = a2(jax.random.key(g["seed"] + 1), jnp.array(ref_x), [jnp.array(v) for v in ref_xi])
        rd = Reader()
        Vs = [read_marg_variates(rd, T, N, n_x, G, conditional=True) for _ in range(1, K)]
        rd.done()
        put(f"alg2/{kind}/ref_x", ref_x)
        for i in range(G):
            put(f"alg2/{kind}/ref_xi{i}", ref_xi[i])
        for name in ("Z", "ZXI0", "U", "TS"):
            put(f"alg2/{kind}/V_{name}", np.stack([np.zeros_like(Vs[0][name])] + [v[name] for v in Vs]))
        put(f"alg2/{kind}/state_trace", r2[0])
        for i in range(G):
            put(f"alg2/{kind}/int_var_trace{i}", r2[1][i])
            for j in range(4):
                put(f"alg2/{kind}/sst{i}_{j}", r2[3][i][j])
        put(f"alg2/{kind}/weights", r2[2]); put(f"alg2/{kind}/obs_trace", r2[4]); put(f"alg2/{kind}/loglik", r2[5])
    put("marg/kinds", np.array(list(MARG_CASES)))


# ------------------------------------------------------------------------------------ PGAS.sample_params at configuration scale
PARAMS_SCALE_CASES = {"smo256": dict(kind="smo", M=256, T=120, N=16, seed=77)}      # M of BASELINE.json configs[3]


def gen_params_scale():
    """the reference's own PGAS.sample_params (src/PGAS.py:288-343) at M = 256: statistics, the three M x M factorisations and the draw"""
    for name, g in PARAMS_SCALE_CASES.items():
        p = helpers.make_problem(g["kind"], T=g["T"], N=g["N"], M=g["M"], seed=g["seed"])
        basis, lik, sd = reference_callables(p)
        n_x, M = p["n_x"], p["M"]
        pg = RP.PGAS(N_samples=g["N"], N_iterations=2, observations=jnp.array(p["obs"]), inputs=jnp.array(p["inputs"]),
                     init_state_mean=jnp.array(p["m0"]), init_state_cov=jnp.array(p["P0"]), likelihood_fcn=lik,
                     GP_prior=tuple(jnp.array(v) for v in p["prior"]), basis_fcn=basis)
        jax.random.tape_reset()
        A, S = pg.sample_params(jax.random.key(g["seed"] + 2), jnp.array(p["ref"]))
        rd = Reader()
        chi2, G, Nrm = read_param_variates(rd, n_x, M)
        rd.done()
        put(f"params_scale/{name}/chi2", chi2); put(f"params_scale/{name}/G", G); put(f"params_scale/{name}/Nrm", Nrm)
        put(f"params_scale/{name}/A", A); put(f"params_scale/{name}/S", S)


if __name__ == "__main__":
    gen_sisr()
    gen_basis()
    gen_mniw()
    gen_group_a()
    gen_group_b()
    gen_params_scale()
    path = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(path, **OUT)
    print("wrote", path, len(OUT), "arrays,", os.path.getsize(path), "bytes")
