"""CPU tests: the C-ABI library builds for sm_100a, loads and exports every symbol the header
declares (no compute calls without a GPU); host-side logic (lattice search, model tracer, keys)."""
import os
import re

import numpy as np
import pytest

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "pgas_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(pgas_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 18
    L = helpers.pkg("_lib")
    for n in sorted(names):
        assert hasattr(built_lib, n), f"{n} declared in include/pgas_b200.h but not exported"
        assert n in L.EXPORTS, f"{n} has no ctypes prototype in _lib.EXPORTS"
    assert built_lib.pgas_version() >= 100


def test_struct_layout_matches_header():
    import ctypes as C
    L = helpers.pkg("_lib")
    # pgas_model_params: 6 int32, ptr, 2 int32, doubles..., checked through field offsets that the C compiler would produce
    assert L.ModelParams.freq.offset == 24
    assert L.ModelParams.center.offset == 40
    assert C.sizeof(L.Rng) == 64


def test_header_constants_match_the_bindings(tmp_path):
    """opcodes, limits, flags and the offsets of the plug-in fields as gcc sees them in include/pgas_b200.h vs the ctypes mirrors"""
    import ctypes as C
    import subprocess
    L = helpers.pkg("_lib")
    names = sorted(L.OPS)
    src = tmp_path / "consts.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pgas_b200.h"\nint main(void){\n'
                   + "".join(f'printf("%d\\n", (int)PGAS_OP_{n});\n' for n in names)
                   + 'printf("%d %d %d %d %d %d %d\\n", PGAS_MAX_PROG, PGAS_PROG_STACK, PGAS_MAX_NX, PGAS_MAX_NY, PGAS_MAX_NU, PGAS_MAX_D, PGAS_MAX_GP);\n'
                   + 'printf("%d %d %d %d %d %d\\n", PGAS_MAP_AFFINE, PGAS_MAP_VEHICLE_SLIP, PGAS_MAP_PROGRAM, PGAS_FLAG_ANCESTOR_GATHER, PGAS_FLAG_INPUT_PREV, PGAS_FLAG_VCHOL_TRANSPOSE);\n'
                   + 'printf("%zu %zu %zu %zu %d\\n", offsetof(pgas_model_params, prog_len), offsetof(pgas_model_params, flags), offsetof(pgas_model_params, lik_prog_len),'
                   + ' offsetof(pgas_model_params, lik_prog_const), PGAS_ABI_VERSION);\nreturn 0;}\n')
    exe = tmp_path / "consts"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    assert [int(v) for v in out[:len(names)]] == [L.OPS[n] for n in names]
    assert [int(v) for v in out[len(names)].split()] == [L.PGAS_MAX_PROG, L.PGAS_PROG_STACK, L.PGAS_MAX_NX, L.PGAS_MAX_NY, L.PGAS_MAX_NU,
                                                         L.PGAS_MAX_D, L.PGAS_MAX_GP]
    assert [int(v) for v in out[len(names) + 1].split()] == [L.MAP_AFFINE, L.MAP_VEHICLE_SLIP, L.MAP_PROGRAM, L.FLAG_ANCESTOR_GATHER,
                                                             L.FLAG_INPUT_PREV, L.FLAG_VCHOL_TRANSPOSE]
    offs = [int(v) for v in out[len(names) + 2].split()]
    assert offs[:4] == [L.ModelParams.prog_len.offset, L.ModelParams.flags.offset, L.ModelParams.lik_prog_len.offset,
                        L.ModelParams.lik_prog_const.offset]
    assert offs[4] == 204


def test_no_cpu_fallback_without_device(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    F = helpers.pkg("Filtering")
    with pytest.raises(helpers.pkg("_lib").PgasError):
        F.systematic_SISR(0.5, np.ones(4))
    p = helpers.make_problem("smo", T=5, N=8)
    cs = helpers.product_csmc(p)
    with pytest.raises(helpers.pkg("_lib").PgasError):
        cs(helpers.pkg("random").key(1), p["ref"], p["Theta"], p["Sigma"])


def test_product_does_not_import_the_oracle():
    pkg_dir = os.path.join(ROOT, helpers.PKG)
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
    for f in os.listdir(os.path.join(ROOT, "src")):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(ROOT, "src", f)).read(), f


@pytest.mark.parametrize("kind", ["smo", "emps", "toy", "vehicle"])
def test_lattice_search_matches_oracle(kind):
    p = helpers.make_problem(kind)
    hgp, sd = helpers.pkg("BasisFunctions").generate_Hilbert_BasisFunction(*p["hgp_args"])
    assert np.array_equal(hgp.freq, p["ohgp"].indices)
    assert np.allclose(sd, p["sd"], rtol=1e-14, atol=0)
    assert np.allclose(hgp.eigen_val, p["ohgp"].eig_val, rtol=1e-15)


def test_model_tracer_affine_families():
    MD, BF = helpers.pkg("models"), helpers.pkg("BasisFunctions")
    hgp3, _ = BF.generate_Hilbert_BasisFunction(27, np.array([[-1.0, 1.0]] * 3), 0.1, 20)
    e = MD.trace_basis(lambda s, u: hgp3(np.hstack([s, u]) / np.array([0.4, 0.4, 160.0])), 2, 1)   # src/EMPS.py:110-113
    assert np.allclose(e.Az, np.diag([2.5, 2.5, 1 / 160.0])) and np.allclose(e.bz, 0)
    hgp1, _ = BF.generate_Hilbert_BasisFunction(9, np.array([-0.2, 0.2]), 0.4 / 9, 20)
    e = MD.trace_basis(lambda s, u: hgp1(s[1]), 2, 1)                                               # src/EMPS.py:90-91
    assert np.allclose(e.Az, [[0, 1, 0]])
    e = MD.trace_basis(lambda s, u: hgp1(2.0 * s[0] - 0.5 + u[0] * 0.1), 2, 1)
    assert np.allclose(e.Az, [[2, 0, 0.1]]) and np.allclose(e.bz, [-0.5])
    e = MD.trace_basis(lambda s, u: hgp1(np.sin(s[0])), 2, 1)        # not affine: expression program (model plug-in), no longer an error
    assert isinstance(e, MD.ProgramBasis) and e.ops == [MD._lib.OPS["PUSH_X"], MD._lib.OPS["SIN"]]
    with pytest.raises(TypeError):
        MD.trace_basis(lambda s, u: hgp1(np.floor(s[0])), 2, 1)      # outside the instruction set
    with pytest.raises(TypeError):
        MD.trace_basis(lambda s, u: s[0], 2, 1)
    lik = MD.resolve_likelihood(MD.gaussian_likelihood(lambda x: x[0], np.eye(1) * 1e-4), 2)       # src/EMPS.py:250-252
    assert np.allclose(lik.H, [[1, 0]]) and np.allclose(lik.h0, 0) and np.allclose(lik.R, 1e-4)
    with pytest.raises(TypeError):
        MD.resolve_likelihood(lambda o, s, i: 0.0, 2)


def test_keys_split_and_uniform():
    R = helpers.pkg("random")
    k = R.key(12345678)
    a, b = R.split(k)
    assert a.seed != b.seed and a.seed != k.seed
    assert [x.seed for x in R.split(k)] == [a.seed, b.seed]          # deterministic
    assert len({x.seed for x in R.split(k, 50)}) == 50
    u = R.uniform(a)
    assert 0.0 <= u < 1.0 and R.uniform(a) == u
    us = R.uniform(b, (1000,))
    assert abs(us.mean() - 0.5) < 0.05
    from oracle import philox as OPH
    assert [int(v) for v in R.philox4x32(0, 0, 0, 0, 0, 0)] == [int(v) for v in OPH.philox(0, 0, 0, 0, 0)]


def test_natural_parameter_host_conversions_match_oracle():
    from oracle import mniw as OM
    BI = helpers.pkg("BayesianInferrence")
    rng = np.random.default_rng(0)
    M, n = 9, 2
    mean = rng.normal(size=(n, M)); B = rng.normal(size=(M, M)); V = B @ B.T + M * np.eye(M)
    Psi = np.array([[2.0, 0.1], [0.1, 1.0]])
    for a, b in zip(BI.prior_mniw_2naturalPara(mean, V, Psi, 5), OM.prior_mniw_2naturalPara(mean, V, Psi, 5)):
        assert np.allclose(a, b, rtol=1e-12, atol=1e-13)
    eta = OM.prior_mniw_2naturalPara(mean, V, Psi, 5)
    for a, b in zip(BI.prior_mniw_2naturalPara_inv(*eta), OM.prior_mniw_2naturalPara_inv(*eta)):
        assert np.allclose(a, b, rtol=1e-11, atol=1e-12)
    assert np.allclose(BI.prior_mniw_mean(eta[0], eta[1]), OM.prior_mniw_mean(eta[0], eta[1]), rtol=1e-11)
    phi = rng.normal(size=M)
    for a, b in zip(BI.prior_mniw_Predictive(mean, V, Psi, 5, phi), OM.prior_mniw_Predictive(mean, V, Psi, 5, phi)):
        assert np.allclose(a, b, rtol=1e-12)
    T = (eta[0], eta[1], eta[2] + np.eye(n), 7.0)
    assert abs(BI.prior_mniw_log_base_measure(*T) - OM.prior_mniw_log_base_measure(*T)) < 1e-9


def test_expression_tracer_compiles_a_non_affine_basis_fcn():
    """model plug-in: a basis_fcn that is not affine is traced into a postfix program; the host mirror of
    the device interpreter reproduces the callable; affine callables keep the compiled-in family"""
    import helpers
    MD = helpers.pkg("models")
    L = helpers.pkg("_lib")

    class FakeBasis:                                   # stands in for HilbertBasis.__call__ without a device
        D, M = 2, 7

        def __call__(self, x):
            return MD.ProgramBasis(self, x) if isinstance(x, MD.Sym) else MD.BasisExpr(self, x.A, x.b)
    hgp = FakeBasis()
    b = MD.trace_basis(lambda s, u: hgp(MD.hstack(helpers.plugin_map(s, u))), 2, 1)
    assert isinstance(b, MD.ProgramBasis) and b.map_kind == L.MAP_PROGRAM and len(b.ops) <= L.PGAS_MAX_PROG
    rng = np.random.default_rng(0)
    for _ in range(20):
        x, u = rng.normal(size=2) * 3, rng.normal(size=1)
        assert np.allclose(MD.run_program(b.ops, b.consts, x, u), np.array(helpers.plugin_map(x, u)), rtol=1e-14, atol=1e-14)
    a = MD.trace_basis(lambda s, u: hgp(MD.hstack([s[0] * 2.0 + u[0], s[1] - 1.0])), 2, 1)
    assert isinstance(a, MD.BasisExpr) and a.map_kind == L.MAP_AFFINE
    with pytest.raises(TypeError):
        MD.trace_basis(lambda s, u: hgp(np.floor(s)), 2, 1)       # not in the instruction set: raises, no host fallback


def test_reference_style_likelihood_lambda_is_traced():
    """the reference's likelihood_fcn form (src/EMPS.py:250-252, src/Toy_Example.py:142-144) with this package's stats module in place
    of jax.scipy.stats reaches the same Gaussian observation model as models.gaussian_likelihood; on numbers it is SciPy's density"""
    import helpers
    import scipy.stats
    MD, ST = helpers.pkg("models"), helpers.pkg("stats")
    R = np.array([[1e-4]])
    f_y = lambda x: x[0]                                                                          # noqa: E731  (src/EMPS.py:215-216)
    lam = lambda obs, state, input: np.squeeze(ST.multivariate_normal.logpdf(obs, mean=f_y(state), cov=R))   # noqa: E731
    a = MD.resolve_likelihood(lam, 2, 1)
    b = MD.resolve_likelihood(MD.gaussian_likelihood(f_y, R), 2, 1)
    assert np.array_equal(a.H, b.H) and np.array_equal(a.h0, b.h0) and np.array_equal(a.R, b.R) and a.H.tolist() == [[1.0, 0.0]]
    R2 = np.array([[2.0, 0.3], [0.3, 1.0]])
    lam2 = lambda obs, state, input: ST.multivariate_normal.logpdf(obs, mean=np.hstack([state[1] * 2.0 + 1.0, state[0] - state[2]]), cov=R2)  # noqa: E731
    c = MD.resolve_likelihood(lam2, 3, 0)
    assert np.allclose(c.H, [[0, 2, 0], [1, 0, -1]]) and np.allclose(c.h0, [1, 0]) and np.array_equal(c.R, R2)
    x, y = np.array([0.3, -1.0, 2.0]), np.array([0.1, 0.2])
    assert np.isclose(lam2(y, x, None), scipy.stats.multivariate_normal.logpdf(y, mean=[-1.0, -1.7], cov=R2), rtol=1e-13)
    for bad in (lambda obs, state, input: ST.multivariate_normal.logpdf(obs, mean=np.sin(state[:1]), cov=R),          # not affine
                lambda obs, state, input: ST.multivariate_normal.logpdf(obs, mean=state[:1] + input, cov=R),         # depends on the input
                lambda obs, state, input: -0.5 * (obs - state[0]) ** 2):                                              # not a traced density
        with pytest.raises(TypeError):
            MD.resolve_likelihood(bad, 2, 1)


def test_likelihood_outside_the_gaussian_family_becomes_a_program():
    """model plug-in (src/PGAS.py:24-43 hands ANY likelihood_fcn to the sampler): a Gaussian with a non-affine
    mean, written with the stats stand-in, and a hand-written Student-t log-density both compile to expression programs over
    (state, input, observation); the host mirror of the device interpreter reproduces the callables"""
    import helpers
    import scipy.stats
    MD, ST, L = helpers.pkg("models"), helpers.pkg("stats"), helpers.pkg("_lib")
    R2 = np.array([[2.0, 0.3], [0.3, 1.0]])
    lam = lambda obs, state, input: np.squeeze(ST.multivariate_normal.logpdf(                     # noqa: E731
        obs, mean=np.hstack([np.sin(state[0]) + input[0], state[1] * state[0]]), cov=R2))
    a = MD.resolve_likelihood(lam, 2, 1, 2)
    b = MD.resolve_likelihood(helpers.plugin_loglik, 2, 1, 1)
    assert isinstance(a, MD.ProgramLikelihood) and isinstance(b, MD.ProgramLikelihood)
    assert max(len(a.ops), len(b.ops)) <= L.PGAS_MAX_PROG and any((i & 0xFF) == L.OPS["PUSH_Y"] for i in b.ops)
    rng = np.random.default_rng(3)
    for _ in range(10):
        x, u, y = rng.normal(size=2), rng.normal(size=1), rng.normal(size=2)
        want = scipy.stats.multivariate_normal.logpdf(y, mean=[np.sin(x[0]) + u[0], x[1] * x[0]], cov=R2)
        assert np.isclose(MD.run_program(a.ops, a.consts, x, u, y)[0], want, rtol=1e-13)
        assert np.isclose(MD.run_program(b.ops, b.consts, x, u, y[:1])[0], helpers.plugin_loglik(y[:1], x, u), rtol=1e-14)
    # the torch evaluation used for the final log-likelihood table of PGAS.__call__ (src/PGAS.py:383-392)
    import torch
    X, Y, U = torch.randn(5, 3, 2, dtype=torch.float64), torch.randn(5, 1, 1, dtype=torch.float64), torch.randn(5, 1, 1, dtype=torch.float64)
    out = b.logpdf_torch(Y, X, U)
    assert out.shape == (5, 3) and np.isclose(float(out[2, 1]), helpers.plugin_loglik(Y[2, 0].numpy(), X[2, 1].numpy(), U[2, 0].numpy()), rtol=1e-14)
    # affine Gaussian stays in the compiled-in family even when the observation count is known
    g = MD.resolve_likelihood(lambda obs, state, input: ST.multivariate_normal.logpdf(obs, mean=state[:1], cov=np.eye(1)), 2, 1, 1)
    assert isinstance(g, MD.GaussianLikelihood)
    with pytest.raises(TypeError):
        MD.resolve_likelihood(lambda obs, state, input: np.floor(state[0]) - obs[0], 2, 1, 1)
