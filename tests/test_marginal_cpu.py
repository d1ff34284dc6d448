"""CPU tests for the marginalised path (SURVEY.md 8a group B): oracle pins (analytic identities, SciPy
cross-checks, golden fixtures), the model tracer against direct evaluation of the shipped models, the
C struct layouts against the header (compiled with gcc), and the no-CPU-fallback guarantee."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers
import helpers_marginal as HM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ------------------------------------------------------------------------------- oracle pins
def test_log_base_measure_against_scipy_densities():
    """exp(log_base_measure) normalises the MNIW density: for n = 1 compare with the closed form built from
    scipy's multivariate-t / inverse-gamma pieces: log Z = log |T1|^{1/2} ... (src/BayesianInferrence.py:111-124)"""
    from scipy.special import gammaln
    from oracle import mniw as OM
    rng = np.random.default_rng(0)
    M = 6
    B = rng.normal(size=(30, M)); y = rng.normal(size=30)
    T1 = B.T @ B + np.eye(M); T0 = (B.T @ y)[:, None]; T2 = np.array([[y @ y + 2.0]]); nu = 31.0
    psi = T2[0, 0] - T0[:, 0] @ np.linalg.solve(T1, T0[:, 0])
    expect = (-0.5 * M * np.log(2 * np.pi) + 0.5 * np.linalg.slogdet(T1)[1] - 0.5 * nu * np.log(2) - gammaln(nu / 2)
              + 0.5 * nu * np.log(psi))
    assert abs(OM.prior_mniw_log_base_measure(T0, T1, T2, nu) - expect) < 1e-10 * abs(expect)


def test_predictive_is_student_t_of_bayesian_linear_regression():
    """prior_mniw_Predictive (src/BayesianInferrence.py:64-89): for n = 1 the predictive of y* = a.phi + e is Student-t
    with mean phi.mean, scale^2 = Psi/df' (1 + phi V phi), checked against the textbook conjugate update."""
    from oracle import mniw as OM
    rng = np.random.default_rng(1)
    M = 5
    Phi = rng.normal(size=(40, M)); y = Phi @ rng.normal(size=M) + 0.1 * rng.normal(size=40)
    V0 = np.diag(rng.uniform(0.5, 2.0, M)); nu0 = 3.0
    eta = OM.prior_mniw_2naturalPara(np.zeros((1, M)), V0, np.eye(1), nu0)
    eta = (eta[0] + (Phi.T @ y)[:, None], eta[1] + Phi.T @ Phi, eta[2] + y @ y, eta[3] + 40)
    mean, V, Psi, df = OM.prior_mniw_2naturalPara_inv(*eta)
    Vn = np.linalg.inv(np.linalg.inv(V0) + Phi.T @ Phi)
    mn = Vn @ Phi.T @ y
    assert np.allclose(mean[0], mn, rtol=1e-10) and np.allclose(V, Vn, rtol=1e-10)
    assert np.isclose(Psi[0, 0], 1.0 + y @ y - mn @ np.linalg.solve(Vn, mn), rtol=1e-10)
    phi = rng.normal(size=M)
    pm, pc, pr, pdf = OM.prior_mniw_Predictive(mean, V, Psi, df, phi)
    assert np.isclose(pm, phi @ mn) and np.isclose(pc[0, 0], 1 + phi @ Vn @ phi) and pdf == df and np.isclose(pr[0, 0], Psi[0, 0] / df)


@pytest.mark.parametrize("kind", ["smo", "vehicle"])
def test_conditional_sweep_keeps_the_reference_path(kind):
    """Algorithm3 pins particle N-1 to the reference (src/Algorithm3.py:132, :147-150, :221-232) and its statistics
    bookkeeping is consistent: after the last step the remaining reference statistics are zero."""
    from oracle import marginal as OMg
    prob = HM.make_marg_problem(kind, T=10, N=12, M=6, seed=4)
    V1 = HM.make_variates(prob, 1.0, seed=5)
    f = OMg.alg1_run(prob["oracle"], 12, 1.0, HM.oracle_variates(V1))
    ref_x = f["state_trace"][:, 3]
    ref_xi = [f["int_var_trace"][g][:, 3, 0] for g in range(prob["G"])]
    rs = OMg.reference_stats(prob["oracle"], ref_x, ref_xi)
    r = OMg.alg3_run(prob["oracle"], 12, ref_x, ref_xi, rs, HM.oracle_variates(HM.make_variates(prob, 1.0, seed=6)))
    assert np.array_equal(r["state_trace"][:, -1], ref_x)
    for g in range(prob["G"]):
        assert np.array_equal(r["int_var_trace"][g][:, -1, 0], ref_xi[g])
        for j in range(4):
            scale = np.max(np.abs(np.asarray(rs[g][j], dtype=float)))
            assert np.max(np.abs(np.asarray(r["ref_st_end"][g][j], dtype=float))) < 1e-12 * max(scale, 1.0)
    # weights of an Algorithm1 run are a probability vector at every step
    assert np.allclose(f["weights_trace"].sum(axis=1), 1.0)


def test_golden_fixtures_marginal():
    sys.path.insert(0, GOLD)
    import make_golden_marginal as G
    with open(os.path.join(GOLD, "oracle_marginal_golden.json")) as fh:
        gold = json.load(fh)
    for name, g in gold.items():
        out = G.run_case(g)
        for k, v in out.items():
            a, b = np.asarray(v, dtype=float), np.asarray(g[k], dtype=float)
            if k.endswith("anc_last") or k == "alg3_idx":
                assert np.array_equal(a, b), (name, k)
            else:
                assert np.allclose(a, b, rtol=1e-9, atol=1e-12), (name, k)


# ------------------------------------------------------------------------------- tracer
def _check_tables(ssm, bases, inputs, n_xi, rng, atol=1e-13):
    tr = helpers.pkg("tracing")
    G = len(n_xi)
    trans, outp, link = ssm.tables(inputs, 2, n_xi)
    state, _ = tr.variables(2, n_xi)
    for t in range(inputs.shape[0]):
        x = rng.normal(size=2) * 0.3
        xi = [rng.normal(size=1) * 0.5 for _ in range(G)]
        v = np.concatenate([x] + xi + [[1.0]])
        assert np.allclose(trans[t] @ v, ssm.transition_model(x, inputs[t], *xi), rtol=1e-12, atol=atol)
        y = outp[t] @ v
        if link == "tanh":
            y = np.tanh(y)
        assert np.allclose(y, np.atleast_1d(ssm.output_model(x, inputs[t], *xi)), rtol=1e-12, atol=atol)
        for b in bases:
            c = b(state, inputs[t])
            z = c.z.A[:, :2] @ x + c.z.b
            if c.z.link == "atan":
                z = np.arctan(z)
            z = c.z.p * z + c.z.q
            hg = c.hgp
            direct = np.prod(np.sqrt(1.0 / hg.half_width) * np.sin(np.sqrt(hg.eigen_val) * ((z - hg.center) + hg.half_width)), axis=1)
            # the same basis through the oracle's closed form of the user-level call
            yield t, x, z, direct


def test_tracer_reproduces_the_shipped_models():
    rng = np.random.default_rng(0)
    from oracle import basis as OB
    S = helpers.pkg("SingleMassOscillator")
    for t, x, z, phi in _check_tables(S.SMO_SSM, [lambda s, u: S.basis_fcn(s)], S.F_ext[::150], [1], rng):
        o = OB.HilbertBasis(S.basis_fcn.freq.astype(np.int64), np.array([[-7.5, 7.5], [-7.5, 7.5]]))
        assert np.allclose(z, x) and np.allclose(phi, o(x), rtol=1e-12, atol=1e-15)
    V = helpers.pkg("Vehicle")
    dom = np.array([-30 / 180 * np.pi, 30 / 180 * np.pi])
    o = OB.HilbertBasis(V.basis_fcn.freq.astype(np.int64), dom)
    k = 0
    for t, x, z, phi in _check_tables(V.Vehicle_SSM, [V.basis_fcn_f, V.basis_fcn_r], V.ctrl_input[100::300], [1, 1], rng):
        af, ar = V.f_alpha(x, V.ctrl_input[100::300][t])
        assert np.isclose(z[0], af if k % 2 == 0 else ar, rtol=1e-12, atol=1e-15)
        assert np.allclose(phi, o(z), rtol=1e-12, atol=1e-15)
        k += 1
    E = helpers.pkg("EMPS")
    for t, x, z, phi in _check_tables(E.EMPS_SSM, [E.basis_fcn_f], E.ctrl_input[::500], [1], rng, atol=1e-12):
        assert np.isclose(z[0], x[1])


def test_tracer_rejects_models_outside_the_families():
    tr, SSMm, BF = helpers.pkg("tracing"), helpers.pkg("StateSpaceModel"), helpers.pkg("BasisFunctions")
    s, xi = tr.variables(2, [1])
    with pytest.raises(TypeError):
        _ = s[0] * s[1]
    with pytest.raises(TypeError):
        _ = np.sin(s[0])
    with pytest.raises(TypeError):
        _ = np.arctan(np.tanh(s))
    bad = SSMm.StateSpaceModel(np.eye(2), np.eye(1), lambda x, u, *v: np.hstack([x[1] * x[0], v[0]]), lambda x, u, *v: x[0])
    with pytest.raises(TypeError):
        bad.tables(np.zeros(3), 2, [1])
    hgp, _ = BF.generate_Hilbert_BasisFunction(5, np.array([-1.0, 1.0]), 0.3, 1.0)
    c = hgp(2.0 - np.arctan(3.0 * s[1] + 1.0))
    assert c.z.link == "atan" and np.allclose(c.z.A, [[0, 3, 0]]) and np.allclose(c.z.b, [1]) and np.allclose(c.z.p, [-1]) and np.allclose(c.z.q, [2])


def test_plugin_model_compiles_to_expression_programs():
    """model plug-in (src/StateSpaceModel.py:19-30 'any callable'): a transition / output model / GP-input map
    outside the coefficient-table families is rejected by the affine tracer and traced into postfix programs; the host mirror of
    the device interpreter reproduces the callables and the oracle's batched twins"""
    MD, SSMm, L = helpers.pkg("models"), helpers.pkg("StateSpaceModel"), helpers.pkg("_lib")
    prob = HM.make_marg_problem("plugin", T=12, N=8, M=6)
    ssm, inputs = prob["prod_kwargs"]["SSM"], prob["prod_kwargs"]["inputs"]
    with pytest.raises(TypeError):
        ssm.tables(inputs, 2, [1])
    tprog, oprog, n_y = ssm.programs(inputs, 2, [1])
    assert n_y == 1 and max(len(tprog[0]), len(oprog[0])) <= L.PGAS_MAX_PROG
    s_sym, _, u_sym = SSMm.program_variables(2, 0, inputs)

    class FakeBasis:                                   # stands in for HilbertBasis.__call__ without a device
        D, M = 1, 6

        def __call__(self, x):
            return MD.ProgramBasis(self, x)
    zprog = FakeBasis()(HM.plugin_z(s_sym))
    rng = np.random.default_rng(1)
    om = prob["oracle"]
    for t in range(12):
        x, xi = rng.normal(size=2), rng.normal(size=1)
        v, u = np.concatenate([x, xi]), np.atleast_1d(inputs[t])
        want_f = ssm.transition_model(x, inputs[t], xi)
        assert np.allclose(MD.run_program(*tprog, v, u), want_f, rtol=1e-14, atol=1e-15)
        assert np.allclose(om.ssm.transition(x[None], inputs[t], xi[None]), want_f[None], rtol=1e-14, atol=1e-15)
        want_g = ssm.output_model(x, inputs[t], xi)
        assert np.allclose(MD.run_program(*oprog, v, u), want_g, rtol=1e-14, atol=1e-15)
        assert np.allclose(om.ssm.output(x[None], inputs[t], xi[None]), want_g, rtol=1e-14, atol=1e-15)
        assert np.allclose(MD.run_program(zprog.ops, zprog.consts, x, u), HM.plugin_z(x), rtol=1e-14, atol=1e-15)
    # outside the instruction set: raises, no host fallback
    bad = SSMm.StateSpaceModel(np.eye(2), np.eye(1), lambda s, u, *xi: np.hstack([np.floor(s[0]), xi[0]]), lambda s, u, *xi: s[0])
    with pytest.raises(TypeError):
        bad.programs(inputs, 2, [1])
    # the shipped models stay on the coefficient tables
    S = helpers.pkg("SingleMassOscillator")
    S.SMO_SSM.tables(S.F_ext[:3], 2, [1])


# ------------------------------------------------------------------------------- ABI
def test_marginal_struct_layouts_match_the_header(tmp_path):
    """compile a tiny C program against include/pgas_b200.h and compare sizeof / offsetof with the ctypes mirrors"""
    import ctypes as C
    L = helpers.pkg("_lib")
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pgas_b200.h"\nint main(void){\n'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(pgas_marg_program), offsetof(pgas_marg_gp, prog),'
                   'offsetof(pgas_marg_params, inputs), offsetof(pgas_marg_params, outp_prog),'
                   'sizeof(pgas_marg_gp), sizeof(pgas_marg_params), sizeof(pgas_marg_rng),'
                   'offsetof(pgas_marg_gp, gp_in), offsetof(pgas_marg_gp, xi_var), offsetof(pgas_marg_params, trans),'
                   'offsetof(pgas_marg_params, Q), offsetof(pgas_marg_params, P0), offsetof(pgas_marg_rng, TS), sizeof(pgas_model_params));'
                   'return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(L.MargProgram), L.MargGP.prog.offset, L.MargParams.inputs.offset, L.MargParams.outp_prog.offset,
            C.sizeof(L.MargGP), C.sizeof(L.MargParams), C.sizeof(L.MargRng), L.MargGP.gp_in.offset, L.MargGP.xi_var.offset,
            L.MargParams.trans.offset, L.MargParams.Q.offset, L.MargParams.P0.offset, L.MargRng.TS.offset, C.sizeof(L.ModelParams)]
    assert got == want


def test_marginal_no_cpu_fallback_without_device(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    prob = HM.make_marg_problem("smo", T=6, N=8, M=6)
    A1 = helpers.pkg("Algorithm1").Algorithm1(forgetting_factor=0.999, **prob["prod_kwargs"])
    with pytest.raises(helpers.pkg("_lib").PgasError):
        A1(helpers.pkg("random").key(1))
    A2 = helpers.pkg("Algorithm2").Algorithm2(N_iterations=3, **prob["prod_kwargs"])
    with pytest.raises(helpers.pkg("_lib").PgasError):
        A2(helpers.pkg("random").key(1), np.zeros((6, 2)), [np.zeros(6)])
