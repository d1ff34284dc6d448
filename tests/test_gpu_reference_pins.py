"""GPU parity against outputs of the REFERENCE'S OWN SOURCE (tests/golden/reference_golden.npz, produced by
tests/golden/make_reference_golden.py from the unmodified reference modules on a NumPy-backed jax stand-in): the
CUDA path is fed the variates the reference run consumed and must reproduce the reference's outputs — indices
exactly, float64 results within 1e-9 relative (BASELINE.json).  The CPU oracle is not involved here."""
import numpy as np
import pytest

import helpers
import helpers_marginal as HM
from test_reference_pins import CSMC_CASES, GOLD, MARG_CASES

pytestmark = pytest.mark.gpu
REL = helpers.REL_TOL


def _np(t):
    return t.detach().cpu().numpy()


def _dev(a):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("kind", list(CSMC_CASES))
def test_sweep_reproduces_reference_source(kind):
    c = CSMC_CASES[kind]
    p = helpers.make_problem(kind, T=c["T"], N=c["N"], seed=c["seed"])
    cs = helpers.product_csmc(p)
    g = cs.sweep(_dev(p["ref"]), _dev(p["Theta"]), _dev(p["Sigma"]),
                 variates=dict(Z=_dev(GOLD[f"csmc/{kind}/Z"][None]), U=_dev(GOLD[f"csmc/{kind}/U"][None])))
    assert helpers.rel_err(_np(g["traj"][0]), GOLD[f"csmc/{kind}/traj"]) < REL


@pytest.mark.parametrize("kind", list(CSMC_CASES))
def test_step_reproduces_reference_source(kind):
    c = CSMC_CASES[kind]
    p = helpers.make_problem(kind, T=c["T"], N=c["N"], seed=c["seed"])
    cs = helpers.product_csmc(p)
    s = lambda k: GOLD[f"step/{kind}/{k}"]  # noqa: E731
    lw, xs, a = (_np(v) for v in cs.step(5, s("logw"), s("state"), p["Theta"], p["Sigma"], p["ref"][5], s("u"), s("z")))
    np.testing.assert_array_equal(a, s("a"))
    assert helpers.rel_err(xs, s("new_state")) < REL
    assert helpers.rel_err(lw, s("new_logw")) < REL


@pytest.mark.parametrize("kind", list(CSMC_CASES))
def test_sample_params_reproduces_reference_source(kind):
    c = CSMC_CASES[kind]
    p = helpers.make_problem(kind, T=c["T"], N=c["N"], seed=c["seed"])
    pg = helpers.product_pgas(p, K=2)
    s = lambda k: GOLD[f"params/{kind}/{k}"]  # noqa: E731
    A, S = pg.sample_params(None, _dev(p["ref"][None]), variates=dict(chi2=_dev(s("chi2")[None]), G=_dev(s("G")[None]), Nrm=_dev(s("Nrm")[None])))
    assert helpers.rel_err(_np(A[0]), s("A")) < REL
    assert helpers.rel_err(_np(S[0]), s("S")) < REL


@pytest.mark.parametrize("name", ["smo256"])
def test_sample_params_reproduces_reference_source_at_config_scale(name):
    """statistics kernels (sine table + DMMA SYRK) and the blocked-Cholesky draw against PGAS.sample_params of the reference's own
    source at M = 256 (src/PGAS.py:288-343)"""
    c = dict(kind="smo", M=256, T=120, N=16, seed=77)          # keep equal to tests/golden/make_reference_golden.py
    p = helpers.make_problem(c["kind"], T=c["T"], N=c["N"], M=c["M"], seed=c["seed"])
    pg = helpers.product_pgas(p, K=2)
    s = lambda k: GOLD[f"params_scale/{name}/{k}"]  # noqa: E731
    A, S = pg.sample_params(None, _dev(p["ref"][None]), variates=dict(chi2=_dev(s("chi2")[None]), G=_dev(s("G")[None]), Nrm=_dev(s("Nrm")[None])))
    assert helpers.rel_err(_np(A[0]), s("A")) < REL
    assert helpers.rel_err(_np(S[0]), s("S")) < REL


def _marg_dev(prefix):
    V = dict(Z=GOLD[prefix + "V_Z"], ZXI0=GOLD[prefix + "V_ZXI0"], U=GOLD[prefix + "V_U"], TS=GOLD[prefix + "V_TS"])
    return HM.device_variates(V)


@pytest.mark.parametrize("kind", list(MARG_CASES))
def test_algorithm1_reproduces_reference_source(kind):
    c = MARG_CASES[kind]
    T, N, M = c["T"], c["N"], c["M"]
    mp = HM.make_marg_problem(kind, T=T, N=N, M=M, seed=c["seed"])
    A1 = helpers.pkg("Algorithm1").Algorithm1(forgetting_factor=mp["lam"], **mp["prod_kwargs"])
    r = A1.filter(variates=_marg_dev(f"alg1/{kind}/"))
    s = lambda k: GOLD[f"alg1/{kind}/{k}"]  # noqa: E731
    assert int(r["status"][0]) == 0
    np.testing.assert_array_equal(_np(r["anc_trace"][0]), s("ancestor_trace"))
    assert HM.rel_err(_np(r["state_trace"][0]), s("state_trace")) < REL
    for g in range(mp["G"]):
        assert HM.rel_err(_np(r["xi_trace"][0, g]), s(f"int_var_trace{g}")[..., 0]) < REL
        for j, sh in enumerate([(T, M), (T, M, M), (T,), (T,)]):
            assert HM.rel_err(_np(r["sst_trace"][4 * g + j][0]).reshape(sh), s(f"sst{g}_{j}").reshape(sh)) < REL
        for j, sh in enumerate([(N, M), (N, M, M), (N,), (N,)]):
            assert HM.rel_err(_np(r["final_stats"][4 * g + j][0]).reshape(sh), s(f"final{g}_{j}").reshape(sh)) < REL


@pytest.mark.parametrize("kind", list(MARG_CASES))
def test_algorithm2_reproduces_reference_source(kind):
    import torch
    c = MARG_CASES[kind]
    mp = HM.make_marg_problem(kind, T=c["T"], N=c["N"], M=c["M"], seed=c["seed"])
    s = lambda k: GOLD[f"alg2/{kind}/{k}"]  # noqa: E731
    K = 3
    A2 = helpers.pkg("Algorithm2").Algorithm2(N_iterations=K, **mp["prod_kwargs"])
    f64 = dict(dtype=torch.float64, device="cuda")
    init_xi = np.stack([s(f"ref_xi{g}") for g in range(mp["G"])])
    r = A2.run(torch.as_tensor(s("ref_x")[None], **f64), torch.as_tensor(init_xi[None], **f64), variates=_marg_dev(f"alg2/{kind}/"))
    assert int(r["status"][0]) == 0
    assert HM.rel_err(_np(r["x_trace"][0]).transpose(1, 0, 2), s("state_trace")) < REL
    for g in range(mp["G"]):
        assert HM.rel_err(_np(r["xi_trace"][0, g]).T, s(f"int_var_trace{g}")[..., 0]) < REL
        for j in range(4):
            got = np.stack([_np(r["sst"][4 * g + j][0, k]).reshape(-1) for k in range(K)])
            assert HM.rel_err(got, s(f"sst{g}_{j}").reshape(K, -1)) < REL
