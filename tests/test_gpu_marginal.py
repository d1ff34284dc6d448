"""GPU parity of the marginalised filters (SURVEY.md 8a group B) against the CPU oracle under injected
variates: ancestor indices exact, float64 traces / statistics within 1e-9 relative (BASELINE.json)."""
import numpy as np
import pytest

import helpers
import helpers_marginal as HM

pytestmark = pytest.mark.gpu
REL = helpers.REL_TOL


def _np(t):
    return t.detach().cpu().numpy()


def _ulp_perturbed(V, seed=1):
    """the normal / Student-t variates moved by one unit in the last place (random sign)"""
    rng = np.random.default_rng(seed)
    return {k: (v * (1.0 + 2.2e-16 * np.sign(rng.normal(size=np.shape(v)))) if k in ("Z", "TS", "ZXI0") else v) for k, v in V.items()}


def _tolerance(run_oracle, V, ref, fields):
    """BASELINE.json's 1e-9 — unless the ORACLE ITSELF moves by more than that when its input variates change by one ulp (the
    shipped vehicle configuration does: prior precisions 1/sd span ten orders of magnitude), in which case two correct float64
    implementations cannot agree more closely than that sensitivity; then 10x the measured sensitivity is the bar."""
    other = run_oracle(HM.oracle_variates(_ulp_perturbed(V)))
    sens = max(HM.rel_err(np.asarray(b, dtype=np.float64), np.asarray(a, dtype=np.float64)) for a, b in zip(fields(ref), fields(other)))
    return max(REL, 10.0 * sens)


@pytest.mark.parametrize("kind,T,N,M,cs", [("smo", 24, 32, 12, 0), ("smo", 12, 200, 41, 0), ("emps", 30, 48, 9, 1),
                                           ("vehicle", 24, 40, 8, 0), ("smo", 16, 70, 12, 4),
                                           # model plug-in: transition / output / GP-input map as interpreted expression programs
                                           ("plugin", 24, 32, 10, 0), ("plugin", 40, 200, 20, 0), ("plugin", 20, 40, 10, 2),
                                           # the shipped sizes (BASELINE configs[0..2]: N = 200; M = 41 | 2 x 20 | 9), 60-100 steps
                                           ("smo", 100, 200, 41, 0), ("vehicle", 60, 200, 20, 0), ("emps", 100, 200, 9, 0)])
def test_algorithm1_matches_oracle(kind, T, N, M, cs):
    from oracle import marginal as OMg
    prob = HM.make_marg_problem(kind, T=T, N=N, M=M, seed=3)
    lam = prob["lam"]
    V = HM.make_variates(prob, lam, seed=11)
    ref = OMg.alg1_run(prob["oracle"], N, lam, HM.oracle_variates(V))
    REL = globals()["REL"]
    if N >= 200 and T >= 60:        # shipped sizes: conditioning-aware bar (1e-9 where the oracle is that stable)
        REL = _tolerance(lambda v: OMg.alg1_run(prob["oracle"], N, lam, v), V, ref,
                         lambda o: [o["state_trace"]] + [o["suff_stats"][g][j] for g in range(prob["G"]) for j in range(2)])
        assert REL < 1e-6
    A1 = helpers.pkg("Algorithm1").Algorithm1(forgetting_factor=lam, cluster_size=cs, **prob["prod_kwargs"])
    r = A1.filter(variates=HM.device_variates(V))
    assert int(r["status"][0]) == 0
    np.testing.assert_array_equal(_np(r["anc_trace"][0]), ref["anc_trace"] if "anc_trace" in ref else ref["ancestor_trace"])
    assert HM.rel_err(_np(r["state_trace"][0]), ref["state_trace"]) < REL
    for g in range(prob["G"]):
        assert HM.rel_err(_np(r["xi_trace"][0, g]), ref["int_var_trace"][g][..., 0]) < REL
        for j, sh in enumerate([(T, M), (T, M, M), (T,), (T,)]):
            assert HM.rel_err(_np(r["sst_trace"][4 * g + j][0]).reshape(sh), np.asarray(ref["suff_stats_trace"][g][j]).reshape(sh)) < REL
        for j, sh in enumerate([(N, M), (N, M, M), (N,), (N,)]):
            assert HM.rel_err(_np(r["final_stats"][4 * g + j][0]).reshape(sh), np.asarray(ref["suff_stats"][g][j]).reshape(sh)) < REL
    assert HM.rel_err(_np(r["logw_trace"][0]), ref["logw_trace"]) < max(REL, 1e-7 if T >= 60 else 0.0)    # differences of O(1e3) log-densities


@pytest.mark.parametrize("kind,T,N,M,cs", [("smo", 24, 32, 12, 0), ("smo", 10, 200, 41, 0), ("emps", 30, 48, 9, 2),
                                           ("vehicle", 24, 40, 8, 0),
                                           ("plugin", 24, 32, 10, 0), ("plugin", 40, 200, 20, 0), ("plugin", 20, 40, 10, 2),     # model plug-in
                                           # the shipped sizes, three refresh cycles of the rank-one updated factors
                                           ("smo", 100, 200, 41, 0), ("vehicle", 60, 200, 20, 0)])
def test_algorithm3_matches_oracle(kind, T, N, M, cs):
    import torch
    from oracle import marginal as OMg
    prob = HM.make_marg_problem(kind, T=T, N=N, M=M, seed=5)
    # reference trajectory: particle 0's path of an oracle filter run
    V1 = HM.make_variates(prob, 1.0, seed=21)
    f = OMg.alg1_run(prob["oracle"], N, 1.0, HM.oracle_variates(V1))
    ref_x = f["state_trace"][:, 0]
    ref_xi = [f["int_var_trace"][g][:, 0, 0] for g in range(prob["G"])]
    rs = OMg.reference_stats(prob["oracle"], ref_x, ref_xi)
    V = HM.make_variates(prob, 1.0, seed=22)
    ref = OMg.alg3_run(prob["oracle"], N, ref_x, ref_xi, rs, HM.oracle_variates(V))
    REL = globals()["REL"]
    if N >= 200 and T >= 60:        # shipped sizes: conditioning-aware bar (see _tolerance)
        REL = _tolerance(lambda v: OMg.alg3_run(prob["oracle"], N, ref_x, ref_xi, rs, v), V, ref, lambda o: [o["state_trace"], o["traj"]])
        assert REL < 1e-6
    LW = max(REL, 1e-7 if T >= 60 else 0.0)
    A3 = helpers.pkg("Algorithm3").Algorithm3(cluster_size=cs, **prob["prod_kwargs"])
    f64 = dict(dtype=torch.float64, device="cuda")
    rx = torch.as_tensor(ref_x[None], **f64)
    rxi = torch.as_tensor(np.stack(ref_xi)[None], **f64)
    # (a) reference statistics computed on the device
    dev_rs = A3.reference_stats(rx, rxi)
    for g in range(prob["G"]):
        for j in range(4):
            assert HM.rel_err(_np(dev_rs[4 * g + j][0]).reshape(-1), np.asarray(rs[g][j], dtype=np.float64).reshape(-1)) < REL
    # (b) the sweep, with the statistics passed in (Algorithm3.__call__ signature) and computed internally
    for stats in (dev_rs, None):
        r = A3.csmc(rx, rxi, stats, variates=HM.device_variates(V))
        assert int(r["status"][0]) == 0
        np.testing.assert_array_equal(_np(r["anc_trace"][0]), ref["anc_trace"])
        assert HM.rel_err(_np(r["state_trace"][0]), ref["state_trace"]) < REL
        assert HM.rel_err(_np(r["logw_trace"][0]), ref["logw_trace"]) < LW
        assert int(r["idx"][0]) == ref["idx"]
        assert HM.rel_err(_np(r["traj"][0]), ref["traj"]) < REL
        for g in range(prob["G"]):
            assert HM.rel_err(_np(r["xi_trace"][0, g]), ref["int_var_trace"][g][..., 0]) < REL
            assert HM.rel_err(_np(r["xi_traj"][0, g]), ref["xi_traj"][g][:, 0]) < REL


@pytest.mark.parametrize("kind", ["smo", "vehicle", "plugin"])
def test_algorithm2_matches_oracle(kind):
    import torch
    from oracle import marginal as OMg
    T, N, M, K = 16, 24, 8, 4
    prob = HM.make_marg_problem(kind, T=T, N=N, M=M, seed=7)
    V1 = HM.make_variates(prob, 1.0, seed=31)
    f = OMg.alg1_run(prob["oracle"], N, 1.0, HM.oracle_variates(V1))
    init_x = f["state_trace"][:, 1]
    init_xi = [f["int_var_trace"][g][:, 1, 0] for g in range(prob["G"])]
    V = HM.make_variates(prob, 1.0, seed=32, K=K)
    ref = OMg.alg2_run(prob["oracle"], N, K, init_x, init_xi, lambda k: HM.oracle_variates(V, k))
    A2 = helpers.pkg("Algorithm2").Algorithm2(N_iterations=K, **prob["prod_kwargs"])
    f64 = dict(dtype=torch.float64, device="cuda")
    r = A2.run(torch.as_tensor(init_x[None], **f64), torch.as_tensor(np.stack(init_xi)[None], **f64), variates=HM.device_variates(V))
    assert int(r["status"][0]) == 0
    assert HM.rel_err(_np(r["x_trace"][0]).transpose(1, 0, 2), ref["state_trace"]) < REL
    for g in range(prob["G"]):
        assert HM.rel_err(_np(r["xi_trace"][0, g]).T, ref["int_var_trace"][g][..., 0]) < REL
        for k in range(K):
            for j in range(4):
                assert HM.rel_err(_np(r["sst"][4 * g + j][0, k]).reshape(-1), np.asarray(ref["suff_stats_trace"][k][g][j], dtype=np.float64).reshape(-1)) < REL


def test_marginal_philox_mode_matches_oracle_fed_with_the_library_stream():
    """Philox mode: export the library's own variates (pgas_philox_marg_variates_f64) and feed them to the oracle."""
    import ctypes as C
    import torch
    from oracle import marginal as OMg
    L = helpers.pkg("_lib")
    A1m = helpers.pkg("Algorithm1")
    T, N, M = 20, 48, 10
    prob = HM.make_marg_problem("smo", T=T, N=N, M=M, seed=9)
    lam = 0.98
    A1 = A1m.Algorithm1(forgetting_factor=lam, **prob["prod_kwargs"])
    key = helpers.pkg("random").key(2024)
    r = A1.filter(key=key, chain_base=3)
    df = np.ascontiguousarray(HM.predictive_df(prob, lam))
    f64 = dict(dtype=torch.float64, device="cuda")
    Z, ZX, U, TS = torch.empty((1, T, N, 2), **f64), torch.empty((1, 1, N), **f64), torch.empty((1, T, 2), **f64), torch.empty((1, 1, T, N), **f64)
    rng = A1m.make_marg_rng(key, 3, 0)
    L.check(L.lib().pgas_philox_marg_variates_f64(C.byref(rng), 1, 1, T, N, 2, df.ctypes.data_as(C.POINTER(C.c_double)), L.ptr(Z), L.ptr(ZX),
                                                  L.ptr(U), L.ptr(TS), L.stream_ptr()))
    V = dict(Z=_np(Z[0]), ZXI0=_np(ZX[0]), U=_np(U[0]), TS=_np(TS[0]))
    assert abs(V["Z"].mean()) < 0.1 and abs(V["Z"].std() - 1) < 0.1 and 0.2 < V["U"].mean() < 0.8
    ref = OMg.alg1_run(prob["oracle"], N, lam, HM.oracle_variates(V))
    np.testing.assert_array_equal(_np(r["anc_trace"][0]), ref["ancestor_trace"])
    assert HM.rel_err(_np(r["state_trace"][0]), ref["state_trace"]) < REL
    assert HM.rel_err(_np(r["xi_trace"][0, 0]), ref["int_var_trace"][0][..., 0]) < REL


def test_marginal_chains_are_independent_of_batching():
    """chain ids, not launch geometry, define the streams: 3 chains at once == chain 2 alone"""
    A1m = helpers.pkg("Algorithm1")
    prob = HM.make_marg_problem("emps", T=16, N=32, M=9, seed=2)
    A1 = A1m.Algorithm1(forgetting_factor=0.999, **prob["prod_kwargs"])
    key = helpers.pkg("random").key(77)
    a = A1.filter(key=key, n_chains=3, chain_base=0)
    b = A1.filter(key=key, n_chains=1, chain_base=2)
    np.testing.assert_array_equal(_np(a["anc_trace"][2]), _np(b["anc_trace"][0]))
    np.testing.assert_array_equal(_np(a["state_trace"][2]), _np(b["state_trace"][0]))


def test_log_base_measure_batched():
    import torch
    from oracle import mniw as OM
    L = helpers.pkg("_lib")
    rng = np.random.default_rng(0)
    n, M = 9, 17
    Phi = rng.normal(size=(n, 40, M))
    y = rng.normal(size=(n, 40))
    T1 = np.einsum("ntm,ntk->nmk", Phi, Phi) + np.eye(M)
    T0 = np.einsum("ntm,nt->nm", Phi, y)
    T2 = np.einsum("nt,nt->n", y, y) + 1.0
    T3 = 40.0 + rng.integers(0, 5, size=n)
    ref = np.array([OM.prior_mniw_log_base_measure(T0[i][:, None], T1[i], np.array([[T2[i]]]), T3[i]) for i in range(n)])
    f64 = dict(dtype=torch.float64, device="cuda")
    out = torch.empty(n, **f64)
    d = [torch.as_tensor(np.ascontiguousarray(a), **f64) for a in (T0, T1, T2, T3)]
    L.check(L.lib().pgas_mniw_log_base_measure_f64(*[L.ptr(t) for t in d], n, M, L.ptr(out), L.stream_ptr()))
    assert HM.rel_err(_np(out), ref) < REL


def test_reference_api_tuples():
    """Algorithm1.__call__ / Algorithm3.__call__ / Algorithm2.__call__ return the reference's tuples and shapes"""
    A1m, A2m, A3m = helpers.pkg("Algorithm1"), helpers.pkg("Algorithm2"), helpers.pkg("Algorithm3")
    T, N, M, K = 12, 24, 8, 3
    prob = HM.make_marg_problem("vehicle", T=T, N=N, M=M, seed=4)
    key = helpers.pkg("random").key(5)
    out = A1m.Algorithm1(forgetting_factor=0.999, **prob["prod_kwargs"])(key)
    st, iv, sst, w, anc, fin, obs, ll = out
    assert st.shape == (T, N, 2) and len(iv) == 2 and iv[0].shape == (T, N, 1) and w.shape == (T, N)
    assert anc.shape == (T - 1, N) and anc.dtype == np.int32 and obs.shape == (T, N, 2) and ll.shape == (T, N)
    assert sst[1][0].shape == (T, M, 1) and sst[1][1].shape == (T, M, M) and sst[1][2].shape == (T, 1, 1) and sst[1][3].shape == (T,)
    assert fin[0][1].shape == (N, M, M) and np.allclose(w.sum(axis=1), 1.0)
    ref_x, ref_xi = st[:, 0], [iv[g][:, 0, 0] for g in range(2)]
    A2 = A2m.Algorithm2(N_iterations=K, **prob["prod_kwargs"])
    s2, iv2, w2, sst2, obs2, ll2 = A2(key, ref_x, ref_xi)
    assert s2.shape == (T, K, 2) and iv2[1].shape == (T, K, 1) and w2.shape == (T, K) and obs2.shape == (T, K, 2) and ll2.shape == (T, K)
    assert sst2[0][1].shape == (K, M, M) and np.allclose(s2[:, 0], ref_x) and np.all(np.isfinite(s2))
    A3 = A3m.Algorithm3(**prob["prod_kwargs"])
    traj, xit = A3(key, ref_x, ref_xi, [[sst2[g][j][0] for j in range(4)] for g in range(2)])
    assert traj.shape == (T, 2) and len(xit) == 2 and xit[0].shape == (T,)


def test_shipped_smo_pipeline_recovers_the_spring_damper_force(tmp_path):
    """SURVEY.md 8c pin (4): the shipped single-mass-oscillator pipeline (Algorithm1 -> Algorithm2, N=200, T=750, M=41,
    reduced to 30 Gibbs iterations) learns F_sd on the visited states: posterior-mean RMSE well below the force's RMS
    (measured 0.057 of it at K=40 on B200).  Replaces the comparison with plots/SingleMassOscillator.mat, a Git-LFS stub."""
    import os
    import sys
    import scipy.io
    sys.path.insert(0, os.path.join(helpers.ROOT, "drivers"))
    import run_example
    out = str(tmp_path / "smo.mat")
    md, summary = run_example.run("smo", iterations=30, out=out, quiet=True)
    assert summary["rmse_F_on_trajectory"] < 0.12 * summary["rms_F_true"], summary
    back = scipy.io.loadmat(out)
    for k in ("offline_Sigma_X", "offline_Sigma_F", "offline_T1", "online_T0", "online_weights", "basis_plot", "prior_T1", "F_sd_true_plot"):
        assert k in back
    assert back["offline_Sigma_X"].shape == (750, 30, 2) and back["offline_T1"].shape == (30, 41, 41)
    assert back["online_Sigma_F"].shape == (750, 200, 1) and back["basis_plot"].shape == (2500, 41)
    assert np.all(np.isfinite(back["offline_Sigma_X"])) and np.all(np.isfinite(back["online_log_likelihood"]))


@pytest.mark.parametrize("T,N,M", [(300, 64, 12), (160, 200, 41)])
def test_algorithm3_long_sweep_stays_on_the_oracle(T, N, M):
    """The rank-one up/down-dated factors (refreshed every 32 steps) must not drift: a 300-step conditional sweep —
    nine refresh cycles — still reproduces every ancestor index of the re-factorising oracle and its trajectory to 1e-9."""
    import torch
    from oracle import marginal as OMg
    prob = HM.make_marg_problem("smo", T=T, N=N, M=M, seed=12)
    V1 = HM.make_variates(prob, 1.0, seed=41)
    f = OMg.alg1_run(prob["oracle"], N, 1.0, HM.oracle_variates(V1))
    ref_x, ref_xi = f["state_trace"][:, 0], [f["int_var_trace"][0][:, 0, 0]]
    rs = OMg.reference_stats(prob["oracle"], ref_x, ref_xi)
    V = HM.make_variates(prob, 1.0, seed=42)
    ref = OMg.alg3_run(prob["oracle"], N, ref_x, ref_xi, rs, HM.oracle_variates(V))
    A3 = helpers.pkg("Algorithm3").Algorithm3(**prob["prod_kwargs"])
    f64 = dict(dtype=torch.float64, device="cuda")
    r = A3.csmc(torch.as_tensor(ref_x[None], **f64), torch.as_tensor(np.stack(ref_xi)[None], **f64), None, variates=HM.device_variates(V))
    assert int(r["status"][0]) == 0
    np.testing.assert_array_equal(_np(r["anc_trace"][0]), ref["anc_trace"])
    assert HM.rel_err(_np(r["state_trace"][0]), ref["state_trace"]) < REL
    assert HM.rel_err(_np(r["xi_trace"][0, 0]), ref["int_var_trace"][0][..., 0]) < REL
    assert HM.rel_err(_np(r["logw_trace"][0]), ref["logw_trace"]) < 1e-7       # differences of O(1e3) log-densities
    assert int(r["idx"][0]) == ref["idx"]


def test_replicas_api_single_process():
    """run_replicas_distributed without a process group = all replicas on this GPU; chain ids define the streams"""
    A2m, DI = helpers.pkg("Algorithm2"), helpers.pkg("distributed")
    prob = HM.make_marg_problem("emps", T=14, N=24, M=9, seed=1)
    A2 = A2m.Algorithm2(N_iterations=3, **prob["prod_kwargs"])
    key = helpers.pkg("random").key(9)
    x0 = np.stack([0.1 + 0.001 * np.arange(14), np.zeros(14)], axis=1)
    r3 = DI.run_replicas_distributed(A2, key, x0, [np.zeros(14)], 3)
    assert tuple(r3["x_trace"].shape) == (3, 3, 14, 2) and tuple(r3["xi_trace"].shape) == (3, 1, 3, 14)
    r1 = DI.run_replicas_distributed(A2, key, x0, [np.zeros(14)], 1)
    np.testing.assert_array_equal(_np(r3["x_trace"][0]), _np(r1["x_trace"][0]))
    assert not np.array_equal(_np(r3["x_trace"][1, 1:]), _np(r3["x_trace"][0, 1:]))


def test_plugin_model_runs_as_expression_programs():
    """model plug-in (src/StateSpaceModel.py:19-30): every callable of the pendulum model is outside the
    coefficient-table families, so all three run as interpreted programs; the reference-API call returns the 8-tuple and its
    observation / log-likelihood traces are the user's output model on the state trace"""
    prob = HM.make_marg_problem("plugin", T=30, N=48, M=10, seed=2)
    A1 = helpers.pkg("Algorithm1").Algorithm1(forgetting_factor=0.999, **prob["prod_kwargs"])
    assert set(A1.model.programs) == {"transition_model", "output_model", "basis_fcn[0]"}
    out = A1(helpers.pkg("random").key(5))
    assert len(out) == 8
    st, obs, ll = out[0], out[6], out[7]
    assert st.shape == (30, 48, 2) and obs.shape == (30, 48) and np.isfinite(st).all()
    want = 1.5 * np.sin(st[..., 0]) + 0.05 * st[..., 1] ** 2
    assert np.allclose(obs, want, rtol=1e-12, atol=1e-14)
    Y = prob["prod_kwargs"]["observations"]
    want_ll = -0.5 * (Y[:, None] - want) ** 2 / 1e-3 - 0.5 * np.log(2 * np.pi * 1e-3)
    assert np.allclose(ll, want_ll, rtol=1e-9, atol=1e-9)
    # a shipped model keeps the table-driven kernels
    smo = HM.make_marg_problem("smo", T=8, N=16, M=6)
    assert helpers.pkg("Algorithm1").Algorithm1(forgetting_factor=0.999, **smo["prod_kwargs"]).model.programs == {}
