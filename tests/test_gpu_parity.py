"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Tolerances: helpers.TIE_TOL (1e-12, CDF ties) and helpers.REL_TOL (1e-9)."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _dev(a):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


# ----------------------------------------------------------------------------- A1 resampling
@pytest.mark.parametrize("N", [1, 2, 7, 200, 1000, 4096, 16384])
def test_systematic_resampling_matches_oracle(built_lib, N):
    from oracle import filtering as OF
    F = helpers.pkg("Filtering")
    rng = np.random.default_rng(N)
    sets = 6
    W = rng.exponential(size=(sets, N)) ** 3
    W[1] = 1.0                      # uniform weights -> arange(N)
    W[2, : N // 2] = 0.0            # zeros
    if N > 2:
        W[3] = 0.0
        W[3, N // 3] = 1.0          # degenerate: a single particle carries everything
    W[4] = -1.0                     # all clipped to zero -> uniform fallback (src/Filtering.py:25)
    W[5] = W[5] * 1e-300            # denormal-scale weights
    u = rng.uniform(size=sets)
    got = F.systematic_SISR(u, W)
    for s in range(sets):
        ref = OF.systematic_SISR(u[s], W[s])
        w = np.clip(W[s], 0, np.inf)
        cdf = helpers.resample_cdf(w) if w.sum() > 0 else np.clip(np.cumsum(np.ones(N) / N), 0, 1)
        ok, nbad, gap = helpers.index_mismatch_is_tie(got[s], ref, cdf, (u[s] + np.arange(N)) / N)
        assert ok, (s, nbad, gap)
    assert np.array_equal(got[1], np.arange(N))
    assert got.min() >= 0 and got.max() <= N - 1


def test_resampling_nan_weights_fall_back_to_uniform(built_lib):
    F = helpers.pkg("Filtering")
    w = np.array([0.1, np.nan, 0.3, 0.2])
    assert np.array_equal(F.systematic_SISR(0.5, w), np.arange(4))     # NaN sum -> 1/N weights (src/Filtering.py:24-25)


# ----------------------------------------------------------------------------- A2 reconstruct
def test_reconstruct_trajectory_matches_oracle(built_lib):
    from oracle import filtering as OF
    F = helpers.pkg("Filtering")
    rng = np.random.default_rng(0)
    for (T, N, n) in [(1, 5, 2), (2, 3, 1), (50, 64, 2), (200, 300, 3)]:
        P = rng.normal(size=(T, N, n))
        anc = rng.integers(0, N, size=(max(T - 1, 1), N))
        idx = int(rng.integers(0, N))
        got = F.reconstruct_trajectory(P if n > 1 else P[..., 0], anc, idx)
        ref = OF.reconstruct_trajectory(P if n > 1 else P[..., 0], anc, idx)
        assert got.shape == ref.shape
        assert np.array_equal(got, ref)        # pure gather: bit exact


# ----------------------------------------------------------------------------- A3/A4 basis
@pytest.mark.parametrize("kind", ["smo", "emps", "toy", "vehicle"])
def test_basis_values_match_oracle(built_lib, kind):
    p = helpers.make_problem(kind)
    hgp, sd = helpers.pkg("BasisFunctions").generate_Hilbert_BasisFunction(*p["hgp_args"])
    assert np.array_equal(hgp.freq, p["ohgp"].indices)
    rng = np.random.default_rng(1)
    D = hgp.D
    X = hgp.center + (rng.uniform(-1.2, 1.2, size=(500, D))) * hgp.half_width       # inside and outside the domain
    got = hgp(X if D > 1 else X[:, 0])
    ref = p["ohgp"].batch(X)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) < 1e-12 * np.max(np.abs(ref)) * 10
    one = hgp(X[0] if D > 1 else X[0, 0])
    assert one.shape == (hgp.M,) and np.allclose(one, ref[0], rtol=0, atol=1e-12)


def test_basis_large_lattices(built_lib):
    from oracle import basis as OB
    BF = helpers.pkg("BasisFunctions")
    for M, dom in [(256, [[-7.5, 7.5]] * 2), (1024, [[-0.5, 0.5]] * 2), (729, [[-1, 1]] * 3)]:
        hgp, sd = BF.generate_Hilbert_BasisFunction(M, np.array(dom, dtype=float), 0.1, 10.0)
        oh, osd = OB.generate_Hilbert_BasisFunction(M, np.array(dom, dtype=float), 0.1, 10.0)
        assert np.array_equal(hgp.freq, oh.indices)
        X = np.random.default_rng(M).uniform(-1, 1, size=(64, hgp.D)) * hgp.half_width + hgp.center
        got, ref = hgp(X), oh.batch(X)
        assert np.max(np.abs(got - ref)) < 5e-12 * np.max(np.abs(ref))


# ----------------------------------------------------------------------------- A9 step
@pytest.mark.parametrize("kind,N,cluster", [
    ("smo", 200, 1), ("smo", 200, 2), ("smo", 1000, 1), ("smo", 1000, 4), ("smo", 777, 8), ("smo", 2048, 16),
    ("emps", 200, 1), ("emps", 512, 2), ("toy", 200, 1), ("toy", 300, 4), ("vehicle", 200, 1), ("vehicle", 640, 8),
    ("smo", 20, 16), ("smo", 2600, 1),
    ("pluginlik", 200, 1), ("pluginlik", 1000, 4),     # model plug-in: likelihood_fcn as an expression program (Student-t, non-affine output map)
])
def test_step_parity_teacher_forced(built_lib, kind, N, cluster):
    p = helpers.make_problem(kind, T=12, N=N, seed=N + cluster)
    r = helpers.run_step_parity(p, n_steps=8, cluster_size=cluster)
    assert r["ok"], str(r)


@pytest.mark.parametrize("flags", [1, 2, 3])
def test_step_parity_quirk_flags(built_lib, flags):
    for cluster in (1, 4):
        p = helpers.make_problem("smo", T=12, N=300, seed=9, flags=flags)
        r = helpers.run_step_parity(p, n_steps=6, cluster_size=cluster)
        assert r["ok"], (flags, cluster, r)


# ----------------------------------------------------------------------------- A11 sweep
@pytest.mark.parametrize("kind,N,T,cluster", [
    ("smo", 200, 60, 1), ("smo", 512, 40, 4), ("smo", 4096, 12, 16), ("emps", 200, 50, 1), ("toy", 200, 40, 2),
    ("vehicle", 256, 40, 1), ("smo", 200, 60, 0),
    ("smo", 4096, 12, 0), ("smo", 2501, 16, 0), ("vehicle", 6000, 8, 0),     # split form with the dedicated resampling kernel: clusters of 2 / 2 / 4
    # model plug-in: a GP-input map of sines / tanh / a rational term, traced into an expression program (split form; fused kernel)
    ("plugin", 512, 40, 0), ("plugin", 200, 30, 1), ("plugin", 4096, 12, 0),
    # ... and a likelihood_fcn outside the Gaussian family as an expression program (fused kernel, with and without workspace / clusters)
    ("pluginlik", 512, 40, 0), ("pluginlik", 200, 30, 1), ("pluginlik", 2048, 12, 4),
])
def test_sweep_parity_injected(built_lib, kind, N, T, cluster):
    p = helpers.make_problem(kind, T=T, N=N, seed=T)
    r = helpers.run_sweep_parity(p, cluster_size=cluster)
    assert r["ok"], str(r)
    assert r["rows_compared"] >= 1


@pytest.mark.parametrize("kind,N,T,M", [("smo", 700, 40, 41), ("vehicle", 1536, 24, 36), ("smo", 512, 20, 256), ("plugin", 300, 30, 41)])
def test_state_kernel_dmma_form_matches_oracle(built_lib, kind, N, T, M, monkeypatch):
    """The contraction of the state kernel on DMMA tiles with the particles on the N dimension (csrc/basis_mma.cuh; opt-in,
    PGAS_STATE_MMA=1): same oracle parity as the FMA row walk — ragged particle counts, both observation dimensions, a lattice with
    five row blocks, the expression-program map."""
    monkeypatch.setenv("PGAS_STATE_MMA", "1")
    monkeypatch.setenv("PGAS_STATE_SMALL", "0")
    p = helpers.make_problem(kind, T=T, N=N, M=M, seed=T + 1)
    r = helpers.run_sweep_parity(p, cluster_size=0)
    assert r["ok"], str(r)
    r = helpers.run_sweep_parity(p, cluster_size=0, philox_seed=77)
    assert r["ok"], str(r)


def test_sweep_gather_mode(built_lib):
    p = helpers.make_problem("smo", T=30, N=384, seed=4, flags=1)
    for cluster in (1, 2):
        r = helpers.run_sweep_parity(p, cluster_size=cluster)
        assert r["ok"], str(r)


def test_sweep_philox_stream_matches_restatement(built_lib):
    """Philox mode: the oracle is fed the library's own stream (oracle/philox.py restates it)."""
    import torch
    import ctypes as C
    from oracle import philox as OPH
    L = helpers.pkg("_lib")
    T, N, n_x, seed = 9, 130, 2, 0xDEADBEEF12345
    rng = L.Rng()
    rng.mode, rng.seed, rng.chain_base, rng.iteration = 0, seed, 2, 3
    Z = torch.empty((1, T, N, n_x), dtype=torch.float64, device="cuda")
    U = torch.empty((1, T, 2), dtype=torch.float64, device="cuda")
    L.check(L.lib().pgas_philox_sweep_variates_f64(C.byref(rng), 1, T, N, n_x, L.ptr(Z), L.ptr(U), L.stream_ptr()))
    Zo, Uo = OPH.sweep_variates(seed, 2, 3, T, N, n_x)
    assert np.array_equal(U[0].cpu().numpy(), Uo)                       # integer -> double mapping: exact
    assert np.max(np.abs(Z[0].cpu().numpy() - Zo)) < 1e-13              # log/sincos differ in the last ulps
    p = helpers.make_problem("smo", T=T, N=N, seed=2)
    r = helpers.run_sweep_parity(p, cluster_size=1, philox_seed=seed)
    assert r["ok"], str(r)


def test_chain_ids_give_identical_results_regardless_of_batching(built_lib):
    """the Philox counter carries the global chain id: chain c gives the same trajectory alone or in a batch"""
    import torch
    p = helpers.make_problem("smo", T=20, N=128, seed=6)
    cs = helpers.product_csmc(p, 1)
    key = helpers.pkg("random").key(77)
    ref = _dev(np.stack([p["ref"]] * 3))
    Th = _dev(np.stack([p["Theta"]] * 3))
    Sg = _dev(np.stack([p["Sigma"]] * 3))
    full = cs.sweep(ref, Th, Sg, key=key, chain_base=0)
    solo = cs.sweep(ref[2:3], Th[2:3], Sg[2:3], key=key, chain_base=2)
    assert torch.equal(full["state_trace"][2], solo["state_trace"][0])
    assert torch.equal(full["anc_trace"][2], solo["anc_trace"][0])
    assert not torch.equal(full["state_trace"][0], full["state_trace"][1])


# ----------------------------------------------------------------------------- A13 statistics + draw
@pytest.mark.parametrize("kind,M,T", [("smo", 41, 80), ("smo", 100, 300), ("smo", 256, 120), ("emps", 64, 60),
                                      ("emps", 200, 90), ("toy", 40, 40), ("vehicle", 36, 100), ("plugin", 64, 120),
                                      # configuration-scale bases: EMPS PGAS baseline (3-D, M = 729, configs[2]), configs[3] (M = 256)
                                      # at its full T, configs[4] (vehicle lattice, M = 1024)
                                      ("emps", 729, 400), ("smo", 256, 2000), ("vehicle", 1024, 700), ("smo", 1024, 300)])
def test_suffstats_and_draw_parity(built_lib, kind, M, T):
    p = helpers.make_problem(kind, T=T, N=32, M=M, seed=M)
    r = helpers.run_draw_parity(p)
    assert r["ok"], str(r)


@pytest.mark.parametrize("kind,M,T,flags", [("smo", 100, 150, 0), ("smo", 130, 90, 4), ("emps", 200, 90, 0), ("smo", 41, 60, 0)])
def test_blocked_and_one_cta_factorisations_agree(built_lib, monkeypatch, kind, M, T, flags):
    """the multi-CTA blocked Cholesky (default for M >= 192) and the one-CTA form give the same draw; both against the oracle"""
    p = helpers.make_problem(kind, T=T, N=32, M=M, seed=M + 1, flags=flags)
    for mode in ("1", "0"):
        monkeypatch.setenv("PGAS_DRAW_BLOCKED", mode)
        r = helpers.run_draw_parity(p)
        assert r["ok"], (mode, str(r))


def test_draw_transpose_flag(built_lib):
    p = helpers.make_problem("smo", T=50, N=32, M=41, seed=1, flags=4)
    r = helpers.run_draw_parity(p)
    assert r["ok"], str(r)


def test_draw_philox_stream(built_lib):
    import torch
    from oracle import philox as OPH, pgas as OP
    p = helpers.make_problem("smo", T=40, N=32, M=41, seed=3)
    pg = helpers.product_pgas(p, K=2)
    seed = 99
    df = p["prior"][3] + p["T"] - 1
    chi2, G, Nrm = OPH.draw_variates(seed, 5, 7, p["n_x"], p["M"], df)
    A_o, S_o, _ = OP.sample_params(p["omodel"], p["prior"], p["ref"], chi2, G, Nrm)
    A_g, S_g = pg.sample_params(helpers.pkg("random").key(seed), _dev(p["ref"][None]), chain_base=5, iteration=7)
    assert helpers.rel_err(A_g[0].cpu().numpy(), A_o) < helpers.REL_TOL
    assert helpers.rel_err(S_g[0].cpu().numpy(), S_o) < helpers.REL_TOL


def test_draw_rejects_indefinite_eta1(built_lib):
    import torch
    BI, PG = helpers.pkg("BayesianInferrence"), helpers.pkg("PGAS")
    M, nx = 40, 2
    e1 = torch.eye(M, dtype=torch.float64, device="cuda")[None].clone()
    e1[0, 7, 7] = -1.0
    rng = PG._make_rng(helpers.pkg("random").key(1))
    A, S, status = BI.mniw_posterior_draw(torch.zeros((1, M, nx), dtype=torch.float64, device="cuda"), e1,
                                          torch.eye(nx, dtype=torch.float64, device="cuda")[None].clone(), 10.0, rng)
    assert int(status[0]) == M - 7          # pivot position in the reversed factorisation order (1-based)
    # the multi-CTA blocked factorisation reports the same position
    M = 200
    e1 = torch.eye(M, dtype=torch.float64, device="cuda")[None].repeat(2, 1, 1)
    e1[1, 150, 150] = -1.0
    A, S, status = BI.mniw_posterior_draw(torch.zeros((2, M, nx), dtype=torch.float64, device="cuda"), e1,
                                          torch.eye(nx, dtype=torch.float64, device="cuda")[None].repeat(2, 1, 1), 10.0, rng)
    assert int(status[0]) == 0 and int(status[1]) == M - 150


# ----------------------------------------------------------------------------- A14 full Gibbs loop
@pytest.mark.parametrize("kind,cluster", [("smo", 1), ("smo", 2), ("toy", 1), ("plugin", 0), ("pluginlik", 0)])
def test_run_chains_matches_oracle(built_lib, kind, cluster):
    from oracle import pgas as OP
    import torch
    K, T, N = 4, 16, 96
    p = helpers.make_problem(kind, T=T, N=N, seed=11)
    pg = helpers.product_pgas(p, K=K, cluster_size=cluster)
    rng = np.random.default_rng(5)
    n_x, M = p["n_x"], p["M"]
    df = p["prior"][3] + T - 1
    V = dict(Z=rng.normal(size=(K, 1, T, N, n_x)), U=rng.uniform(size=(K, 1, T, 2)),
             chi2=rng.chisquare(df - np.arange(n_x), size=(K, 1, n_x)), G=rng.normal(size=(K, 1, n_x, n_x)),
             Nrm=rng.normal(size=(K, 1, n_x, M)))
    out = pg.run_chains(None, p["ref"], n_chains=1, variates={k: _dev(v) for k, v in V.items()})
    st_o, ll_o, A_o, S_o = OP.pgas_run(p["omodel"], N, K, p["prior"], p["ref"],
                                       lambda k: {n: V[n][k, 0] for n in V})
    st_g = out["state_trace"][0].cpu().numpy()            # (K,T,n_x)
    assert np.array_equal(st_g[0], p["ref"])
    # Iteration k is compared as long as iteration k-1 agreed: the Gibbs chain feeds each sweep the previous
    # draw, and the sampled dynamics Theta^(k) are not contractive in general, so rounding-level differences
    # (or one flipped CDF tie) legitimately grow from one iteration to the next.  Every sweep / draw is
    # additionally replayed from the ORACLE's previous iterate, which isolates each call of the loop.
    cs = pg.cSMC
    agreed = 0
    for k in range(K):
        if k >= 1:
            sw = cs.sweep(_dev(st_o[:, k - 1]), _dev(A_o[k - 1]), _dev(S_o[k - 1]),
                          variates=dict(Z=_dev(V["Z"][k]), U=_dev(V["U"][k])))
            o = OP.csmc_sweep(p["omodel"], N, st_o[:, k - 1], A_o[k - 1], S_o[k - 1], V["Z"][k, 0], V["U"][k, 0])
            assert np.array_equal(sw["anc_trace"][0].cpu().numpy(), o["anc_trace"]), k
            assert int(sw["idx"][0]) == o["idx"], k
            assert helpers.rel_err(sw["state_trace"][0, 1].cpu().numpy(), o["state_trace"][1]) < helpers.REL_TOL, k
        A_g, S_g = pg.sample_params(None, _dev(st_o[:, k][None]),
                                    variates=dict(chi2=_dev(V["chi2"][k]), G=_dev(V["G"][k]), Nrm=_dev(V["Nrm"][k])))
        assert helpers.rel_err(A_g[0].cpu().numpy(), A_o[k]) < helpers.REL_TOL, k
        assert helpers.rel_err(S_g[0].cpu().numpy(), S_o[k]) < helpers.REL_TOL, k
        if agreed == k and helpers.rel_err(st_g[k], st_o[:, k]) < 1e-7 and \
                helpers.rel_err(out["A_trace"][0, k].cpu().numpy(), A_o[k]) < 1e-7:
            agreed = k + 1
    assert agreed >= 2, agreed                 # initial draw and the first full Gibbs iteration
    if kind == "smo":
        assert agreed == K, agreed             # contractive synthetic dynamics: the whole run agrees


def test_pgas_reference_call_signature(built_lib):
    p = helpers.make_problem("smo", T=14, N=64, seed=2)
    pg = helpers.product_pgas(p, K=3)
    st, ll = pg(helpers.pkg("random").key(12345678), p["ref"])
    assert st.shape == (14, 3, 2) and ll.shape == (14, 3)
    assert np.array_equal(st[:, 0], p["ref"]) and np.all(np.isfinite(st)) and np.all(np.isfinite(ll))
    from oracle import pgas as OP
    ll_o = np.stack([p["omodel"].loglik(p["obs"][t], st[t], p["inputs"][t]) for t in range(14)])
    assert np.allclose(ll, ll_o, rtol=1e-10, atol=1e-10)
    cs = helpers.product_csmc(p)
    tr = cs(helpers.pkg("random").key(1), p["ref"], p["Theta"], p["Sigma"])
    assert tr.shape == (14, 2) and np.all(np.isfinite(tr))


@pytest.mark.parametrize("kind,N,T,chains,cluster,dedicated", [
    ("smo", 300, 40, 3, 0, 1), ("vehicle", 700, 50, 2, 2, 1), ("smo", 1100, 150, 2, 4, 1),
    ("smo", 4096, 70, 2, 0, 1),          # dedicated resampling kernel, one CTA per chain walking two slices (the bench shape)
    ("smo", 5001, 20, 1, 0, 1),          # cluster of 4, ragged last CTA, odd N (scalar loads / stores)
    ("vehicle", 8192, 12, 1, 0, 1),      # cluster of 4, full
    ("smo", 300, 40, 3, 0, 2), ("smo", 4096, 70, 2, 0, 2), ("smo", 2500, 30, 1, 0, 2),   # cluster form forced (1 / 2 / 2 CTAs)
    ("smo", 2049, 30, 2, 0, 1), ("vehicle", 3000, 20, 1, 0, 1),   # one CTA, two slices, ragged second slice
    ("smo", 300, 40, 3, 0, 0), ("smo", 4096, 70, 2, 0, 0),    # general resampling kernel (csmc_sweep_kernel<PRE>)
    # latency form (weights_lat.cu: offspring scatter, st.async + mbarrier hand-offs): clusters of 1 / 8 / 10 (ragged, odd N) / 16 / 5 / 2
    ("smo", 300, 40, 3, 0, 3), ("smo", 4096, 150, 2, 0, 3), ("smo", 5001, 20, 1, 0, 3), ("vehicle", 8192, 12, 1, 0, 3),
    ("smo", 2049, 30, 2, 0, 3), ("vehicle", 700, 50, 2, 2, 3),
    # three-dimensional basis (EMPS baseline shape, src/EMPS.py:101-123): the row walk with one more level, both geometries
    ("emps", 200, 80, 1, 0, 3), ("emps", 1500, 40, 3, 0, 1), ("emps", 700, 30, 2, 0, 0),
    # latency form with four particles per thread (N > 8192): clusters of 10 (ragged) and 16
    ("smo", 10001, 12, 1, 0, 3), ("vehicle", 16384, 10, 2, 0, 3)])
def test_split_and_fused_sweeps_agree(built_lib, kind, N, T, chains, cluster, dedicated, monkeypatch):
    """The split form (state kernel ahead of the resampling kernel, csrc/sweep.cu) and the fused kernel are two schedules
    of the same arithmetic: identical ancestors and traces, for particle counts that are not multiples of the tile sizes,
    several chains, chunk boundaries (T > 64 + 1) and both observation dimensions."""
    import os
    import torch
    monkeypatch.setenv("PGAS_WEIGHTS_KERNEL", str(dedicated))
    if dedicated != 3:          # state kernel: the fill-the-GPU geometry <256, 2>; the latency-form cases run the few-chains one <64, 1>
        monkeypatch.setenv("PGAS_STATE_SMALL", "0")
    p = helpers.make_problem(kind, T=T, N=N, seed=5)
    cs = helpers.product_csmc(p, cluster)
    dev = lambda x: torch.as_tensor(np.ascontiguousarray(x)).cuda()
    ref, Th, Sg = (dev(np.stack([p[k]] * chains)) for k in ("ref", "Theta", "Sigma"))
    key = helpers.pkg("random").key(11)
    os.environ.pop("PGAS_SWEEP_FUSED", None)
    a = cs.sweep(ref, Th, Sg, key=key, chain_base=4)
    os.environ["PGAS_SWEEP_FUSED"] = "1"
    try:
        b = cs.sweep(ref, Th, Sg, key=key, chain_base=4)
    finally:
        os.environ.pop("PGAS_SWEEP_FUSED", None)
    for k in ("anc_trace", "idx"):
        assert torch.equal(a[k], b[k]), k
    # two-dimensional bases: both forms run the same row walk (same bits up to the contraction of fused multiply-adds); three-
    # dimensional ones: the split form walks the lattice with FMAs, the fused kernel contracts DMMA tiles — another summation order
    rtol, atol = (1e-13, 1e-300) if kind != "emps" else (1e-9, 1e-12)
    for k in ("state_trace", "logw_last", "traj"):
        assert torch.allclose(a[k], b[k], rtol=rtol, atol=atol), k
    assert bool(torch.isfinite(a["state_trace"]).all())


@pytest.mark.parametrize("kernel", [0, 1, 2, 3])
def test_sweep_with_dead_and_degenerate_weights(built_lib, kernel, monkeypatch):
    """Weights that vanish for whole warps and a step whose weights are all -inf, in every resampling kernel of the split sweep
    (0 general, 1 one CTA per chain, 2 cluster form, 3 latency form): jax.nn.softmax gives -inf entries weight 0, and
    systematic_SISR falls back to uniform weights when the sum is not > 0 (src/Filtering.py:24-25) — after which every
    log-weight is NaN and every later step is uniform too.  Ancestors must equal the oracle's exactly."""
    import warnings
    import torch
    from oracle import pgas as OP
    monkeypatch.setenv("PGAS_WEIGHTS_KERNEL", str(kernel))
    p = helpers.make_problem("smo", T=24, N=1024, seed=9)
    p["obs"] = np.array(p["obs"], dtype=np.float64, copy=True)
    p["obs"][17] = 1.0e200                                  # log p(y_17 | .) = -inf for every particle (the squared residual overflows)
    p["omodel"].observations = p["obs"]
    Z, U = helpers.sweep_variates(p, 4)
    Z[5, 128:384] = 1.0e3                                   # 256 consecutive particles thrown far off: log-weights ~ -5e8 -> weight 0
    cs = helpers.product_csmc(p)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    g = cs.sweep(dev(p["ref"]), dev(p["Theta"]), dev(p["Sigma"]), variates=dict(Z=dev(Z[None]), U=dev(U[None])))
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        o = OP.csmc_sweep(p["omodel"], p["N"], p["ref"], p["Theta"], p["Sigma"], Z, U, keep_weights=True, **helpers.oracle_flags(p))
    an_g, an_o = g["anc_trace"][0].cpu().numpy(), np.asarray(o["anc_trace"])
    for t in range(1, p["T"]):
        assert np.array_equal(an_g[t - 1], an_o[t - 1]), (kernel, t, int((an_g[t - 1] != an_o[t - 1]).sum()))
    assert np.array_equal(an_o[17 - 1][:-1], np.arange(p["N"] - 1))          # the degenerate step resampled uniformly
    assert helpers.rel_err(g["state_trace"][0].cpu().numpy(), o["state_trace"]) < helpers.REL_TOL


def test_run_chains_resume_is_bit_exact(built_lib):
    """checkpoint / resume: the Gibbs state is (reference trajectory, key, iteration index);
    6 iterations == 4 iterations + a resumed run of 3 starting from trajectory 3"""
    import torch
    p = helpers.make_problem("smo", T=40, N=256, seed=8)
    pg = helpers.product_pgas(p, K=6)
    key = helpers.pkg("random").key(21)
    full = pg.run_chains(key, p["ref"], n_chains=2, chain_base=5)
    first = pg.run_chains(key, p["ref"], n_chains=2, chain_base=5, K=4)
    rest = pg.run_chains(key, first["state_trace"][:, 3].contiguous(), n_chains=2, chain_base=5, iteration=3, K=3)
    assert torch.equal(full["state_trace"][:, :4], first["state_trace"])
    assert torch.equal(full["state_trace"][:, 3:], rest["state_trace"])
    assert torch.equal(full["A_trace"][:, 3:], rest["A_trace"])


def test_split_and_fused_sweeps_agree_at_config5_shape(built_lib):
    """BASELINE.json configs[4] shape: N=16384 particles, 2-D basis with M=1024 (max lattice index 36), cluster of 16"""
    import os
    import torch
    p = helpers.make_problem("vehicle", T=24, N=16384, M=1024, seed=6)
    cs = helpers.product_csmc(p, 0)
    dev = lambda x: torch.as_tensor(np.ascontiguousarray(x)).cuda()
    ref, Th, Sg = (dev(p[k][None]) for k in ("ref", "Theta", "Sigma"))
    key = helpers.pkg("random").key(3)
    os.environ.pop("PGAS_SWEEP_FUSED", None)
    a = cs.sweep(ref, Th, Sg, key=key)
    os.environ["PGAS_SWEEP_FUSED"] = "1"
    try:
        b = cs.sweep(ref, Th, Sg, key=key)
    finally:
        os.environ.pop("PGAS_SWEEP_FUSED", None)
    assert torch.equal(a["anc_trace"], b["anc_trace"]) and torch.equal(a["idx"], b["idx"])
    assert torch.allclose(a["state_trace"], b["state_trace"], rtol=1e-12, atol=1e-300)
    assert bool(torch.isfinite(a["traj"]).all())


@pytest.mark.parametrize("kind,N,T,M,chains", [("smo", 4096, 2000, 256, 2), ("vehicle", 16384, 5000, 1024, 1)])
def test_full_size_sweep_properties_and_teacher_forced_steps(built_lib, kind, N, T, M, chains):
    """BASELINE.json configs[3] (N = 4096 particles, T = 2000 steps, M = 256 basis functions) and configs[4] (N = 16384,
    T = 5000, M = 1024, two observations, slip-angle GP inputs) at FULL size, Philox mode.
    The oracle's Python loops cannot run 2000 steps of 4096 particles in test time, so the sweep is checked through
    (a) size-independent properties — ancestors sorted and in range, conditioned particle pinned to the reference,
    the returned trajectory equals reconstruct_trajectory of the traces — and (b) TEACHER-FORCED oracle steps at
    sampled depths: the log-weights entering step t are rebuilt from the GPU traces with the oracle's formulas
    (logw_{t-1} = log p(y_{t-1}|x_{t-1}) - l_aux_{t-1}[a_{t-2}], src/PGAS.py:137-147), then ONE oracle step with the
    restated Philox variates must reproduce the GPU's ancestors (exactly, CDF ties within 1e-12 excepted) and states
    (1e-9 relative) at that depth."""
    import torch
    from oracle import filtering as OF, pgas as OP, philox as OPH
    chain_base, it, seed = 5, 3, 0x5EEDBA5E
    p = helpers.make_problem(kind, T=T, N=N, M=M, seed=12)
    cs = helpers.product_csmc(p)
    dev = lambda x: torch.as_tensor(np.ascontiguousarray(x)).cuda()
    ref, Th, Sg = (dev(np.stack([p[k]] * chains)) for k in ("ref", "Theta", "Sigma"))
    out = cs.sweep(ref, Th, Sg, key=helpers.pkg("random").key(seed), chain_base=chain_base, iteration=it)
    st, an = out["state_trace"], out["anc_trace"]
    assert bool(torch.isfinite(st).all())
    # (a) properties
    assert bool((an[:, :, :-1] >= 0).all()) and bool((an[:, :, :-1] <= N - 1).all())
    assert bool((an[:, :, 1:-1] >= an[:, :, :-2]).all())                     # systematic resampling of sorted points
    assert bool((an[:, :, -1] >= 0).all()) and bool((an[:, :, -1] <= N).all())   # reference ancestor: unclipped searchsorted
    assert torch.equal(st[:, :, -1, :], ref)                                 # src/PGAS.py:134,194
    an_h, st_h = an.cpu().numpy(), st.cpu().numpy()
    for c in range(chains):
        want = OF.reconstruct_trajectory(st_h[c], np.clip(an_h[c], 0, N - 1), min(int(out["idx"][c]), N - 1))
        assert np.array_equal(out["traj"][c].cpu().numpy(), want.reshape(T, -1))
    # (b) teacher-forced oracle steps at sampled depths
    model, c = p["omodel"], chains - 1
    ii = np.arange(N)
    for t in (2, 3, 417, 1000, T - 1):
        x2, x1 = st_h[c, t - 2], st_h[c, t - 1]
        aux_prev = model.basis(x2, model.inputs[t - 1]) @ p["Theta"].T       # mu(x_{t-2}) (src/PGAS.py:45-57)
        la_prev = model.loglik(model.observations[t - 1], aux_prev, model.inputs[t - 1])
        a_prev = np.clip(an_h[c, t - 2], 0, N - 1)
        logw = model.loglik(model.observations[t - 1], x1, model.inputs[t - 1]) - la_prev[a_prev]
        za, zb = OPH.normal2(seed, OPH.PURPOSE_STATE, chain_base + c, it, np.full(N, t), ii)
        ua, ub = OPH.uniform2(seed, OPH.PURPOSE_STEP_U, chain_base + c, it, np.array([t]), 0)
        o = OP.csmc_step(model, t, logw, x1, p["Theta"], p["Sigma"], p["ref"][t], float(ua[0]), float(ub[0]), np.stack([za, zb], axis=1))
        r = helpers.compare_step(p, o, (o[0], st_h[c, t], an_h[c, t - 1]), float(ua[0]), float(ub[0]))   # log-weights are not traced
        assert r["ok"], (t, r)


@pytest.mark.parametrize("M,n,K,G", [(41, 1, 6, 70), (64, 2, 3, 300), (9, 1, 17, 5), (256, 2, 2, 33)])
def test_posterior_predictive_batch_matches_oracle(built_lib, M, n, K, G):
    """vmap(prior_mniw_2naturalPara_inv) over K statistics sets + prior_mniw_Predictive on a grid
    (src/BayesianInferrence.py:35-45, :64-89; SingleMassOscillator_Figures.py:58-89, :131-140), batched on the device."""
    from oracle import mniw as OM
    rng = np.random.default_rng(M + K)
    prior = OM.prior_mniw_2naturalPara(np.zeros((n, M)), np.diag(rng.uniform(0.5, 2.0, size=M)), np.eye(n), 3)
    e0, e1, e2, e3 = [], [], [], []
    for k in range(K):
        Phi, Y = rng.normal(size=(40 + 7 * k, M)), rng.normal(size=(40 + 7 * k, n))
        e0.append(prior[0] + Phi.T @ Y); e1.append(prior[1] + Phi.T @ Phi); e2.append(prior[2] + Y.T @ Y); e3.append(prior[3] + Phi.shape[0])
    basis = rng.normal(size=(G, M))
    r = helpers.pkg("BayesianInferrence").posterior_predictive_batch(np.stack(e0), np.stack(e1), np.stack(e2), np.array(e3, dtype=float), basis)
    assert int(r["status"].abs().sum()) == 0
    for k in range(K):
        mean, col_cov, row_scale, df = OM.prior_mniw_2naturalPara_inv(e0[k], e1[k], e2[k], e3[k])
        assert helpers.rel_err(r["mean"][k].cpu().numpy(), mean) < helpers.REL_TOL
        assert helpers.rel_err(r["row_scale"][k].cpu().numpy(), row_scale) < helpers.REL_TOL
        assert float(r["df"][k]) == df
        pm, pcs, prs, pdf = OM.prior_mniw_Predictive(mean, col_cov, row_scale, df, basis)
        assert helpers.rel_err(r["pred_mean"][k].cpu().numpy().reshape(np.shape(pm)), pm) < helpers.REL_TOL
        assert helpers.rel_err(r["pred_col_scale"][k].cpu().numpy(), np.diag(pcs)) < helpers.REL_TOL
        assert helpers.rel_err(r["pred_row_scale"][k].cpu().numpy(), prs) < helpers.REL_TOL
        assert float(r["pred_df"][k]) == pdf
