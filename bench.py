#!/usr/bin/env python
"""bench.py — PGAS particle-steps/s on the BASELINE.json workload "scaled single-mass oscillator":
N=4096 particles, T=2000 steps, M=256 basis functions (2-D Hilbert GP), 64 independent chains per GPU
(weak scaling: the path shards over independent chains, every GPU runs the one-GPU workload on its own
chain ids; chains never communicate, one NCCL all-gather of the per-chain trajectories at the end;
`--scaling strong` shards 64 chains in total instead).

One "step" = one full Gibbs iteration of every chain of the job: conditional-SMC sweep (persistent
kernel) -> final pick + backward trace -> sufficient statistics -> MNIW posterior draw.

  python bench.py --gpus 1 --steps K --warmup W          # this repo (CUDA, sm_100a)
  python bench.py --impl reference ...                    # CPU restatement of the reference (oracle port:
                                                          # JAX/equinox are not installable here)
Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "bayesian_inference_with_explicit_and_implicit_prior_knowledge_b200"

N_PART, T_STEPS, M_BASIS, CHAINS_TOTAL = 4096, 2000, 256, 64
FLOP_PER_PSTEP = 2 * M_BASIS * 2 + M_BASIS * 2          # 2 M n_x + M D  (SURVEY.md 8d), n_x = D = 2
NCU_DRAM_BYTES_PER_PSTEP = (0.267264e6 + 32.3328e6) / (32 * 4096 * 16)        # profiles/r01_state_kernel_raw.csv (one 16-step launch of one chain group)
STATE_BYTES_PER_PSTEP = 8 * 2 + 3 * 8                   # trace row (n_x doubles) + the three log-densities handed to the resampling kernel
SEED = 12345678                                          # the reference's seed (src/SingleMassOscillator.py:82)


# ------------------------------------------------------------------------------- workload
def smo_truth(T, rng):
    """ground-truth single-mass oscillator (src/SingleMassOscillator.py:17-48, 85-97): RK4, dt = 0.02,
    three-level external force, y = x[0] + N(0, 1e-3)"""
    m, c1, c2, d1, d2, dt = 0.2, 5.0, 2.0, 0.4, 0.4, 0.02
    F_ext = np.ones(T) * 9.81 * m
    F_ext[T // 3:] = 0.0
    F_ext[2 * T // 3:] = -9.81 * m

    def dx(x, F):
        f_sd = c1 * x[0] + c2 * x[0] ** 3 + d1 * x[1] / (1 + d2 * x[1] * np.tanh(x[1]))
        return np.array([x[1], (-f_sd + F) / m])
    X = np.zeros((T, 2))
    Q = np.sqrt(np.array([5e-8, 5e-9]))
    for t in range(1, T):
        x, F = X[t - 1], F_ext[t - 1]
        k1 = dx(x, F); k2 = dx(x + dt / 2 * k1, F); k3 = dx(x + dt / 2 * k2, F); k4 = dx(x + dt * k3, F)
        X[t] = x + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4) + Q * rng.normal(size=2)
    Y = X[:, 0] + np.sqrt(1e-3) * rng.normal(size=T)
    return X, Y, F_ext


def workload(T=T_STEPS):
    rng = np.random.default_rng(SEED)
    X, Y, F = smo_truth(T, rng)
    return dict(X=X, Y=Y, F=F, domain=np.array([[-7.5, 7.5], [-7.5, 7.5]]), lengthscale=15.0 / M_BASIS, scale=100.0,
                m0=np.zeros(2), P0=np.diag([1e-4, 1e-4]), R=np.array([[1e-3]]), df=3)


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop = index, [], threading.Event()
        self._th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------- CPU baseline
def _oracle_model(w, T):
    from oracle import basis as OB, mniw as OM, pgas as OP
    hgp, sd = OB.generate_Hilbert_BasisFunction(M_BASIS, w["domain"], w["lengthscale"], w["scale"])
    model = OP.ThetaModel(w["Y"][:T], np.zeros((T, 0)), w["m0"], w["P0"], OP.affine_hgp_basis(hgp, np.eye(2), np.zeros(2)),
                          OP.gaussian_loglik([[1.0, 0.0]], [0.0], w["R"]))
    prior = OM.prior_mniw_2naturalPara(np.zeros((2, M_BASIS)), np.diag(sd), np.eye(2), w["df"])
    return model, prior


def _cpu_chain_sample(args):
    """one bounded sample of the workload on one core: a sweep of `T` steps at full N and M (the sweep is
    exactly linear in T) + statistics + draw, with the NumPy restatement of the reference"""
    T, seed = args
    from threadpoolctl import threadpool_limits
    from oracle import pgas as OP
    w = workload(max(T, 32))
    model, prior = _oracle_model(w, T)
    rng = np.random.default_rng(seed)
    df = prior[3] + T - 1
    A, S, _ = OP.sample_params(model, prior, w["X"][:T], rng.chisquare(df - np.arange(2)), rng.normal(size=(2, 2)), rng.normal(size=(2, M_BASIS)))
    Z, U = rng.normal(size=(T, N_PART, 2)), rng.uniform(size=(T, 2))
    with threadpool_limits(limits=1):
        t0 = time.perf_counter()
        sw = OP.csmc_sweep(model, N_PART, w["X"][:T], A, S, Z, U)
        OP.sample_params(model, prior, sw["traj"], rng.chisquare(df - np.arange(2)), rng.normal(size=(2, 2)), rng.normal(size=(2, M_BASIS)))
        dt = time.perf_counter() - t0
    return N_PART * (T - 1), dt


def cpu_baseline_single(T_sample=41):
    _cpu_chain_sample((9, 0))                              # warm-up (imports, page-in)
    ps, dt = _cpu_chain_sample((T_sample, 1))
    return {"value": ps / dt, "unit": "particle-steps/s", "cores": 1, "kind": "port",
            "sample": f"1 chain, N={N_PART}, M={M_BASIS}, {T_sample - 1} of {T_STEPS - 1} steps (sweep is linear in T), NumPy restatement "
                      f"of the reference (oracle/pgas.py), single thread; jax/equinox not installable offline"}


def run_reference_arm(args):
    """the reference's CPU implementation of the path = the oracle port on all host cores, one chain per core"""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    T_sample = 21
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for _ in range(max(args.warmup, 1) if args.warmup else 0):
            pool.map(_cpu_chain_sample, [(9, 100 + c) for c in range(cores)])
        times, psteps = [], 0
        for k in range(args.steps):
            res = pool.map(_cpu_chain_sample, [(T_sample, 1000 * k + c) for c in range(cores)])
            times.append(max(r[1] for r in res))              # slowest chain of the step (setup excluded)
            psteps += sum(r[0] for r in res)
    total = sum(times)
    value = psteps / total
    sample = (f"{cores} chains in parallel (one per host core), N={N_PART}, M={M_BASIS}, {T_sample - 1} of {T_STEPS - 1} steps per chain "
              f"per step, NumPy restatement of the reference (oracle/); jax 0.4.38 / equinox 0.12.2 not installable offline")
    line = {"impl": "reference", "metric": "pgas_particle_steps_per_s", "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "scaled single-mass oscillator: N=4096, T=2000, M=256, 64 chains per GPU (BASELINE.json configs[3])",
                       "sample": sample},
            "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- marginalised path (group B)
def marginalised_leg(key_mod):
    """Shipped single-mass-oscillator configuration (BASELINE.json configs[0]: N=200, T=750, M=41) through the
    marginalised PGAS (Algorithm2 -> Algorithm3): conditional sweeps per second for 1 chain and for 7 replicas
    (7 clusters of 16 CTAs are co-resident on B200, profiles/r01_microbench.md), beside the NumPy restatement of one Algorithm3 step on one host core."""
    import importlib
    import torch
    S = importlib.import_module("src.SingleMassOscillator")
    a1, a2 = S.SMO_Algorithm1, S.SMO_Algorithm2
    m = a1.model
    N, T = a1.N_samples, m.T
    key = key_mod.key(SEED)
    r = a1.filter(key=key)
    x0, xi0 = r["state_trace"][0, :, 0].contiguous(), r["xi_trace"][0, :, :, 0].contiguous()
    out = {"workload": f"SingleMassOscillator as shipped: N={N}, T={T}, M={m.M[0]}, Algorithm2/3 (marginalised PGAS)"}
    for nc in (1, 7):
        ix, ixi = x0[None].repeat(nc, 1, 1), xi0[None].repeat(nc, 1, 1)
        a2.run(ix, ixi, key=key, K=2, want_sst=True)
        torch.cuda.synchronize()
        K = 4
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rr = a2.run(ix, ixi, key=key, K=K + 1, want_sst=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        assert bool(torch.isfinite(rr["x_trace"]).all()) and int(rr["status"].abs().sum()) == 0
        out[f"chains_{nc}"] = {"ms_per_sweep": ms, "us_per_step": 1e3 * ms / (T - 1), "sweeps_per_s": nc / (ms * 1e-3),
                               "particle_steps_per_s": nc * N * (T - 1) / (ms * 1e-3)}
    # CPU: the oracle's Algorithm3 step (NumPy restatement, one core), a few steps at the same N and M
    try:
        from threadpoolctl import threadpool_limits
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import helpers_marginal as HM
        from oracle import marginal as OMg
        steps = 4
        prob = HM.make_marg_problem("smo", T=steps + 1, N=N, M=m.M[0], seed=1)
        V = HM.make_variates(prob, 1.0, seed=2)
        ref_x, ref_xi = np.zeros((steps + 1, 2)), [np.zeros(steps + 1)]
        rs = OMg.reference_stats(prob["oracle"], ref_x, ref_xi)
        with threadpool_limits(limits=1):
            t0 = time.perf_counter()
            OMg.alg3_run(prob["oracle"], N, ref_x, ref_xi, rs, HM.oracle_variates(V))
            dt = time.perf_counter() - t0
        out["cpu_port"] = {"particle_steps_per_s": N * steps / dt, "cores": 1, "kind": "port",
                           "sample": f"{steps} Algorithm3 steps at N={N}, M={m.M[0]} (oracle/marginal.py, single thread)"}
    except Exception as e:          # the oracle is test infrastructure; the GPU numbers stand without it
        out["cpu_port"] = {"error": str(e)}
    return out


# ------------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import importlib
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the PGAS hot path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    L = importlib.import_module(PKG + "._lib")
    lib = L.lib()
    BF, MD, PG, DI, BI = (importlib.import_module(PKG + "." + n) for n in ("BasisFunctions", "models", "PGAS", "distributed", "BayesianInferrence"))
    RND = importlib.import_module(PKG + ".random")

    w = workload()
    T = args.T
    hgp, sd = BF.generate_Hilbert_BasisFunction(M_BASIS, w["domain"], w["lengthscale"], w["scale"])
    prior = BI.prior_mniw_2naturalPara(np.zeros((2, M_BASIS)), np.diag(sd), np.eye(2), w["df"])
    if args.scaling == "weak":
        args.chains = args.chains * world                    # per-GPU work fixed: 64 chains on every rank
    first, count = DI.shard_chains(args.chains, rank, world)
    K = args.steps + args.warmup + 1                         # iteration 0 is the initial draw
    del K
    pg = PG.PGAS(N_samples=args.particles, N_iterations=2, observations=w["Y"][:T], inputs=np.zeros((T, 0)), init_state_mean=w["m0"],
                 init_state_cov=w["P0"], likelihood_fcn=MD.gaussian_likelihood(lambda x: x[0], w["R"]), GP_prior=prior,
                 basis_fcn=lambda state, inp: hgp(state), cluster_size=args.cluster)
    m = pg.cSMC.model
    key = RND.key(SEED)
    N = args.particles

    # ---- device-resident leg: K iterations, timed per iteration with CUDA events on the launching stream
    ref = torch.as_tensor(np.broadcast_to(w["X"][:T], (count, T, 2)).copy()).cuda()
    nbytes = lib.pgas_run_chains_workspace_bytes(m.handle, N, count)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device="cuda")
    p0, p1, p2 = pg._prior()

    def run_iterations(k_first, n_it, ref_t, out_t):
        """iterations k_first .. k_first+n_it-1 continuing from ref_t; out_t (count, n_it+1, T, 2), row 0 = ref_t"""
        rng = PG._make_rng(key, first, k_first)
        L.check(lib.pgas_run_chains_f64(m.handle, N, n_it + 1, count, L.ptr(p0), L.ptr(p1), L.ptr(p2), pg.GP_prior[3], L.ptr(ref_t),
                                        C.byref(rng), L.ptr(out_t), C.c_void_p(0), C.c_void_p(0), args.cluster, L.ptr(ws), nbytes,
                                        L.stream_ptr()))

    # sweep-kernel-only timing for the roofline (same launches as inside the loop)
    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    W = args.warmup
    out_w = torch.empty((count, W + 1, T, 2), dtype=torch.float64, device="cuda")
    out_t = torch.empty((count, args.steps + 1, T, 2), dtype=torch.float64, device="cuda")
    run_iterations(0, W, ref, out_w)                          # warm-up: W full iterations
    cur = out_w[:, W].contiguous()
    sync_all()
    launches0 = lib.pgas_launch_count()
    with ClockSampler(local) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_iterations(W, args.steps, cur, out_t)             # EXACTLY `steps` Gibbs iterations of every chain of this rank
        cur = out_t[:, args.steps].contiguous()
        gathered = DI.gather_chain_outputs(cur, args.chains)  # the only collective: final gather of the trajectories
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
    launches = lib.pgas_launch_count() - launches0
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms[0])
    psteps_job = args.chains * N * (T - 1) * args.steps
    value = psteps_job / (ms * 1e-3)

    # ---- sweep kernel alone (the dominant kernel), CUDA events, for the roofline
    st = torch.empty((count, T, N, 2), dtype=torch.float64, device="cuda")
    an = torch.empty((count, T - 1, N), dtype=torch.int32, device="cuda")
    lw = torch.empty((count, N), dtype=torch.float64, device="cuda")
    T0, T1, T2, T3 = BI.trajectory_statistics(m, cur)
    A0, S0, _ = BI.mniw_posterior_draw(p0 + T0, p1 + T1, p2 + T2, pg.GP_prior[3] + T3, PG._make_rng(key, first, 999))
    sw_bytes = int(lib.pgas_csmc_sweep_workspace_bytes(m.handle, N, count))
    sw_ws = torch.empty((sw_bytes,), dtype=torch.uint8, device="cuda")
    sweep_ms = []
    for r in range(3):
        rng = PG._make_rng(key, first, 1000 + r)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        L.check(lib.pgas_csmc_sweep_f64(m.handle, N, count, L.ptr(cur), L.ptr(A0), L.ptr(S0), C.byref(rng), L.ptr(st), L.ptr(an),
                                        L.ptr(lw), C.c_void_p(0), C.c_void_p(0), args.cluster, L.ptr(sw_ws), sw_bytes, L.stream_ptr()))
        a1.record()
        torch.cuda.synchronize()
        sweep_ms.append(a0.elapsed_time(a1))
    sweep_avg = float(np.mean(sweep_ms[1:]))
    # the dominant kernel alone: csmc_state_kernel over all T-1 steps (same launches as inside the sweep, no resampling kernel)
    state_ms = []
    for r in range(3):
        rng = PG._make_rng(key, first, 2000 + r)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        L.check(lib.pgas_debug_state_kernel_f64(m.handle, N, count, L.ptr(cur), L.ptr(A0), L.ptr(S0), C.byref(rng), L.ptr(st), L.ptr(sw_ws),
                                                sw_bytes, L.stream_ptr()))
        a1.record()
        torch.cuda.synchronize()
        state_ms.append(a0.elapsed_time(a1))
    state_avg = float(np.mean(state_ms[1:]))
    dfma, dmma = C.c_double(), C.c_double()
    L.check(lib.pgas_measure_fp64_peaks(C.byref(dfma), C.byref(dmma), L.stream_ptr()))
    flops_launch = count * N * (T - 1) * FLOP_PER_PSTEP
    achieved = flops_launch / (state_avg * 1e-3) / 1e12
    achieved_sweep = flops_launch / (sweep_avg * 1e-3) / 1e12
    peak = max(dfma.value, dmma.value)
    hbm_bytes = count * N * (T - 1) * STATE_BYTES_PER_PSTEP
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)

    # ---- end-to-end leg: the public API with HOST buffers (pinned), H2D of the reference trajectories and D2H of the
    #      new trajectories inside the timed region, every step
    pg2 = PG.PGAS(N_samples=N, N_iterations=2, observations=w["Y"][:T], inputs=np.zeros((T, 0)), init_state_mean=w["m0"],
                  init_state_cov=w["P0"], likelihood_fcn=MD.gaussian_likelihood(lambda x: x[0], w["R"]), GP_prior=prior,
                  basis_fcn=lambda state, inp: hgp(state), cluster_size=args.cluster)
    pg2.cSMC._model = m
    host_ref = torch.as_tensor(np.broadcast_to(w["X"][:T], (count, T, 2)).copy()).pin_memory()
    host_out = torch.empty((count, T, 2), dtype=torch.float64).pin_memory()
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step(k):
        dref = host_ref.cuda(non_blocking=True)
        res = pg2.run_chains(RND.key(SEED + k), dref, n_chains=count, chain_base=first, want_params=False)
        host_out.copy_(res["state_trace"][:, 1], non_blocking=True)
        torch.cuda.synchronize()
        host_ref.copy_(host_out)
    e2e_step(0)
    sync_all()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(1 + k)
    sync_all()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = args.chains * N * (T - 1) * e2e_steps / float(t_e[0])

    if rank == 0:
        cpu = cpu_baseline_single() if (world == 1 and not args.no_cpu_baseline) else None
        line = {
            "metric": "pgas_particle_steps_per_s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"scaled single-mass oscillator (BASELINE.json configs[3]): N={N} particles, T={T} steps, M={M_BASIS} basis "
                                   f"functions (2-D Hilbert GP), {args.chains} independent chains in total, {count} on rank 0 ({args.scaling} scaling)",
                       "step": "one Gibbs iteration of every chain: cSMC sweep + pick/backward trace + sufficient statistics + MNIW draw",
                       "cluster_size": args.cluster, "l2": "per-step working set (state + ancestor traces, "
                       f"{count * T * N * 20 / 1e9:.1f} GB on rank 0) exceeds the 126 MB L2", "rng": "Philox-4x32-10 in-kernel"},
            "sweeps_per_s": args.chains * args.steps / (ms * 1e-3),
            "roofline": {"bound": "tensor", "kernel": "csmc_state_kernel (FP64 FMA row walk; timed alone over all T-1 steps, "
                         f"{(T - 1 + 15) // 16} launches of <= 16 steps per chain group, two chain groups on two streams exactly as inside the sweep)", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_PSTEP * count * N * (T - 1),
                         "traffic_note": "dram__bytes_read+write of this kernel from the ncu --set full capture in profiles/r01_state_kernel_summary.md "
                                         f"({NCU_DRAM_BYTES_PER_PSTEP:.1f} B per particle-step measured on one 16-step launch, scaled to all launches; algorithmic {STATE_BYTES_PER_PSTEP} B: "
                                         "most of the rows a 16-step launch writes are still in L2 when the capture ends)",
                         "note": f"compute bound = FP64 pipe (on B200 the FP64 FMA and FP64 tensor (DMMA) pipes have the same measured rate); algorithmic flops = "
                                 f"{FLOP_PER_PSTEP} per particle-step (2 M n_x + M D) x {count * N * (T - 1)} particle-steps; peak = FP64 measured on this GPU in this "
                                 f"run (register-resident DFMA {dfma.value:.1f}, DMMA {dmma.value:.1f} TFLOP/s; MEASURED_PEAKS.json has no FP64 figure); state kernel "
                                 f"{state_avg:.2f} ms; whole sweep (state kernel overlapped with the resampling kernel csmc_weights1_kernel) {sweep_avg:.2f} ms",
                         "sweep_ms": sweep_avg, "sweep_achieved": achieved_sweep, "sweep_frac": achieved_sweep / peak,
                         "hbm_achieved_gbs": hbm_bytes / (state_avg * 1e-3) / 1e9, "hbm_peak_gbs": hbm_peak,
                         "hbm_frac": hbm_bytes / (state_avg * 1e-3) / 1e9 / hbm_peak},
            "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": int(count * T * 2 * 8),
                    "d2h_bytes_per_step": int(count * T * 2 * 8), "steps": e2e_steps,
                    "api": "PGAS.run_chains(key, host reference trajectories) -> host trajectories (pinned buffers)"},
            "gpu_launches": int(launches), "clocks": clk.summary(),
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if world == 1 and not args.no_marginalised:
            line["marginalised"] = marginalised_leg(RND)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chains", type=int, default=CHAINS_TOTAL, help="chains per GPU (weak scaling) or in total (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-marginalised", action="store_true", help="skip the marginalised-path (Algorithm2/3) leg")
    ap.add_argument("--particles", type=int, default=N_PART)
    ap.add_argument("--T", type=int, default=T_STEPS)
    ap.add_argument("--cluster", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            run_reference_arm(args)
        return
    run_gpu_arm(args)


if __name__ == "__main__":
    main()
