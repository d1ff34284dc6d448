#!/usr/bin/env python
"""bench.py — PGAS particle-steps/s on the BASELINE.json workloads
  --config 4 (default)  "scaled single-mass oscillator": N=4096 particles, T=2000 steps, M=256 basis functions, 64 chains per GPU
  --config 5            "scaled vehicle model": N=16384, T=5000, M=1024 (2-D tensor-product basis over the slip angles), 16 chains per GPU
The path shards over independent chains: weak scaling (default) runs the one-GPU workload on every GPU with its own chain
ids; chains never communicate, one NCCL all-gather of the per-chain trajectories at the end.  Multi-GPU runs add a
"strong" record: the one-GPU job (64 chains in total, BASELINE.json's named split) sharded over the ranks.

One "step" = one full Gibbs iteration of every chain of the job: conditional-SMC sweep (state kernel in 16-step launches
overlapped with the resampling kernel, ~290 launches) -> final pick + backward trace -> sufficient statistics -> MNIW draw.
One-GPU runs add "split_8gpu_share": one eighth of the chains on this GPU = the per-rank work of the named 8-GPU split.

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

SEED = 12345678                                          # the reference's seed (src/SingleMassOscillator.py:82)

# BASELINE.json configs[3] and configs[4].  `chains` = chains per GPU of the bench workload.
CONFIGS = {
    4: dict(kind="smo", N=4096, T=2000, M=256, chains=64, n_y=1,
            name="scaled single-mass oscillator (BASELINE.json configs[3])"),
    5: dict(kind="vehicle", N=16384, T=5000, M=1024, chains=16, n_y=2,
            name="scaled vehicle model, 2-D tensor-product basis over the slip angles (BASELINE.json configs[4])"),
}
STATE_BYTES_PER_PSTEP = 8 * 2 + 3 * 8                   # trace row (n_x doubles) + the three log-densities handed to the resampling kernel


def flop_per_pstep(M):
    return 2 * M * 2 + M * 2                             # 2 M n_x + M D  (SURVEY.md 8d), n_x = D = 2


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


def vehicle_truth(T, rng):
    """ground-truth single-track vehicle (src/Vehicle.py:17-128, 180-208): x = [yaw rate, lateral velocity],
    u = [steering angle, 11 m/s], magic-formula tyres, RK4 with dt = 0.02, steering profile of :200-208 stretched to T steps"""
    m, Izz, lf, lr, g, mux = 1720.0, 1827.5, 1.16, 1.47, 9.81, 0.9
    mu, B, Cc, E = 0.9, 10.0, 1.9, 0.97
    dt = 0.02
    time_ = np.arange(T) * dt
    t_end = T * dt
    U = np.zeros((T, 2))
    U[:, 0] = 10 / 180 * np.pi * np.sin(2 * np.pi * time_ / 5) * np.exp(-0.5 * (time_ - t_end / 2) ** 2 / (t_end / 5) ** 2)
    U[:, 1] = 11.0
    Fzf, Fzr = m * g * lr / (lf + lr), m * g * lf / (lf + lr)

    def mu_y(al):
        ta = np.tan(al)
        return mu * np.sin(Cc * np.arctan(B * (1 - E) * ta + E * np.arctan(B * ta)))

    def dx(x, u):
        af = u[0] - np.arctan((x[1] + x[0] * lf) / u[1])
        ar = -np.arctan((x[1] - x[0] * lr) / u[1])
        myf, myr = mu_y(af), mu_y(ar)
        ddpsi = (lf * Fzf * myf * np.cos(u[0]) - lr * Fzr * myr + lf * Fzf * mux * np.sin(u[0])) / Izz
        ay = (Fzf * myf * np.cos(u[0]) + Fzr * myr + Fzf * mux * np.sin(u[0])) / m - u[1] * x[0]
        return np.array([ddpsi, ay])
    X = np.zeros((T, 2))
    for t in range(1, T):
        x, u = X[t - 1], U[t - 1]
        k1 = dx(x, u); k2 = dx(x + dt / 2 * k1, u); k3 = dx(x + dt / 2 * k2, u); k4 = dx(x + dt * k3, u)
        X[t] = x + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4) + 1e-4 * rng.normal(size=2)
    R = np.diag([0.001 / 180 * np.pi, 1e-3])
    Y = X + rng.normal(size=(T, 2)) * np.sqrt(np.diag(R))
    return X, Y, U, R


def workload(cfg, T=None):
    """synthetic trajectories + model constants of one BASELINE configuration"""
    T = T or cfg["T"]
    rng = np.random.default_rng(SEED)
    M = cfg["M"]
    if cfg["kind"] == "smo":
        X, Y, F = smo_truth(max(T, 32), rng)
        return dict(kind="smo", X=X[:T], Y=Y[:T], U=np.zeros((T, 0)), domain=np.array([[-7.5, 7.5], [-7.5, 7.5]]), lengthscale=15.0 / M,
                    scale=100.0, hgp_extra=(), m0=np.zeros(2), P0=np.diag([1e-4, 1e-4]), H=np.array([[1.0, 0.0]]), R=np.array([[1e-3]]), df=3)
    X, Y, U, R = vehicle_truth(max(T, 32), rng)
    a = 30 / 180 * np.pi
    return dict(kind="vehicle", X=X[:T], Y=Y[:T], U=U[:T], domain=np.array([[-a, a], [-a, a]]), lengthscale=2 / 180 * np.pi, scale=50.0,
                hgp_extra=(), m0=np.zeros(2), P0=np.diag([1e-4, 1e-4]), H=np.eye(2), R=R, df=3, slip=(1.16, 1.47))


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
def _oracle_model(cfg, w, T):
    from oracle import basis as OB, mniw as OM, pgas as OP
    M = cfg["M"]
    hgp, sd = OB.generate_Hilbert_BasisFunction(M, w["domain"], w["lengthscale"], w["scale"])
    if w["kind"] == "vehicle":
        basis = OP.vehicle_slip_basis(hgp, *w["slip"])
    else:
        basis = OP.affine_hgp_basis(hgp, np.eye(2), np.zeros(2))
    model = OP.ThetaModel(w["Y"][:T], w["U"][:T], w["m0"], w["P0"], basis, OP.gaussian_loglik(w["H"], np.zeros(w["H"].shape[0]), w["R"]))
    prior = OM.prior_mniw_2naturalPara(np.zeros((2, M)), np.diag(sd), np.eye(2), w["df"])
    return model, prior


def _cpu_chain_sample(args):
    """one bounded sample of the workload on one core: a sweep of `T` steps at full N and M (the sweep is
    exactly linear in T) + statistics + draw, with the NumPy restatement of the reference"""
    config, T, seed = args
    from threadpoolctl import threadpool_limits
    from oracle import pgas as OP
    cfg = CONFIGS[config]
    N, M = cfg["N"], cfg["M"]
    w = workload(cfg, max(T, 32))
    model, prior = _oracle_model(cfg, w, T)
    rng = np.random.default_rng(seed)
    df = prior[3] + T - 1
    A, S, _ = OP.sample_params(model, prior, w["X"][:T], rng.chisquare(df - np.arange(2)), rng.normal(size=(2, 2)), rng.normal(size=(2, M)))
    Z, U = rng.normal(size=(T, N, 2)), rng.uniform(size=(T, 2))
    with threadpool_limits(limits=1):
        t0 = time.perf_counter()
        sw = OP.csmc_sweep(model, N, w["X"][:T], A, S, Z, U)
        OP.sample_params(model, prior, sw["traj"], rng.chisquare(df - np.arange(2)), rng.normal(size=(2, 2)), rng.normal(size=(2, M)))
        dt = time.perf_counter() - t0
    return N * (T - 1), dt


def cpu_baseline_single(config):
    cfg = CONFIGS[config]
    T_sample = 41 if config == 4 else 9
    _cpu_chain_sample((config, 5, 0))                      # warm-up (imports, page-in)
    ps, dt = _cpu_chain_sample((config, T_sample, 1))
    return {"value": ps / dt, "unit": "particle-steps/s", "cores": 1, "kind": "port",
            "sample": f"1 chain, N={cfg['N']}, M={cfg['M']}, {T_sample - 1} of {cfg['T'] - 1} steps (sweep is linear in T), NumPy restatement "
                      f"of the reference (oracle/pgas.py), single thread; jax/equinox not installable offline"}


def run_reference_arm(args):
    """the reference's CPU implementation of the path = the oracle port on all host cores, one chain per core"""
    import multiprocessing as mp
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    T_sample = 21 if args.config == 4 else 6
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for _ in range(max(args.warmup, 1) if args.warmup else 0):
            pool.map(_cpu_chain_sample, [(args.config, 5, 100 + c) for c in range(cores)])
        times, psteps = [], 0
        for k in range(args.steps):
            res = pool.map(_cpu_chain_sample, [(args.config, T_sample, 1000 * k + c) for c in range(cores)])
            times.append(max(r[1] for r in res))              # slowest chain of the step (setup excluded)
            psteps += sum(r[0] for r in res)
    total = sum(times)
    value = psteps / total
    sample = (f"{cores} chains in parallel (one per host core), N={cfg['N']}, M={cfg['M']}, {T_sample - 1} of {cfg['T'] - 1} steps per chain "
              f"per step, NumPy restatement of the reference (oracle/); jax 0.4.38 / equinox 0.12.2 not installable offline")
    line = {"impl": "reference", "metric": "pgas_particle_steps_per_s", "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{cfg['name']}: N={cfg['N']}, T={cfg['T']}, M={cfg['M']}, {cfg['chains']} chains per GPU", "sample": sample},
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
    # roofline of the group-B step: the reference does ~6 batched M x M factorisation-class operations per particle and
    # step (src/Algorithm3.py:95-106: two log base measures = Cholesky + LU each; src/Algorithm1.py:211-217, :249-256: two more
    # inverses) ~ 6 M^3 / 3 flop; this implementation updates two augmented factors by rank-one rotations and does three forward
    # solves, ~ 3 * 4 (M+1)^2 + 3 (M+1)^2 flop.  Both fractions are tiny by construction: a particle's step is one serial
    # instruction stream on one warp (profiles/r01_marg_sweep_summary.md) — the bound is dependent-issue latency, not a pipe.
    try:
        M0 = int(m.M[0])
        fl_ref, fl_impl = 6.0 * M0 ** 3 / 3.0, 15.0 * (M0 + 1) ** 2
        a_, b_ = C.c_double(), C.c_double()
        L_ = importlib.import_module(PKG + "._lib")
        L_.check(L_.lib().pgas_measure_fp64_peaks(C.byref(a_), C.byref(b_), L_.stream_ptr()))
        peak = max(a_.value, b_.value)
        rate = out["chains_7"]["particle_steps_per_s"]
        out["roofline"] = {"bound": "latency (one warp per particle, one warp per scheduler)", "peak": peak, "unit": "TFLOP/s",
                           "reference_algorithm_flop_per_particle_step": fl_ref, "achieved_vs_reference_algorithm": rate * fl_ref / 1e12,
                           "frac_vs_reference_algorithm": rate * fl_ref / 1e12 / peak,
                           "implementation_flop_per_particle_step": fl_impl, "achieved": rate * fl_impl / 1e12, "frac": rate * fl_impl / 1e12 / peak,
                           "note": "7 replicas; per-particle statistics and factors (8.6 MB per chain) stay in L2: DRAM traffic ~1.4 MB per 100 steps"}
    except Exception as e:
        out["roofline"] = {"error": str(e)}
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
def ncu_traffic(kernel_regex, capture):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the kernel, from the committed `ncu --set full` raw page of this
    round (profiles/<capture>); None when the capture does not hold that kernel.  Several matching launches: the largest.  The
    capture is named in the result — the number is an ncu measurement of the same kernel at the same launch geometry, not of this run."""
    import csv
    import re
    path = os.path.join(ROOT, "profiles", capture)
    best = None
    try:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            if re.search(kernel_regex, r[ik]):
                b = float(r[ir]) * scale.get(units[ir], 1.0) + float(r[iw]) * scale.get(units[iw], 1.0)
                if best is None or b > best["bytes_per_launch"]:
                    best = {"bytes_per_launch": b, "capture": "profiles/" + capture, "kernel": r[ik][:80]}
    except Exception:
        return None
    return best


def cuda_time(torch, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(); fn(); a1.record()
        torch.cuda.synchronize()
        ts.append(a0.elapsed_time(a1))
    return float(np.mean(ts))


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

    cfg = CONFIGS[args.config]
    M = cfg["M"]
    T = args.T or cfg["T"]
    N = args.particles or cfg["N"]
    chains_per_gpu = args.chains or cfg["chains"]
    w = workload(cfg, T)
    hgp, sd = BF.generate_Hilbert_BasisFunction(M, w["domain"], w["lengthscale"], w["scale"])
    prior = BI.prior_mniw_2naturalPara(np.zeros((2, M)), np.diag(sd), np.eye(2), w["df"])
    if w["kind"] == "vehicle":
        basis_fcn = MD.VehicleSlipBasis(hgp, *w["slip"])
        lik = MD.GaussianLikelihood(w["H"], np.zeros(2), w["R"])
    else:
        basis_fcn = lambda state, inp: hgp(state)                   # noqa: E731
        lik = MD.gaussian_likelihood(lambda x: x[0], w["R"])

    def make_pgas():
        return PG.PGAS(N_samples=N, N_iterations=2, observations=w["Y"], inputs=w["U"], init_state_mean=w["m0"], init_state_cov=w["P0"],
                       likelihood_fcn=lik, GP_prior=prior, basis_fcn=basis_fcn, cluster_size=args.cluster)
    pg = make_pgas()
    m = pg.cSMC.model
    key = RND.key(SEED)
    p0, p1, p2 = pg._prior()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_job(chains_total, steps, warmup):
        """`steps` Gibbs iterations of every chain of the job after `warmup` untimed ones; device-resident inputs; CUDA events on the
        launching stream, max over ranks; the only collective is the final gather of the trajectories"""
        first, count = DI.shard_chains(chains_total, rank, world)
        cnt = max(count, 1)
        ref = torch.as_tensor(np.broadcast_to(w["X"], (cnt, T, 2)).copy()).cuda()
        nbytes = lib.pgas_run_chains_workspace_bytes(m.handle, N, cnt)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device="cuda")

        def run_iterations(k_first, n_it, ref_t, out_t):
            rng = PG._make_rng(key, first, k_first)
            L.check(lib.pgas_run_chains_f64(m.handle, N, n_it + 1, cnt, L.ptr(p0), L.ptr(p1), L.ptr(p2), pg.GP_prior[3], L.ptr(ref_t),
                                            C.byref(rng), L.ptr(out_t), C.c_void_p(0), C.c_void_p(0), args.cluster, L.ptr(ws), nbytes,
                                            L.stream_ptr()))
        out_w = torch.empty((cnt, warmup + 1, T, 2), dtype=torch.float64, device="cuda")
        out_t = torch.empty((cnt, steps + 1, T, 2), dtype=torch.float64, device="cuda")
        run_iterations(0, warmup, ref, out_w)
        cur = out_w[:, warmup].contiguous()
        sync_all()
        l0 = lib.pgas_launch_count()
        with ClockSampler(local) as clk:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run_iterations(warmup, steps, cur, out_t)         # EXACTLY `steps` Gibbs iterations of every chain of this rank
            cur = out_t[:, steps].contiguous()
            DI.gather_chain_outputs(cur[:count], chains_total)
            e1.record()
            sync_all()
            ms = e0.elapsed_time(e1)
        launches = lib.pgas_launch_count() - l0
        t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        ms = float(t_ms[0])
        del ws, out_w, out_t, ref
        return dict(ms=ms, value=chains_total * N * (T - 1) * steps / (ms * 1e-3), launches=int(launches), clocks=clk.summary(),
                    first=first, count=count, cur=cur, chains_total=chains_total)

    chains_total = chains_per_gpu * world if args.scaling == "weak" else chains_per_gpu
    job = timed_job(chains_total, args.steps, args.warmup)
    ms, value, first, count, cur = job["ms"], job["value"], job["first"], job["count"], job["cur"]
    # the named split of BASELINE.json (configs[3]: 64 chains sharded over the GPUs of the box) beside the weak-scaling headline
    strong = None
    if world > 1 and args.scaling == "weak" and not args.no_strong:
        sj = timed_job(chains_per_gpu, args.steps, args.warmup)
        strong = {"chains_total": chains_per_gpu, "chains_per_gpu": chains_per_gpu / world, "value": sj["value"], "unit": "particle-steps/s",
                  "ms_per_step": sj["ms"] / args.steps, "note": "strong scaling: the one-GPU job (same chain ids, same results) sharded over the ranks"}

    # one GPU: this GPU's SHARE of the named 8-GPU split (configs[3]: 64 chains over 8 GPUs = 8 chains per GPU; chains never
    # communicate, so the share runs here exactly as it would on rank r of 8) -> projected strong-scaling efficiency
    share = None
    if world == 1 and args.scaling == "weak" and not args.no_strong and chains_per_gpu >= 8:
        sj = timed_job(chains_per_gpu // 8, args.steps, args.warmup)
        share = {"chains_on_this_gpu": chains_per_gpu // 8, "ms_per_step": sj["ms"] / args.steps, "value": sj["value"], "unit": "particle-steps/s",
                 "projected_value_8_gpus": 8 * sj["value"],
                 "projected_strong_scaling_efficiency_8_gpus": (ms / args.steps / 8) / (sj["ms"] / args.steps),
                 "note": f"one eighth of the one-GPU job ({chains_per_gpu // 8} of {chains_per_gpu} chains) on one GPU: what each rank of the named "
                         "8-GPU split executes (no data-path collective); efficiency = (one-GPU time / 8) / this time"}

    # ---- the kernels of one iteration alone (CUDA events), for the rooflines
    cnt = max(count, 1)
    st = torch.empty((cnt, T, N, 2), dtype=torch.float64, device="cuda")
    an = torch.empty((cnt, T - 1, N), dtype=torch.int32, device="cuda")
    lw = torch.empty((cnt, N), dtype=torch.float64, device="cuda")
    T0, T1, T2, T3 = BI.trajectory_statistics(m, cur)
    e0_, e1_, e2_ = (p0 + T0).contiguous(), (p1 + T1).contiguous(), (p2 + T2).contiguous()
    A0, S0, _ = BI.mniw_posterior_draw(e0_, e1_, e2_, pg.GP_prior[3] + T3, PG._make_rng(key, first, 999))
    sw_bytes = int(lib.pgas_csmc_sweep_workspace_bytes(m.handle, N, cnt))
    sw_ws = torch.empty((sw_bytes,), dtype=torch.uint8, device="cuda")
    it = [1000]

    def sweep_once():
        it[0] += 1
        rng = PG._make_rng(key, first, it[0])
        L.check(lib.pgas_csmc_sweep_f64(m.handle, N, cnt, L.ptr(cur), L.ptr(A0), L.ptr(S0), C.byref(rng), L.ptr(st), L.ptr(an),
                                        L.ptr(lw), C.c_void_p(0), C.c_void_p(0), args.cluster, L.ptr(sw_ws), sw_bytes, L.stream_ptr()))

    def state_once():
        it[0] += 1
        rng = PG._make_rng(key, first, it[0])
        L.check(lib.pgas_debug_state_kernel_f64(m.handle, N, cnt, L.ptr(cur), L.ptr(A0), L.ptr(S0), C.byref(rng), L.ptr(st), L.ptr(sw_ws),
                                                sw_bytes, L.stream_ptr()))
    sweep_avg = cuda_time(torch, sweep_once, 2)
    state_avg = cuda_time(torch, state_once, 2)
    suff_avg = cuda_time(torch, lambda: BI.trajectory_statistics(m, cur), 5)
    draw_avg = cuda_time(torch, lambda: BI.mniw_posterior_draw(e0_, e1_, e2_, pg.GP_prior[3] + T3, PG._make_rng(key, first, 998)), 5)
    dfma, dmma = C.c_double(), C.c_double()
    L.check(lib.pgas_measure_fp64_peaks(C.byref(dfma), C.byref(dmma), L.stream_ptr()))
    FLOP = flop_per_pstep(M)
    psteps_rank = cnt * N * (T - 1)
    flops_sweep = psteps_rank * FLOP
    achieved = flops_sweep / (state_avg * 1e-3) / 1e12
    achieved_sweep = flops_sweep / (sweep_avg * 1e-3) / 1e12
    peak = max(dfma.value, dmma.value)
    n_state_launches = ((T - 1 + 15) // 16) * (2 if cnt >= 4 else 1)
    hbm_bytes = psteps_rank * STATE_BYTES_PER_PSTEP
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        hbm_src = "fallback (B200_PROFILING.md)"
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    flops_suff = cnt * ((T - 1) * M * (M + 1) + 2 * (T - 1) * M * 2)
    flops_draw = cnt * (M ** 3 / 3.0 + 3 * 2 * M * M)          # what this implementation performs: one factorisation + three triangular solves with n_x columns
    # ncu captures of the same kernels at the same launch geometry (the state kernel's: one 16-step launch of one chain group of
    # configs[3]; the statistics': configs[3], 64 chains; the draw's: configs[4], 16 chains) — null for a configuration that was not captured
    tr_state = ncu_traffic(rf"csmc_state_kernel<2, {cfg['n_y']}, 0, 256, 2", "r02_state_kernel_raw.csv") if cnt * ((N + 511) // 512) >= 222 else None
    tr_suff = (ncu_traffic(r"suffstats_kernel", "r02_suffstats_kernel_raw.csv") if (args.config == 4 and cnt == 64) else
               ncu_traffic(r"suffstats_kernel", "r02_suffstats_kernel_cfg5_raw.csv") if (args.config == 5 and cnt == 16) else None)
    tr_draw = ncu_traffic(r"chol_update_kernel", "r02_tail_kernels_raw.csv") if args.config == 5 else None

    # ---- end-to-end leg: the public API with HOST buffers (pinned), H2D of the reference trajectories and D2H of the
    #      new trajectories inside the timed region, every step
    pg2 = make_pgas()
    pg2.cSMC._model = m
    host_ref = torch.as_tensor(np.broadcast_to(w["X"], (cnt, T, 2)).copy()).pin_memory()
    host_out = torch.empty((cnt, T, 2), dtype=torch.float64).pin_memory()
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step(k):
        dref = host_ref.cuda(non_blocking=True)
        res = pg2.run_chains(RND.key(SEED + k), dref, n_chains=cnt, chain_base=first, want_params=False)
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
    e2e_value = chains_total * N * (T - 1) * e2e_steps / float(t_e[0])

    if rank == 0:
        cpu = cpu_baseline_single(args.config) if (world == 1 and not args.no_cpu_baseline) else None
        pk = f"FP64 peak measured on this GPU in this run (register-resident DFMA {dfma.value:.1f}, DMMA {dmma.value:.1f} TFLOP/s; MEASURED_PEAKS.json has no FP64 figure)"
        line = {
            "metric": "pgas_particle_steps_per_s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{cfg['name']}: N={N} particles, T={T} steps, M={M} basis functions (2-D Hilbert GP), {chains_total} independent "
                                   f"chains in total, {count} on rank 0 ({args.scaling} scaling)",
                       "step": "one Gibbs iteration of every chain: cSMC sweep + pick/backward trace + sufficient statistics + MNIW draw",
                       "cluster_size": args.cluster, "l2": "per-step working set (state + ancestor traces, "
                       f"{cnt * T * N * 20 / 1e9:.1f} GB on rank 0) exceeds the 126 MB L2", "rng": "Philox-4x32-10 in-kernel"},
            "sweeps_per_s": chains_total * args.steps / (ms * 1e-3),
            "roofline": {"bound": "fp64", "kernel": "csmc_state_kernel (FP64 FMA row walk; timed alone over all T-1 steps, launched exactly as inside the sweep)",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "launches": n_state_launches, "flop_per_launch": flops_sweep / n_state_launches, "avg_launch_ms": state_avg / n_state_launches,
                         "traffic": tr_state["bytes_per_launch"] if tr_state else None,
                         "traffic_capture": tr_state["capture"] if tr_state else None,
                         "algorithmic_bytes_per_launch": hbm_bytes / n_state_launches,
                         "note": f"the bound is the FP64 pipe, not HBM and not the bf16 tensor pipe (tcgen05 has no f64 kind; on B200 FP64 FMA and FP64 DMMA share "
                                 f"one pipe: interleaved they reach the single-pipe rate, profiles/r02_microbench.md); algorithmic flops = {FLOP} per "
                                 f"particle-step (2 M n_x + M D) x {psteps_rank} particle-steps; peak = {pk}; state kernel {state_avg:.2f} ms; whole sweep "
                                 f"(state kernel overlapped with the resampling kernel) {sweep_avg:.2f} ms",
                         "sweep_ms": sweep_avg, "sweep_achieved": achieved_sweep, "sweep_frac": achieved_sweep / peak,
                         "hbm_achieved_gbs": hbm_bytes / (state_avg * 1e-3) / 1e9, "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src,
                         "hbm_frac": hbm_bytes / (state_avg * 1e-3) / 1e9 / hbm_peak},
            "roofline_kernels": {
                "sweep": {"bound": "fp64", "ms": sweep_avg, "achieved": achieved_sweep, "peak": peak, "unit": "TFLOP/s", "frac": achieved_sweep / peak},
                "suffstats": {"bound": "fp64", "kernel": "suffstats_kernel (SYRK over time, mma.sync.m8n8k4.f64)", "ms": suff_avg,
                              "achieved": flops_suff / (suff_avg * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                              "frac": flops_suff / (suff_avg * 1e-3) / 1e12 / peak, "flops": flops_suff,
                              "traffic": tr_suff["bytes_per_launch"] if tr_suff else None, "traffic_capture": tr_suff["capture"] if tr_suff else None},
                "mniw_draw": {"bound": "fp64 (latency-limited panels)", "kernel": "chol_prep/panel/update kernels + mniw_draw_kernel (one factorisation, M^3/3)",
                              "ms": draw_avg, "achieved": flops_draw / (draw_avg * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                              "frac": flops_draw / (draw_avg * 1e-3) / 1e12 / peak, "flops": flops_draw,
                              "traffic": tr_draw["bytes_per_launch"] if tr_draw else None, "traffic_capture": tr_draw["capture"] if tr_draw else None}},
            "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": int(cnt * T * 2 * 8),
                    "d2h_bytes_per_step": int(cnt * T * 2 * 8), "steps": e2e_steps,
                    "api": "PGAS.run_chains(key, host reference trajectories) -> host trajectories (pinned buffers)"},
            "gpu_launches": job["launches"], "clocks": job["clocks"],
        }
        if strong is not None:
            line["strong"] = strong
        if share is not None:
            line["split_8gpu_share"] = share
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if world == 1 and args.config == 4 and not args.no_marginalised:
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
    ap.add_argument("--config", type=int, default=4, choices=[4, 5], help="BASELINE.json configuration (1-based): 4 = scaled oscillator, 5 = scaled vehicle")
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (weak scaling) or in total (strong); default: the configuration's")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-strong", action="store_true", help="multi-GPU runs: skip the additional strong-scaling leg")
    ap.add_argument("--no-marginalised", action="store_true", help="skip the marginalised-path (Algorithm2/3) leg")
    ap.add_argument("--particles", type=int, default=0)
    ap.add_argument("--T", type=int, default=0)
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
