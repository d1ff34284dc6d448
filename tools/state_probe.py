"""Developer aid: the state kernel alone (bench workload, configs[3]) with in-kernel Philox noise vs pre-generated noise read
from global memory, and the noise generator alone.  usage: state_probe.py [chains] [T]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import helpers  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 401
cfgn = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cfg = bench.CONFIGS[cfgn]
N, M = cfg["N"], cfg["M"]
L, BF, MD, PG, BI, RND = (helpers.pkg(n) for n in ("_lib", "BasisFunctions", "models", "PGAS", "BayesianInferrence", "random"))
lib = L.lib()
w = bench.workload(cfg, T)
hgp, sd = BF.generate_Hilbert_BasisFunction(M, w["domain"], w["lengthscale"], w["scale"])
prior = BI.prior_mniw_2naturalPara(np.zeros((2, M)), np.diag(sd), np.eye(2), w["df"])
if w["kind"] == "vehicle":
    basis_fcn, lik = MD.VehicleSlipBasis(hgp, *w["slip"]), MD.GaussianLikelihood(w["H"], np.zeros(2), w["R"])
else:
    basis_fcn, lik = (lambda s, u: hgp(s)), MD.gaussian_likelihood(lambda x: x[0], w["R"])
pg = PG.PGAS(N_samples=N, N_iterations=2, observations=w["Y"], inputs=w["U"], init_state_mean=w["m0"], init_state_cov=w["P0"],
             likelihood_fcn=lik, GP_prior=prior, basis_fcn=basis_fcn)
m = pg.cSMC.model
key = RND.key(1)
cur = torch.as_tensor(np.broadcast_to(w["X"], (nc, T, 2)).copy()).cuda()
p0, p1, p2 = pg._prior()
T0, T1, T2, T3 = BI.trajectory_statistics(m, cur)
A0, S0, _ = BI.mniw_posterior_draw(p0 + T0, p1 + T1, p2 + T2, pg.GP_prior[3] + T3, PG._make_rng(key, 0, 999))
st = torch.empty((nc, T, N, 2), dtype=torch.float64, device="cuda")
sw_bytes = int(lib.pgas_csmc_sweep_workspace_bytes(m.handle, N, nc))
sw_ws = torch.empty((sw_bytes,), dtype=torch.uint8, device="cuda")
dfma, dmma = C.c_double(), C.c_double()
L.check(lib.pgas_measure_fp64_peaks(C.byref(dfma), C.byref(dmma), L.stream_ptr()))
peak = max(dfma.value, dmma.value)
flops = nc * N * (T - 1) * bench.flop_per_pstep(M)


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def state(rng):
    L.check(lib.pgas_debug_state_kernel_f64(m.handle, N, nc, L.ptr(cur), L.ptr(A0), L.ptr(S0), C.byref(rng), L.ptr(st), L.ptr(sw_ws),
                                            sw_bytes, L.stream_ptr()))


rng_p = PG._make_rng(key, 0, 5)
ms_p = timeit(lambda: state(rng_p))
ref_trace = st.clone()
Z = torch.empty((nc, T, N, 2), dtype=torch.float64, device="cuda")
U = torch.empty((nc, T, 2), dtype=torch.float64, device="cuda")
ms_z = timeit(lambda: L.check(lib.pgas_philox_sweep_variates_f64(C.byref(rng_p), nc, T, N, 2, L.ptr(Z), L.ptr(U), L.stream_ptr())))
rng_i = PG._make_rng(None, 0, 5, variates=dict(Z=Z, U=U))
ms_i = timeit(lambda: state(rng_i))
same = bool(torch.equal(st, ref_trace))
print(dict(config=cfgn, chains=nc, T=T, peak=peak, state_philox_ms=ms_p, frac_philox=flops / ms_p / 1e9 / peak, state_injected_ms=ms_i,
           frac_injected=flops / ms_i / 1e9 / peak, variates_kernel_ms=ms_z, same_trace=same,
           env={k: v for k, v in os.environ.items() if k.startswith("PGAS_")}), flush=True)
