"""phase clocks of the resampling kernel on a bench workload (developer aid): python tools/ticks_bench.py [iterations] [chains] [config]"""
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

n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 3
count = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg = bench.CONFIGS[int(sys.argv[3]) if len(sys.argv) > 3 else 4]
N, M, T = cfg["N"], cfg["M"], cfg["T"]
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
ref = torch.as_tensor(np.broadcast_to(w["X"], (count, T, 2)).copy()).cuda()
nbytes = lib.pgas_run_chains_workspace_bytes(m.handle, N, count)
ws = torch.empty((nbytes,), dtype=torch.uint8, device="cuda")
p0, p1, p2 = pg._prior()
out = torch.empty((count, n_it + 1, T, 2), dtype=torch.float64, device="cuda")
dbg = torch.zeros((64 * 2 * 8,), dtype=torch.int64, device="cuda")
lib.pgas_debug_set_split_ticks.argtypes = [C.c_void_p]
lib.pgas_debug_set_split_ticks(L.ptr(dbg))
rng = PG._make_rng(RND.key(bench.SEED), 0, 0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
L.check(lib.pgas_run_chains_f64(m.handle, N, n_it + 1, count, L.ptr(p0), L.ptr(p1), L.ptr(p2), pg.GP_prior[3], L.ptr(ref), C.byref(rng),
                                L.ptr(out), C.c_void_p(0), C.c_void_p(0), 0, L.ptr(ws), nbytes, L.stream_ptr()))
e1.record()
torch.cuda.synchronize()
print("chains", count, "ms per iteration", e0.elapsed_time(e1) / n_it, {k: v for k, v in os.environ.items() if k.startswith("PGAS_")})
lib.pgas_debug_set_split_ticks(None)
d = dbg.cpu().numpy().reshape(64, 2, 8)
rows = int((d[:, 0, 0] > 0).sum())
d = d[:rows]
names = ["A loads+softmax", "wait 1", "X fold", "B cdf / scatter", "wait 2", "C new log-weights"]
for th in (0, 1):
    seg = np.diff(d[2:, th, :7], axis=1)
    print("CTA", th, "cycles/step", np.median(np.diff(d[2:, th, 0])))
    for k, n in enumerate(names):
        print(f"   {n:28s} {np.median(seg[:, k]):8.0f}")
    print(f"   {'loop back':28s} {np.median(d[3:, th, 0] - d[2:-1, th, 6]):8.0f}")
