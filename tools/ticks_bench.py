"""phase clocks of the resampling kernel on the bench workload (developer aid): python tools/ticks_bench.py [iterations]"""
import ctypes as C, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
PKG = bench.PKG
L = importlib.import_module(PKG + "._lib"); lib = L.lib()
BF, MD, PG, BI, RND = (importlib.import_module(PKG + "." + n) for n in ("BasisFunctions", "models", "PGAS", "BayesianInferrence", "random"))
w = bench.workload(); T = bench.T_STEPS; N = bench.N_PART; count = 64
n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 3
hgp, sd = BF.generate_Hilbert_BasisFunction(bench.M_BASIS, w["domain"], w["lengthscale"], w["scale"])
prior = BI.prior_mniw_2naturalPara(np.zeros((2, bench.M_BASIS)), np.diag(sd), np.eye(2), w["df"])
pg = PG.PGAS(N_samples=N, N_iterations=2, observations=w["Y"][:T], inputs=np.zeros((T, 0)), init_state_mean=w["m0"], init_state_cov=w["P0"],
             likelihood_fcn=MD.gaussian_likelihood(lambda x: x[0], w["R"]), GP_prior=prior, basis_fcn=lambda state, inp: hgp(state))
m = pg.cSMC.model
ref = torch.as_tensor(np.broadcast_to(w["X"][:T], (count, T, 2)).copy()).cuda()
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
print("ms per iteration", e0.elapsed_time(e1) / n_it)
lib.pgas_debug_set_split_ticks(None)
d = dbg.cpu().numpy().reshape(64, 2, 8)
rows = int((d[:, 0, 0] > 0).sum()); d = d[:rows]
names = ["A loads+softmax", "sync1", "X1 folds + cluster barrier", "B1 cdf -> cluster", "barrier 2", "B2 + C"]
for th in (0, 1):
    seg = np.diff(d[2:, th, :7], axis=1)
    print("cluster rank", th, "cycles/step", np.median(np.diff(d[2:, th, 0])))
    for k, n in enumerate(names):
        print(f"   {n:28s} {np.median(seg[:, k]):8.0f}")
    print(f"   {'loop back':28s} {np.median(d[3:, th, 0] - d[2:-1, th, 6]):8.0f}")
