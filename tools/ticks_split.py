"""phase clocks of the GENERAL resampling kernel csmc_sweep_kernel<PRE> of the split sweep (developer aid; the dedicated kernels are
clocked by tools/ticks_bench.py): python tools/ticks_split.py kind N T M chains"""
import ctypes as C, sys, os
os.environ["PGAS_WEIGHTS_KERNEL"] = "0"
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
kind, N, T, M, nc = sys.argv[1], *map(int, sys.argv[2:6])
L = helpers.pkg("_lib"); lib = L.lib()
p = helpers.make_problem(kind, T=T, N=N, M=M, seed=1)
cs = helpers.product_csmc(p, 0)
dev = lambda x: torch.as_tensor(np.ascontiguousarray(x)).cuda()
ref = dev(np.stack([p["ref"]] * nc)); Th = dev(np.stack([p["Theta"]] * nc)); Sg = dev(np.stack([p["Sigma"]] * nc))
dbg = torch.zeros((64 * 2 * 8,), dtype=torch.int64, device="cuda")
lib.pgas_debug_set_split_ticks.argtypes = [C.c_void_p]
lib.pgas_debug_set_split_ticks(L.ptr(dbg))
for _ in range(2):
    out = cs.sweep(ref, Th, Sg, key=helpers.pkg("random").key(1))
torch.cuda.synchronize()
lib.pgas_debug_set_split_ticks(None)
d = dbg.cpu().numpy().reshape(64, 2, 8)
rows = int((d[:, 0, 0] > 0).sum())
d = d[:rows]
names = ["A loads+softmax4", "wait sync1", "X1 fold/cluster/fold", "wait sync2", "B1 cdf+count+sync", "B2 resample", "X2+C"]
for th in (0, 1):
    seg = np.diff(d[2:, th, :7], axis=1)
    print("CTA", th, "(thread 96) cycles/step", np.median(np.diff(d[2:, th, 0])))
    for k, n in enumerate(names[:6]):
        print(f"   {n:22s} {np.median(seg[:, k]):8.0f}")
    print(f"   {'X2 cluster barrier':22s} {np.median(d[2:, th, 7] - d[2:, th, 6]):8.0f}")
    print(f"   {'C + loop back':22s} {np.median(d[3:, th, 0] - d[2:-1, th, 7]):8.0f}")
