"""one sweep launch for ncu (developer aid): python tools/prof_sweep.py kind N T M chains cluster"""
import ctypes as C, sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
kind, N, T, M, nc, cl = sys.argv[1], *map(int, sys.argv[2:7])
L = helpers.pkg("_lib"); lib = L.lib()
p = helpers.make_problem(kind, T=T, N=N, M=M, seed=1)
cs = helpers.product_csmc(p, cl)
dev = lambda x: torch.as_tensor(np.ascontiguousarray(x)).cuda()
ref = dev(np.stack([p["ref"]] * nc)); Th = dev(np.stack([p["Theta"]] * nc)); Sg = dev(np.stack([p["Sigma"]] * nc))
out = cs.sweep(ref, Th, Sg, key=helpers.pkg("random").key(1))
torch.cuda.synchronize()
print("ok", bool(torch.isfinite(out["traj"]).all()))
