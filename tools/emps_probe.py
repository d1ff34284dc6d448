"""EMPS PGAS-baseline shape (3-D basis M = 729, N = 200 particles, one chain; src/EMPS.py:100-123, :240-255) through the sweep API,
per cluster size (developer aid): python tools/emps_probe.py [T] [chains]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
T = int(sys.argv[1]) if len(sys.argv) > 1 else 400
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 1
p = helpers.make_problem("emps", T=T, N=200, M=729, seed=1)
dev = lambda x: torch.as_tensor(np.ascontiguousarray(x)).cuda()
ref, Th, Sg = (dev(np.stack([p[k]] * nc)) for k in ("ref", "Theta", "Sigma"))
key = helpers.pkg("random").key(1)
for cl in (1, 2, 4, 8):
    cs = helpers.product_csmc(p, cl)
    out = cs.sweep(ref, Th, Sg, key=key)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = cs.sweep(ref, Th, Sg, key=key); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(dict(cluster=cl, ms=best, us_per_step=1e3 * best / (T - 1), finite=bool(torch.isfinite(out["traj"]).all())), flush=True)
