"""One Algorithm3 sweep of a synthetic problem, for ncu captures (developer tool).
usage: prof_marg.py kind T N M [n_chains]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import helpers  # noqa: E402
import helpers_marginal as HM  # noqa: E402

kind, T, N, M = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
nc = int(sys.argv[5]) if len(sys.argv) > 5 else 1
prob = HM.make_marg_problem(kind, T=T, N=N, M=M, seed=1)
A3 = helpers.pkg("Algorithm3").Algorithm3(**prob["prod_kwargs"])
key = helpers.pkg("random").key(3)
rng = np.random.default_rng(0)
Y = prob["prod_kwargs"]["observations"]
f64 = dict(dtype=torch.float64, device="cuda")
rx = torch.zeros((nc, T, 2), **f64)
rx[:, :, 0] = torch.as_tensor(np.atleast_2d(Y.reshape(T, -1))[:, 0], **f64)
rxi = torch.zeros((nc, prob["G"], T), **f64)
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = A3.csmc(rx, rxi, None, key=key, iteration=it)
    e1.record()
    torch.cuda.synchronize()
    print("ms", e0.elapsed_time(e1), "us/step", 1e3 * e0.elapsed_time(e1) / (T - 1), "status", r["status"].tolist(),
          "finite", bool(torch.isfinite(r["traj"]).all()))
