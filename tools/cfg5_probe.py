"""BASELINE config 5 shape (vehicle-like 2-D basis M=1024, N=16384 particles) through the sweep API: split vs fused form.
usage: cfg5_probe.py [T] [chains] [cluster]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 301
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cl = int(sys.argv[3]) if len(sys.argv) > 3 else 0
N, M = 16384, 1024
p = helpers.make_problem("vehicle", T=T, N=N, M=M, seed=1)
cs = helpers.product_csmc(p, cl)
dev = lambda x: torch.as_tensor(np.ascontiguousarray(x)).cuda()
ref, Th, Sg = (dev(np.stack([p[k]] * nc)) for k in ("ref", "Theta", "Sigma"))
key = helpers.pkg("random").key(1)
res = {}
for mode in ("split", "fused"):
    if mode == "fused":
        os.environ["PGAS_SWEEP_FUSED"] = "1"
    else:
        os.environ.pop("PGAS_SWEEP_FUSED", None)
    out = cs.sweep(ref, Th, Sg, key=key)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = cs.sweep(ref, Th, Sg, key=key)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    ps = nc * N * (T - 1) / (best * 1e-3)
    res[mode] = out
    print(mode, dict(ms=best, us_per_step=1e3 * best / (T - 1), particle_steps_per_s=ps, alg_tflops=ps * (2 * M * 2 + M * 2) / 1e12,
                     finite=bool(torch.isfinite(out["traj"]).all())), flush=True)
print("ancestors equal:", bool(torch.equal(res["split"]["anc_trace"], res["fused"]["anc_trace"])),
      "states close:", bool(torch.allclose(res["split"]["state_trace"], res["fused"]["state_trace"], rtol=1e-12, atol=1e-300)))
