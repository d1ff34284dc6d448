"""timing of the batched posterior / predictive op at the shipped figure sizes (developer aid):
python tools/gp_post_probe.py [M] [n] [K] [G]   (defaults: single-mass oscillator, M=41, n=1, K=800 iterations, 50x50 grid)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
M, n, K, G = (int(v) for v in (sys.argv[1:5] + [41, 1, 800, 2500][len(sys.argv) - 1:]))
rng = np.random.default_rng(0)
Phi = rng.normal(size=(200, M))
e1 = np.broadcast_to(np.eye(M) + Phi.T @ Phi, (K, M, M)).copy()
e0 = rng.normal(size=(K, M, n)); e2 = np.broadcast_to(50.0 * np.eye(n), (K, n, n)).copy(); e3 = np.full(K, 10.0)
basis = rng.normal(size=(G, M))
BI = helpers.pkg("BayesianInferrence")
d = [torch.as_tensor(x, device="cuda") for x in (e0, e1, e2, e3, basis)]
BI.posterior_predictive_batch(*d)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); r = BI.posterior_predictive_batch(*d); t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1)
flops = K * (M ** 3 / 3 + 2 * M * M * n + G * (M * M + 2 * M * n))
print(dict(M=M, n=n, K=K, G=G, ms=ms, gflops=flops / ms / 1e6, status=int(r["status"].abs().sum())))
