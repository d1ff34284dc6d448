"""Timing probe for the marginalised filters at the shipped configurations (developer tool)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "smo"
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    chains = [int(c) for c in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["1"])]
    import importlib
    mod = importlib.import_module("src." + {"smo": "SingleMassOscillator", "vehicle": "Vehicle", "emps": "EMPS"}[which])
    A1 = {"smo": "SMO_Algorithm1", "vehicle": "Vehicle_Algorithm1", "emps": "EMPS_Algorithm1"}[which]
    A2 = {"smo": "SMO_Algorithm2", "vehicle": "Vehicle_Algorithm2", "emps": "EMPS_Algorithm2"}[which]
    a1, a2 = getattr(mod, A1), getattr(mod, A2)
    from bayesian_inference_with_explicit_and_implicit_prior_knowledge_b200 import random as R
    key = R.key(1)
    t0 = time.time()
    m = a1.model
    print(json.dumps({"model_build_s": time.time() - t0, "T": m.T, "G": m.G, "M": m.M}))
    N, T = a1.N_samples, m.T
    ms = timed(lambda: a1.filter(key=key), reps=2)
    r = a1.filter(key=key)
    print(json.dumps({"alg": "Algorithm1", "ms": ms, "us_per_step": 1e3 * ms / (T - 1), "status": int(r["status"][0]),
                      "finite": bool(torch.isfinite(r["state_trace"]).all())}))
    x0 = r["state_trace"][0, :, 0].contiguous()
    xi0 = r["xi_trace"][0, :, :, 0].contiguous()
    for nc in chains:
        ix = x0[None].repeat(nc, 1, 1)
        ixi = xi0[None].repeat(nc, 1, 1)
        ms = timed(lambda: a2.run(ix, ixi, key=key, K=K + 1, want_sst=True), reps=2)
        rr = a2.run(ix, ixi, key=key, K=K + 1)
        per = ms / K
        print(json.dumps({"alg": "Algorithm2", "chains": nc, "K": K, "ms_per_sweep": per, "us_per_step": 1e3 * per / (T - 1),
                          "particle_steps_per_s": nc * N * (T - 1) / (per * 1e-3), "status": rr["status"].tolist(),
                          "finite": bool(torch.isfinite(rr["x_trace"]).all())}))


main()
