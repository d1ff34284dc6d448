"""GPU probe (developer aid): FP64 microbenchmarks and sweep timings at the BASELINE shapes."""
import ctypes as C
import json
import sys
import os
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402

g.build()
import helpers  # noqa: E402

L = helpers.pkg("_lib")
lib = L.lib()
out = {}
a, b = C.c_double(), C.c_double()
L.check(lib.pgas_measure_fp64_peaks(C.byref(a), C.byref(b), L.stream_ptr()))
out["dfma_tflops"], out["dmma_tflops"] = a.value, b.value
lib.pgas_microbench_f64.restype = C.c_int
lib.pgas_microbench_f64.argtypes = [C.POINTER(C.c_double), C.c_void_p]
arr = (C.c_double * 8)()
L.check(lib.pgas_microbench_f64(arr, L.stream_ptr()))
out["dfma_latency_cyc"], out["dfma_issue_cyc"], out["cluster16_barrier_cyc"], out["cluster8_barrier_cyc"], out["dfma_2w_ilp4_tflops"], out["dmma_latency_cyc"], out["dmma_1w_ilp5_cyc"], out["dmma_16w_ilp5_cyc"] = list(arr)
lib.pgas_microbench_mix_f64.restype = C.c_int
lib.pgas_microbench_mix_f64.argtypes = [C.POINTER(C.c_double), C.c_void_p]
arr4 = (C.c_double * 4)()
L.check(lib.pgas_microbench_mix_f64(arr4, L.stream_ptr()))
out["mix_dfma8_tflops"], out["mix_dmma8_tflops"], out["mix_dfma8_dmma8_tflops"], out["mix_dfma8_dmma2_tflops"] = list(arr4)
print(json.dumps(out), flush=True)
if len(sys.argv) > 1: sys.exit(0)


def time_sweep(kind, N, T, M, n_chains, cluster, reps=3):
    p = helpers.make_problem(kind, T=T, N=N, M=M, seed=1)
    cs = helpers.product_csmc(p, cluster)
    dev = lambda x: torch.as_tensor(np.ascontiguousarray(x)).cuda()
    ref = dev(np.stack([p["ref"]] * n_chains)); Th = dev(np.stack([p["Theta"]] * n_chains)); Sg = dev(np.stack([p["Sigma"]] * n_chains))
    key = helpers.pkg("random").key(1)
    m = cs.model
    st = torch.empty((n_chains, m.T, N, m.n_x), dtype=torch.float64, device="cuda")
    anc = torch.empty((n_chains, m.T - 1, N), dtype=torch.int32, device="cuda")
    lw = torch.empty((n_chains, N), dtype=torch.float64, device="cuda")
    rng = helpers.pkg("PGAS")._make_rng(key)
    def run():
        L.check(lib.pgas_csmc_sweep_f64(m.handle, N, n_chains, L.ptr(ref), L.ptr(Th), L.ptr(Sg), C.byref(rng), L.ptr(st), L.ptr(anc),
                                        L.ptr(lw), C.c_void_p(0), C.c_void_p(0), cluster, C.c_void_p(0), 0, L.stream_ptr()))
    try:
        run(); torch.cuda.synchronize()
    except Exception as e:
        return dict(kind=kind, N=N, T=T, M=M, chains=n_chains, cluster=cluster, error=str(e))
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    ps = n_chains * N * (T - 1) / (best * 1e-3)
    return dict(kind=kind, N=N, T=T, M=p["M"], chains=n_chains, cluster=cluster, ms=best, us_per_step=best * 1e3 / (T - 1),
                particle_steps_per_s=ps, alg_tflops=ps * (2 * p["M"] * 2 + p["M"] * 2) / 1e12, finite=bool(torch.isfinite(st).all()))


pp = helpers.make_problem("smo", T=5, N=4096, M=256, seed=1)
csq = helpers.product_csmc(pp, 16)
lib.pgas_debug_max_active_clusters.restype = C.c_int
lib.pgas_debug_max_active_clusters.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
print(json.dumps({"max_active_clusters": {c: lib.pgas_debug_max_active_clusters(csq.model.handle, 4096, c) for c in range(2, 17)}}), flush=True)

for cfg in [("smo", 4096, 201, 256, 1, 16), ("smo", 4096, 201, 256, 4, 16), ("smo", 4096, 201, 256, 8, 16), ("smo", 4096, 201, 256, 8, 8), ("smo", 4096, 201, 256, 16, 8), ("smo", 4096, 201, 256, 37, 4),
            ("smo", 4096, 201, 256, 74, 2), ("smo", 2048, 201, 256, 148, 1), ("smo", 200, 751, 41, 1, 1), ("smo", 200, 751, 41, 148, 1),
            ("vehicle", 16384, 101, 1024, 8, 16), ("vehicle", 16384, 101, 1024, 9, 16)]:
    print(json.dumps(time_sweep(*cfg)), flush=True)
