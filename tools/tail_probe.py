"""Developer aid: time the sufficient statistics and the MNIW draw at configuration-scale bases.
usage: tail_probe.py kind M T chains [kind M T chains ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402

BI, PG, RND = helpers.pkg("BayesianInferrence"), helpers.pkg("PGAS"), helpers.pkg("random")
L = helpers.pkg("_lib")
import ctypes as C
lib = L.lib()
dfma, dmma = C.c_double(), C.c_double()
L.check(lib.pgas_measure_fp64_peaks(C.byref(dfma), C.byref(dmma), L.stream_ptr()))
peak = max(dfma.value, dmma.value)
print(dict(dfma=dfma.value, dmma=dmma.value), flush=True)


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


args = sys.argv[1:]
for i in range(0, len(args), 4):
    kind, M, T, nc = args[i], int(args[i + 1]), int(args[i + 2]), int(args[i + 3])
    p = helpers.make_problem(kind, T=T, N=32, M=M, seed=1)
    pg = helpers.product_pgas(p, K=2)
    m = pg.cSMC.model
    M = p["M"]
    traj = torch.as_tensor(np.stack([p["ref"]] * nc)).cuda()
    T0, T1, T2, T3 = BI.trajectory_statistics(m, traj)
    ms_s = timeit(lambda: BI.trajectory_statistics(m, traj))
    p0, p1, p2 = pg._prior()
    e0, e1, e2 = (p0 + T0).contiguous(), (p1 + T1).contiguous(), (p2 + T2).contiguous()
    rng = PG._make_rng(RND.key(3), 0, 1)
    ms_d = timeit(lambda: BI.mniw_posterior_draw(e0, e1, e2, p["prior"][3] + T3, rng))
    A, S, status = BI.mniw_posterior_draw(e0, e1, e2, p["prior"][3] + T3, rng)
    fl_s = nc * ((T - 1) * M * (M + 1) + 2 * (T - 1) * M * p["n_x"])
    fl_d = nc * (M ** 3 / 3.0)
    print(dict(kind=kind, M=M, T=T, chains=nc, suffstats_ms=ms_s, suff_tflops=fl_s / ms_s / 1e9, suff_frac=fl_s / ms_s / 1e9 / peak,
               draw_ms=ms_d, draw_tflops_m3_3=fl_d / ms_d / 1e9, draw_frac=fl_d / ms_d / 1e9 / peak,
               status=int(status.abs().sum()), finite=bool(torch.isfinite(A).all())), flush=True)
