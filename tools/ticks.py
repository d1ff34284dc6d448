"""phase clocks of one sweep (developer aid): python tools/ticks.py kind N T M chains cluster"""
import ctypes as C, sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
kind, N, T, M, nc, cl = sys.argv[1], *map(int, sys.argv[2:7])
L = helpers.pkg("_lib"); lib = L.lib()
p = helpers.make_problem(kind, T=T, N=N, M=M, seed=1)
cs = helpers.product_csmc(p, cl)
m = cs.model
dev = lambda x: torch.as_tensor(np.ascontiguousarray(x)).cuda()
ref = dev(np.stack([p["ref"]] * nc)); Th = dev(np.stack([p["Theta"]] * nc)); Sg = dev(np.stack([p["Sigma"]] * nc))
st = torch.empty((nc, T, N, 2), dtype=torch.float64, device="cuda"); an = torch.empty((nc, T - 1, N), dtype=torch.int32, device="cuda")
lw = torch.empty((nc, N), dtype=torch.float64, device="cuda")
dbg = torch.zeros(((T - 1) * 2 * 8,), dtype=torch.int64, device="cuda")
lib.pgas_debug_sweep_ticks.restype = C.c_int
lib.pgas_debug_sweep_ticks.argtypes = [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 6 + [C.c_int32, C.c_void_p, C.c_void_p]
for _ in range(2):
    L.check(lib.pgas_debug_sweep_ticks(m.handle, N, nc, L.ptr(ref), L.ptr(Th), L.ptr(Sg), L.ptr(st), L.ptr(an), L.ptr(lw), cl, L.ptr(dbg), L.stream_ptr()))
torch.cuda.synchronize()
d = dbg.cpu().numpy().reshape(T - 1, 2, 8)
names = ["A (DMMA+softmax)", "wait sync1", "X1 fold (warp0)", "wait sync2", "B1 cdf+count", "B2 resample", "X2+C"]
for th in (0, 1):
    seg = np.diff(d[5:, th, :], axis=1)                    # phases within a step
    nxt = d[6:, th, 0] - d[5:-1, th, 7]
    print("CTA", th, "(thread 96) cycles/step", np.median(d[6:, th, 0] - d[5:-1, th, 0]))
    for k, n in enumerate(names):
        print(f"   {n:20s} {np.median(seg[:, k]):8.0f}")
    print(f"   {'loop back':20s} {np.median(nxt):8.0f}")
if hasattr(lib, "pgas_debug_fine_ticks"):
    buf = (C.c_longlong * 64)()
    lib.pgas_debug_fine_ticks(buf)
    f = np.array(list(buf), dtype=np.int64)
    lab = {0: "start", 1: "gp_input", 2: "seeds"}
    for h in range(4):
        lab.update({3 + 4 * h: f"p{h} sines", 4 + 4 * h: f"p{h} mma", 5 + 4 * h: f"p{h} scale", 6 + 4 * h: f"p{h} reduce"})
    lab.update({20: "loglik", 21: "warp_max", 22: "exp", 23: "scan"})
    keys = sorted(lab)
    prev = f[0]
    for k in keys[1:]:
        print(f"   fine {lab[k]:12s} {f[k] - prev:7d}")
        prev = f[k]
