import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from oracle import pgas as OP
np.set_printoptions(precision=6, linewidth=200)
K, T, N = 4, 16, 96
p = helpers.make_problem("toy", T=T, N=N, seed=11)
pg = helpers.product_pgas(p, K=K, cluster_size=1)
rng = np.random.default_rng(5)
n_x, M = p["n_x"], p["M"]
df = p["prior"][3] + T - 1
V = dict(Z=rng.normal(size=(K, 1, T, N, n_x)), U=rng.uniform(size=(K, 1, T, 2)),
         chi2=rng.chisquare(df - np.arange(n_x), size=(K, 1, n_x)), G=rng.normal(size=(K, 1, n_x, n_x)),
         Nrm=rng.normal(size=(K, 1, n_x, M)))
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
out = pg.run_chains(None, p["ref"], n_chains=1, variates={k: dev(v) for k, v in V.items()})
st_o, ll_o, A_o, S_o = OP.pgas_run(p["omodel"], N, K, p["prior"], p["ref"], lambda k: {n: V[n][k, 0] for n in V})
st_g = out["state_trace"][0].cpu().numpy()
cs = pg.cSMC
for k in range(1, K):
    print("iter", k, "traj err", helpers.rel_err(st_g[k], st_o[:, k]), "A err", helpers.rel_err(out["A_trace"][0, k].cpu().numpy(), A_o[k]))
    # replay sweep k on both sides from the ORACLE's previous trajectory and parameters
    sw = cs.sweep(dev(st_o[:, k - 1]), dev(A_o[k - 1]), dev(S_o[k - 1]), variates=dict(Z=dev(V["Z"][k]), U=dev(V["U"][k])))
    o = OP.csmc_sweep(p["omodel"], N, st_o[:, k - 1], A_o[k - 1], S_o[k - 1], V["Z"][k, 0], V["U"][k, 0], keep_weights=True)
    an = sw["anc_trace"][0].cpu().numpy()
    bad = np.argwhere(an != o["anc_trace"])
    print("  replay: idx", int(sw["idx"][0]), o["idx"], "anc mismatches", len(bad), bad[:3].tolist(),
          "state err", helpers.rel_err(sw["state_trace"][0].cpu().numpy(), o["state_trace"]),
          "traj err", helpers.rel_err(sw["traj"][0].cpu().numpy(), o["traj"]))
    if len(bad):
        t, j = bad[0]
        w_aux, w_anc = o["cdfs"][t]
        cdf = helpers.resample_cdf(w_aux); pts = (V["U"][k, 0, t + 1, 0] + np.arange(N)) / N
        print("   first mismatch t", t, "j", j, "gpu", an[t, j], "ora", o["anc_trace"][t, j], "gap", np.min(np.abs(cdf - pts[j])) if j < N - 1 else np.min(np.abs(np.cumsum(w_anc) - V["U"][k, 0, t + 1, 1])))
    # final pick check
    cdf_f = np.cumsum(o["w_final"]); print("  final u", V["U"][k, 0, 0, 0], "gap", np.min(np.abs(cdf_f - V["U"][k, 0, 0, 0])))
    A_g, S_g = pg.sample_params(None, dev(st_o[:, k][None]), variates=dict(chi2=dev(V["chi2"][k]), G=dev(V["G"][k]), Nrm=dev(V["Nrm"][k])))
    print("  draw replay A err", helpers.rel_err(A_g[0].cpu().numpy(), A_o[k]), "S err", helpers.rel_err(S_g[0].cpu().numpy(), S_o[k]))
