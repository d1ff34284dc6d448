import sys, os, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
g.build()
import helpers
from oracle import pgas as OP
np.set_printoptions(precision=6, linewidth=200)
p = helpers.make_problem("emps", T=50, N=200, seed=50)
print("emps sweep:", helpers.run_sweep_parity(p, cluster_size=1))
# run_chains toy
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
for k in range(K):
    print("iter", k, "traj err", helpers.rel_err(st_g[k], st_o[:, k]), "A err", helpers.rel_err(out["A_trace"][0, k].cpu().numpy(), A_o[k]),
          "S", out["S_trace"][0, k].cpu().numpy().ravel(), S_o[k].ravel())
# direct sweep for iteration 1 with oracle's A,S
cs = pg.cSMC
sw = cs.sweep(dev(p["ref"]), dev(A_o[0]), dev(S_o[0]), variates=dict(Z=dev(V["Z"][1]), U=dev(V["U"][1])))
o = OP.csmc_sweep(p["omodel"], N, p["ref"], A_o[0], S_o[0], V["Z"][1, 0], V["U"][1, 0], keep_weights=True)
print("direct sweep: idx", int(sw["idx"][0]), o["idx"], "anc equal", np.array_equal(sw["anc_trace"][0].cpu().numpy(), o["anc_trace"]),
      "state err", helpers.rel_err(sw["state_trace"][0].cpu().numpy(), o["state_trace"]), "traj err", helpers.rel_err(sw["traj"][0].cpu().numpy(), o["traj"]))
print("traj g", sw["traj"][0].cpu().numpy().ravel()[:8], "\ntraj o", o["traj"].ravel()[:8])
print("run_chains traj1", st_g[1].ravel()[:8])
# occupancy query for clusters
