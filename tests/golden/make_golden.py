"""Regenerates tests/golden/oracle_golden.json.

The reference (JAX 0.4.38 / equinox 0.12.2) cannot be imported in the build container and ships
no golden vectors of its own, so these fixtures freeze outputs of the ORACLE
restatement on seeded inputs: they guard the oracle against silent regressions; they do not pin
it to the reference ("parity unpinned", see oracle/__init__.py).
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import helpers  # noqa: E402
from oracle import pgas as OP  # noqa: E402

CASES = {
    "smo_T25_N50": dict(kind="smo", T=25, N=50, seed=1, vseed=11),
    "emps_T20_N40": dict(kind="emps", T=20, N=40, seed=2, vseed=12),
    "toy_T30_N30": dict(kind="toy", T=30, N=30, seed=3, vseed=13),
    "vehicle_T20_N64": dict(kind="vehicle", T=20, N=64, seed=4, vseed=14),
}

out = {}
for name, g in CASES.items():
    p = helpers.make_problem(g["kind"], T=g["T"], N=g["N"], seed=g["seed"])
    Z, U = helpers.sweep_variates(p, g["vseed"])
    o = OP.csmc_sweep(p["omodel"], g["N"], p["ref"], p["Theta"], p["Sigma"], Z, U)
    rng = np.random.default_rng(g["vseed"])
    df = p["prior"][3] + g["T"] - 1
    chi2 = rng.chisquare(df - np.arange(p["n_x"]))
    G = rng.normal(size=(p["n_x"],) * 2)
    Nrm = rng.normal(size=(p["n_x"], p["M"]))
    A, S, _ = OP.sample_params(p["omodel"], p["prior"], p["ref"], chi2, G, Nrm)
    out[name] = dict(g, anc_last=o["anc_trace"][-1].tolist(), idx=int(o["idx"]), traj=o["traj"].ravel().tolist(),
                     S=S.ravel().tolist(), A_head=A.ravel()[:16].tolist())
with open(os.path.join(HERE, "oracle_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(out), "cases")
