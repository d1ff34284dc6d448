"""Regenerates tests/golden/oracle_marginal_golden.json: frozen outputs of the ORACLE restatement of the
marginalised filters (oracle/marginal.py) on seeded inputs.  Like oracle_golden.json these guard the oracle
against silent regressions; they do not pin it to the reference, which cannot be run here.
Run:  python tests/golden/make_golden_marginal.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import helpers_marginal as HM  # noqa: E402
from oracle import marginal as OMg  # noqa: E402

CASES = {
    "smo_T14_N24_M10": dict(kind="smo", T=14, N=24, M=10, seed=1, vseed=11),
    "emps_T16_N20_M9": dict(kind="emps", T=16, N=20, M=9, seed=2, vseed=12),
    "vehicle_T12_N20_M6": dict(kind="vehicle", T=12, N=20, M=6, seed=3, vseed=13),
}


def run_case(g):
    prob = HM.make_marg_problem(g["kind"], T=g["T"], N=g["N"], M=g["M"], seed=g["seed"])
    V = HM.make_variates(prob, 0.999, seed=g["vseed"])
    f = OMg.alg1_run(prob["oracle"], g["N"], 0.999, HM.oracle_variates(V))
    V1 = HM.make_variates(prob, 1.0, seed=g["vseed"] + 100)
    f1 = OMg.alg1_run(prob["oracle"], g["N"], 1.0, HM.oracle_variates(V1))
    ref_x = f1["state_trace"][:, 0]
    ref_xi = [f1["int_var_trace"][k][:, 0, 0] for k in range(prob["G"])]
    rs = OMg.reference_stats(prob["oracle"], ref_x, ref_xi)
    V3 = HM.make_variates(prob, 1.0, seed=g["vseed"] + 200)
    c = OMg.alg3_run(prob["oracle"], g["N"], ref_x, ref_xi, rs, HM.oracle_variates(V3))
    return dict(alg1_anc_last=f["ancestor_trace"][-1].tolist(), alg1_state_last=f["state_trace"][-1].ravel().tolist(),
                alg1_logw_last=f["logw_trace"][-1].tolist(), alg1_sst_T0_last=np.asarray(f["suff_stats_trace"][0][0][-1]).ravel().tolist(),
                alg3_anc_last=c["anc_trace"][-1].tolist(), alg3_idx=int(c["idx"]), alg3_traj=c["traj"].ravel().tolist(),
                alg3_xi_traj=c["xi_traj"][0].ravel().tolist(), ref_T1_trace=float(np.trace(rs[0][1])))


if __name__ == "__main__":
    out = {name: dict(g, **run_case(g)) for name, g in CASES.items()}
    with open(os.path.join(HERE, "oracle_marginal_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", len(out), "cases")
