#!/usr/bin/env python
"""Electro-mechanical-positioning-system experiment on the B200 path — the pipeline of the reference's EMPS_Simulation.py (:25-161):
Algorithm1 online, sampled reference path, Algorithm2 offline (friction GP, M = 9), the Theta-conditioned PGAS baseline with the
three-dimensional basis (M = 729), free-run validation of both learned models, results in plots/EMPS.mat with the reference's
variable names (:128-160).  The imports from `src.*` are the reference's own lines; only the `jax` calls differ.

Two deviations, both forced by the reference as shipped: (i) its measurement files are not in its repository — without them
`src.EMPS` synthesises data of the measured shape and the validation runs on that record; (ii) its statistics of the baseline
(:102-113) materialise a (T-1, K, M, M) array (25 TB at the shipped sizes); here the same sums come from the device kernel
(BayesianInferrence.trajectory_statistics), one trajectory at a time."""
import numpy as np

from _common import initial_reference, options, put_statistics, resize, rnd, save, timed

from src.Filtering import reconstruct_trajectory
from src.EMPS import (
    GP_prior,
    Y,
    X,
    EMPS_Algorithm1,
    EMPS_Algorithm2,
    time,
    basis_fcn,
    key,
    EMPS_PGAS_baseline,
    basis_fcn_f_PGAS,
    GP_prior_PGAS,
    ctrl_input,
    EMPS_Validation_Simulation,
)
import src.BayesianInferrence as BI
import src.EMPS as _emps

opts = options(__doc__, "plots/EMPS.mat")
resize(EMPS_Algorithm1, EMPS_Algorithm2, opts, baseline=EMPS_PGAS_baseline)

print("\n=== Online Algorithm ===")
key, key_sim = rnd.split(key)
online = timed("Algorithm1", EMPS_Algorithm1, key_sim)
online_Sigma_X, online_Sigma_F, online_GP_stats, online_weights = online[:4]
online_Sigma_Y, online_log_likelihood = online[6:]

print("\n=== Offline Algorithm ===")
key, key_sim, key_traj = rnd.split(key, 3)
init_ref_state, init_ref_int_var = initial_reference(EMPS_Algorithm1, key_sim, key_traj, reconstruct_trajectory)
offline = timed(f"Algorithm2 ({EMPS_Algorithm2.N_iterations} iterations)", EMPS_Algorithm2, key, init_ref_state, init_ref_int_var)
offline_Sigma_X, offline_Sigma_F, offline_weights, offline_GP_stats, offline_Sigma_Y, offline_log_likelihood = offline
offline_mean = BI.prior_mniw_2naturalPara_inv(*[GP_prior[j] + np.mean(offline_GP_stats[0][j], axis=0) for j in range(4)])[0]

print("\n=== Offline Algorithm (PGAS) ===")
offline_Sigma_X_PGAS, offline_log_likelihood_PGAS = timed(
    f"PGAS baseline ({EMPS_PGAS_baseline.N_iterations} iterations)", EMPS_PGAS_baseline, key, init_ref_state)
# iteration-averaged statistics of the sampled trajectories (EMPS_Simulation.py:98-113), 64 trajectories per device call
import torch
traj = torch.as_tensor(np.ascontiguousarray(np.moveaxis(np.asarray(offline_Sigma_X_PGAS), 1, 0))).cuda()      # (K, T, n_x)
sums = [0.0, 0.0, 0.0, 0.0]
for k0 in range(0, traj.shape[0], 64):
    part = BI.trajectory_statistics(EMPS_PGAS_baseline.cSMC.model, traj[k0:k0 + 64])
    for j in range(3):
        sums[j] = sums[j] + part[j].sum(dim=0)
    sums[3] += part[3] * part[0].shape[0]
PGAS_Posterior_Stats = [np.asarray(GP_prior_PGAS[j]) + (sums[j].cpu().numpy() if j < 3 else sums[j]) / traj.shape[0] for j in range(4)]
PGAS_mean = BI.prior_mniw_2naturalPara_inv(*PGAS_Posterior_Stats)[0]

validation = None if not _emps.data_is_synthetic else _emps.data     # no DATA_EMPS_PULSES.mat: validate on the identification record
RMSE_Alg2, RMSE_PGAS = EMPS_Validation_Simulation(offline_mean, PGAS_mean, validation)
print(f"RMSE_Alg2: {RMSE_Alg2}")
print(f"RMSE_PGAS: {RMSE_PGAS}")

dq_plot = np.linspace(-0.15, 0.15, 500)
mdict = {"time": time, "X": X, "Y": Y, "dq_plot": dq_plot, "basis_plot": np.asarray(basis_fcn(dq_plot)),
         "offline_Sigma_X_PGAS": offline_Sigma_X_PGAS, "offline_log_likelihood_PGAS": offline_log_likelihood_PGAS,
         "RMSE_Alg2": RMSE_Alg2, "RMSE_PGAS": RMSE_PGAS}
for side, (Sigma_X, Sigma_Y, Sigma_F, weights, log_likelihood, stats_g) in {
        "offline": (offline_Sigma_X, offline_Sigma_Y, offline_Sigma_F[0], offline_weights, offline_log_likelihood, offline_GP_stats[0]),
        "online": (online_Sigma_X, online_Sigma_Y, online_Sigma_F[0], online_weights, online_log_likelihood, online_GP_stats[0])}.items():
    mdict.update({f"{side}_Sigma_X": Sigma_X, f"{side}_Sigma_Y": Sigma_Y, f"{side}_Sigma_F": Sigma_F, f"{side}_weights": weights,
                  f"{side}_log_likelihood": log_likelihood})
    put_statistics(mdict, side, stats_g)
put_statistics(mdict, "prior", GP_prior)
save(opts.out, mdict)
