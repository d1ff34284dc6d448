#!/usr/bin/env python
"""Lateral-vehicle-dynamics experiment on the B200 path — the pipeline of the reference's VehicleSimulation_Simulation.py (:21-155):
Algorithm1 online, sampled reference path, Algorithm2 offline with two tyre-friction GPs (front / rear), results in plots/Vehicle.mat
with the reference's variable names (:105-154).  The imports from `src.*` are the reference's own lines; only the `jax` calls differ."""
import numpy as np

from _common import initial_reference, options, put_statistics, resize, rnd, save, timed

from src.Filtering import reconstruct_trajectory
from src.Vehicle import mu_y, basis_fcn, f_alpha, ctrl_input
from src.Vehicle import time, Vehicle_Algorithm1, Vehicle_Algorithm2
from src.Vehicle import (
    GP_prior_f,
    GP_prior_r,
    key,
    X,
    Y,
    mu_f,
    mu_r,
)

opts = options(__doc__, "plots/Vehicle.mat")
resize(Vehicle_Algorithm1, Vehicle_Algorithm2, opts)


def slip_angles(Sigma_X):
    """f_alpha over a state trace (T, n, 2) with the input of each time step (the reference nests two jax.vmap, :45-47)"""
    states = np.moveaxis(np.asarray(Sigma_X), -1, 0)                          # (2, T, n)
    return f_alpha(states, (ctrl_input[:, 0][:, None], ctrl_input[:, 1][:, None]))


print("\n=== Online Algorithm ===")
key, key_sim = rnd.split(key)
online = timed("Algorithm1", Vehicle_Algorithm1, key_sim)
online_Sigma_X, online_Sigma_mu, online_GP_stats, online_weights = online[:4]
online_Sigma_Y, online_log_likelihood = online[6:]

print("\n=== Offline Algorithm ===")
key, key_sim, key_traj = rnd.split(key, 3)
init_ref_state, init_ref_int_var = initial_reference(Vehicle_Algorithm1, key_sim, key_traj, reconstruct_trajectory)
offline = timed(f"Algorithm2 ({Vehicle_Algorithm2.N_iterations} iterations)", Vehicle_Algorithm2, key, init_ref_state, init_ref_int_var)
offline_Sigma_X, offline_Sigma_mu, offline_weights, offline_GP_stats, offline_Sigma_Y, offline_log_likelihood = offline

alpha_plot = np.linspace(-20 / 180 * np.pi, 20 / 180 * np.pi, 500)
mdict = {"time": time, "X": X, "Y": Y, "mu_f": mu_f, "mu_r": mu_r, "alpha_plot": alpha_plot,
         "mu_true_plot": mu_y(alpha=alpha_plot), "basis_plot": np.asarray(basis_fcn(alpha_plot))}
for side, (Sigma_X, Sigma_Y, Sigma_mu, weights, log_likelihood, stats) in {
        "offline": (offline_Sigma_X, offline_Sigma_Y, offline_Sigma_mu, offline_weights, offline_log_likelihood, offline_GP_stats),
        "online": (online_Sigma_X, online_Sigma_Y, online_Sigma_mu, online_weights, online_log_likelihood, online_GP_stats)}.items():
    alpha_f, alpha_r = slip_angles(Sigma_X)
    mdict.update({f"{side}_Sigma_X": Sigma_X, f"{side}_Sigma_Y": Sigma_Y, f"{side}_weights": weights, f"{side}_log_likelihood": log_likelihood,
                  f"{side}_Sigma_mu_f": Sigma_mu[0], f"{side}_Sigma_mu_r": Sigma_mu[1],
                  f"{side}_Sigma_alpha_f": alpha_f, f"{side}_Sigma_alpha_r": alpha_r})
    put_statistics(mdict, side, stats[0], "_f")
    put_statistics(mdict, side, stats[1], "_r")
# the reference stores the FRONT statistic under this rear key (VehicleSimulation_Simulation.py:136); kept so that the file
# is interchangeable with one written by the reference
mdict["online_T2_r"] = mdict["online_T2_f"]
put_statistics(mdict, "prior", GP_prior_f, "_f")
put_statistics(mdict, "prior", GP_prior_r, "_r")
save(opts.out, mdict)
