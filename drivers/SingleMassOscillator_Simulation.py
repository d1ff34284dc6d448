#!/usr/bin/env python
"""Single-mass-oscillator experiment on the B200 path — the pipeline of the reference's SingleMassOscillator_Simulation.py
(:16-125): online marginalised filter (Algorithm1), a second filter run whose sampled path starts the Gibbs sampler, marginalised
PGAS (Algorithm2), results in plots/SingleMassOscillator.mat with the reference's variable names (:94-124).
The imports from `src.*` are the reference's own lines; only the `jax` calls differ (drivers/_common.py)."""
import numpy as np

from _common import initial_reference, options, put_statistics, resize, rnd, save, timed, vmap

from src.Filtering import reconstruct_trajectory
from src.SingleMassOscillator import F_spring, F_damper, basis_fcn, time, key
from src.SingleMassOscillator import (
    GP_prior,
    SMO_Algorithm1,
    X,
    Y,
    F_sd,
    SMO_Algorithm2,
)

opts = options(__doc__, "plots/SingleMassOscillator.mat")
resize(SMO_Algorithm1, SMO_Algorithm2, opts)

print("\n=== Online Algorithm ===")
key, key_sim = rnd.split(key)
online = timed("Algorithm1", SMO_Algorithm1, key_sim)
online_Sigma_X, online_Sigma_F, online_GP_stats, online_weights = online[:4]
online_Sigma_Y, online_log_likelihood = online[6:]

print("\n=== Offline Algorithm ===")
key, key_sim, key_traj = rnd.split(key, 3)
init_ref_state, init_ref_int_var = initial_reference(SMO_Algorithm1, key_sim, key_traj, reconstruct_trajectory)
offline = timed(f"Algorithm2 ({SMO_Algorithm2.N_iterations} iterations)", SMO_Algorithm2, key, init_ref_state, init_ref_int_var)
offline_Sigma_X, offline_Sigma_F, offline_weights, offline_GP_stats, offline_Sigma_Y, offline_log_likelihood = offline

# input grid of the spring-damper force for the figures (:81-91)
axis = np.linspace(-3.5, 3.5, 50)
X_plot = np.stack([g.flatten() for g in np.meshgrid(axis, axis, indexing="xy")], axis=1)

mdict = {"time": time, "X": X, "Y": Y, "F_sd": F_sd, "X_plot": X_plot,
         "basis_plot": np.asarray(basis_fcn(X_plot)),                       # batched device evaluation instead of jax.vmap
         "F_sd_true_plot": vmap(F_spring, X_plot[:, 0]) + vmap(F_damper, X_plot[:, 1])}
for side, (Sigma_X, Sigma_Y, Sigma_F, weights, log_likelihood, stats) in {
        "offline": (offline_Sigma_X, offline_Sigma_Y, offline_Sigma_F[0], offline_weights, offline_log_likelihood, offline_GP_stats[0]),
        "online": (online_Sigma_X, online_Sigma_Y, online_Sigma_F[0], online_weights, online_log_likelihood, online_GP_stats[0])}.items():
    mdict.update({f"{side}_Sigma_X": Sigma_X, f"{side}_Sigma_Y": Sigma_Y, f"{side}_Sigma_F": Sigma_F, f"{side}_weights": weights,
                  f"{side}_log_likelihood": log_likelihood})
    put_statistics(mdict, side, stats)
put_statistics(mdict, "prior", GP_prior)
save(opts.out, mdict)
