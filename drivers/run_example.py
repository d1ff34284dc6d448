#!/usr/bin/env python
"""End-to-end driver for the three shipped examples on the B200 path.

Same pipeline as the reference's *_Simulation.py scripts (SingleMassOscillator_Simulation.py:16-125,
VehicleSimulation_Simulation.py:21-155, EMPS_Simulation.py:25-161): online marginalised filter
(Algorithm1) -> a second filter run whose sampled path initialises the Gibbs sampler -> marginalised
PGAS (Algorithm2) -> [EMPS: Theta-conditioned PGAS baseline] -> results in a .mat file with the
reference's keys.  The reference scripts additionally need `jax` for key handling and `jax.vmap` in
the post-processing; this script uses the library's Philox keys and numpy instead.

  python drivers/run_example.py smo --iterations 800 --out plots/SingleMassOscillator.mat
"""
import argparse
import os
import sys
import time as _time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def posterior_mean(prior, T0, T1, burn_in=0):
    """GP coefficient mean from iteration-averaged statistics (SingleMassOscillator_Figures.py:58-73): (1, M)"""
    from src.BayesianInferrence import prior_mniw_mean
    return prior_mniw_mean(np.asarray(prior[0]) + np.mean(T0[burn_in:], axis=0), np.asarray(prior[1]) + np.mean(T1[burn_in:], axis=0))


def run(example, iterations=None, particles=None, pgas_iterations=None, out=None, quiet=False):
    from src.Filtering import reconstruct_trajectory
    import bayesian_inference_with_explicit_and_implicit_prior_knowledge_b200.random as rnd
    say = (lambda *a: None) if quiet else print
    if example == "smo":
        import src.SingleMassOscillator as E
        alg1, alg2, names = E.SMO_Algorithm1, E.SMO_Algorithm2, ["F"]
    elif example == "vehicle":
        import src.Vehicle as E
        alg1, alg2, names = E.Vehicle_Algorithm1, E.Vehicle_Algorithm2, ["mu_f", "mu_r"]
    elif example == "emps":
        import src.EMPS as E
        alg1, alg2, names = E.EMPS_Algorithm1, E.EMPS_Algorithm2, ["F"]
    else:
        raise SystemExit(f"unknown example {example}")
    if iterations:
        alg2.N_iterations = int(iterations)
    if particles:
        alg1.N_samples = alg2.cSMC.N_samples = int(particles)
    key = E.key

    say("=== Online Algorithm ===")
    key, key_sim = rnd.split(key)
    t0 = _time.time()
    on_X, on_xi, on_stats, on_w, _, _, on_Y, on_ll = alg1(key_sim)
    say(f"Algorithm1: {_time.time() - t0:.2f} s")

    say("=== Offline Algorithm ===")
    key, key_sim, key_traj = rnd.split(key, 3)
    ref_X, ref_xi, _, ref_w, ref_anc, _, _, _ = alg1(key_sim)
    idx = int(np.searchsorted(np.cumsum(ref_w), rnd.uniform(key_traj)))      # flattened cumsum, as the reference does (:55)
    init_x = reconstruct_trajectory(ref_X, ref_anc, idx)
    init_xi = tuple(reconstruct_trajectory(v, ref_anc, idx) for v in ref_xi)
    t0 = _time.time()
    off_X, off_xi, off_w, off_stats, off_Y, off_ll = alg2(key, init_x, init_xi)
    dt = _time.time() - t0
    K = alg2.N_iterations
    say(f"Algorithm2: {K} iterations in {dt:.2f} s ({(K - 1) / dt:.1f} sweeps/s)")

    md = {"offline_Sigma_X": off_X, "offline_Sigma_Y": off_Y, "offline_weights": off_w, "offline_log_likelihood": off_ll,
          "online_Sigma_X": on_X, "online_Sigma_Y": on_Y, "online_weights": on_w, "online_log_likelihood": on_ll,
          "time": E.time, "X": E.X, "Y": E.Y}
    summary = {"K": K, "alg2_seconds": dt}
    if example == "smo":
        md.update({"offline_Sigma_F": off_xi[0], "online_Sigma_F": on_xi[0], "F_sd": E.F_sd})
        for j in range(4):
            md[f"offline_T{j}"], md[f"online_T{j}"], md[f"prior_T{j}"] = off_stats[0][j], on_stats[0][j], E.GP_prior[j]
        g = np.linspace(-3.5, 3.5, 50)
        gx, gy = np.meshgrid(g, g, indexing="xy")
        X_plot = np.vstack([gx.flatten(), gy.flatten()]).T
        basis_plot = np.asarray(E.basis_fcn(X_plot))
        md.update({"X_plot": X_plot, "basis_plot": basis_plot, "F_sd_true_plot": E.F_spring(X_plot[:, 0]) + E.F_damper(X_plot[:, 1])})
        # posterior-mean check on the states the true system visited (where the data inform the GP)
        A = posterior_mean(E.GP_prior, off_stats[0][0], off_stats[0][1], burn_in=K // 4)
        phi = np.asarray(E.basis_fcn(E.X))
        summary["rmse_F_on_trajectory"] = float(np.sqrt(np.mean((phi @ A[0] - (E.F_spring(E.X[:, 0]) + E.F_damper(E.X[:, 1]))) ** 2)))
        summary["rms_F_true"] = float(np.sqrt(np.mean((E.F_spring(E.X[:, 0]) + E.F_damper(E.X[:, 1])) ** 2)))
    elif example == "vehicle":
        def slip(Xs):                                   # (T, n, 2) states -> slip angles (T, n) front / rear
            x = np.moveaxis(Xs, -1, 0)
            u = (E.ctrl_input[:, 0][:, None], E.ctrl_input[:, 1][:, None])
            return E.f_alpha(x, u)
        oaf, oar = slip(off_X)
        naf, nar = slip(on_X)
        md.update({"offline_Sigma_mu_f": off_xi[0], "offline_Sigma_mu_r": off_xi[1], "online_Sigma_mu_f": on_xi[0], "online_Sigma_mu_r": on_xi[1],
                   "offline_Sigma_alpha_f": oaf, "offline_Sigma_alpha_r": oar, "online_Sigma_alpha_f": naf, "online_Sigma_alpha_r": nar,
                   "mu_f": E.mu_f, "mu_r": E.mu_r})
        for j in range(4):
            md[f"offline_T{j}_f"], md[f"offline_T{j}_r"] = off_stats[0][j], off_stats[1][j]
            md[f"online_T{j}_f"], md[f"online_T{j}_r"] = on_stats[0][j], on_stats[1][j]
            md[f"prior_T{j}_f"], md[f"prior_T{j}_r"] = E.GP_prior_f[j], E.GP_prior_r[j]
        alpha_plot = np.linspace(-20 / 180 * np.pi, 20 / 180 * np.pi, 500)
        md.update({"alpha_plot": alpha_plot, "basis_plot": np.asarray(E.basis_fcn(alpha_plot)), "mu_true_plot": E.mu_y(alpha_plot)})
        Af = posterior_mean(E.GP_prior_f, off_stats[0][0], off_stats[0][1], burn_in=K // 4)
        af_true, _ = E.f_alpha(E.X.T, (E.ctrl_input[:, 0], E.ctrl_input[:, 1]))
        summary["rmse_mu_f_on_trajectory"] = float(np.sqrt(np.mean((np.asarray(E.basis_fcn(af_true)) @ Af[0] - E.mu_f) ** 2)))
        summary["rms_mu_f_true"] = float(np.sqrt(np.mean(E.mu_f ** 2)))
    else:
        md.update({"offline_Sigma_F": off_xi[0], "online_Sigma_F": on_xi[0]})
        for j in range(4):
            md[f"offline_T{j}"], md[f"online_T{j}"], md[f"prior_T{j}"] = off_stats[0][j], on_stats[0][j], E.GP_prior[j]
        dq_plot = np.linspace(-0.15, 0.15, 500)
        md.update({"dq_plot": dq_plot, "basis_plot": np.asarray(E.basis_fcn(dq_plot))})
        if pgas_iterations != 0:
            say("=== PGAS baseline ===")
            if pgas_iterations:
                E.EMPS_PGAS_baseline.N_iterations = int(pgas_iterations)
            key, key_pgas = rnd.split(key)
            t0 = _time.time()
            px, pll = E.EMPS_PGAS_baseline(key_pgas, init_x)
            say(f"PGAS baseline: {E.EMPS_PGAS_baseline.N_iterations} iterations in {_time.time() - t0:.2f} s")
            md.update({"offline_Sigma_X_PGAS": px, "offline_log_likelihood_PGAS": pll})
        A = posterior_mean(E.GP_prior, off_stats[0][0], off_stats[0][1], burn_in=K // 4)
        dq = E.X[:, 1]
        F_true = 203.5 * dq + 20.39 * np.sign(dq) - 3.16                      # src/EMPS.py:171
        summary["rmse_F_on_trajectory"] = float(np.sqrt(np.mean((np.asarray(E.basis_fcn(dq)) @ A[0] - F_true) ** 2)))
        summary["rms_F_true"] = float(np.sqrt(np.mean(F_true ** 2)))
    if out:
        import scipy.io
        os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
        scipy.io.savemat(out, md)
        say("wrote", out)
    say(summary)
    return md, summary


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("example", choices=["smo", "vehicle", "emps"])
    ap.add_argument("--iterations", type=int, default=None, help="Algorithm2 iterations (default: as shipped, 800)")
    ap.add_argument("--particles", type=int, default=None)
    ap.add_argument("--pgas-iterations", type=int, default=None, help="EMPS baseline iterations (0 = skip; default as shipped, 2400)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    run(a.example, a.iterations, a.particles, a.pgas_iterations, a.out)
