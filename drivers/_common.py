"""Helpers shared by the three example drivers (drivers/*_Simulation.py).

The reference drivers need `jax` for three things only: key handling (`jax.random.split`, `jax.random.uniform`), `jax.vmap` over
NumPy-style callables in the post-processing, and `jnp.linspace`.  Here keys are the library's Philox keys
(bayesian_inference_with_explicit_and_implicit_prior_knowledge_b200.random) and `vmap` is a plain loop / a batched device call.
"""
import argparse
import os
import sys
import time as _time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import bayesian_inference_with_explicit_and_implicit_prior_knowledge_b200.random as rnd  # noqa: E402


def options(description, default_out):
    """Command line of a driver: the defaults are the shipped settings (the reference drivers take no options)."""
    ap = argparse.ArgumentParser(description=description)
    ap.add_argument("--iterations", type=int, default=None, help="Algorithm2 iterations (default: as shipped)")
    ap.add_argument("--particles", type=int, default=None, help="particles (default: as shipped)")
    ap.add_argument("--pgas-iterations", type=int, default=None, help="EMPS only: iterations of the Theta-conditioned baseline")
    ap.add_argument("--out", default=default_out, help="result file (.mat)")
    return ap.parse_args()


def resize(alg1, alg2, opts, baseline=None):
    if opts.iterations:
        alg2.N_iterations = int(opts.iterations)
    if opts.particles:
        alg1.N_samples = alg2.cSMC.N_samples = int(opts.particles)
        if baseline is not None:
            baseline.cSMC.N_samples = int(opts.particles)
    if baseline is not None and opts.pgas_iterations:
        baseline.N_iterations = int(opts.pgas_iterations)


def vmap(fn, *columns):
    """Row-wise map of a NumPy-style callable (stand-in for jax.vmap in the drivers' post-processing)."""
    return np.stack([np.asarray(fn(*row)) for row in zip(*columns)])


def timed(label, fn, *args):
    t0 = _time.time()
    out = fn(*args)
    print(f"{label}: {_time.time() - t0:.2f} s")
    return out


def initial_reference(alg1, key_sim, key_traj, reconstruct_trajectory):
    """A second filter run and one sampled path of it (the reference picks the index from the flattened (T, N) cumulative sum of
    the weight trace: SingleMassOscillator_Simulation.py:55, VehicleSimulation_Simulation.py:67, EMPS_Simulation.py:62)."""
    states, int_vars, _, weights, ancestors, _, _, _ = alg1(key_sim)
    idx = np.searchsorted(np.cumsum(weights), rnd.uniform(key_traj))
    ref_state = reconstruct_trajectory(states, ancestors, idx)
    ref_int_var = tuple(reconstruct_trajectory(v, ancestors, idx) for v in int_vars)
    return ref_state, ref_int_var


def put_statistics(mdict, prefix, stats, suffix=""):
    """offline_T0 .. offline_T3 (+ suffix) from one GP's 4-tuple of traced statistics"""
    for j, value in enumerate(stats):
        mdict[f"{prefix}_T{j}{suffix}"] = value


def save(path, mdict):
    import scipy.io
    os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
    scipy.io.savemat(path, mdict)
    print("wrote", path, f"({len(mdict)} variables)")
