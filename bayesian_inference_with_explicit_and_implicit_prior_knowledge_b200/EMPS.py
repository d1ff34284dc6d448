"""Electro-mechanical positioning system example (reference src/EMPS.py): a motor-driven carriage
whose friction force F(dq) is the unknown function (1-D Hilbert-space GP), plus the Theta-conditioned
PGAS baseline with a 3-D basis (M = 729).  Module-level names match the reference.

The reference loads `src/Measurements/DATA_EMPS.mat`, which is not part of its repository
(SURVEY.md 8c).  When the file is absent this module synthesises a trajectory of the same shape
(T = 2484 samples at 100 Hz after the reference's decimation by 10) from the reference's own linear
friction model (src/EMPS.py:168-174) driven by a band-limited force within the +-100 N range of
`vir * gtau`; set EMPS_DATA to the .mat path to use the real measurements."""
import os

import numpy as np

from . import models as _models
from . import random as _random
from . import stats as _stats
from ._examples import rk4_step
from .Algorithm1 import Algorithm1
from .Algorithm2 import Algorithm2
from .BasisFunctions import generate_Hilbert_BasisFunction
from .BayesianInferrence import prior_mniw_2naturalPara
from .PGAS import PGAS
from .StateSpaceModel import StateSpaceModel

N_particles, N_PGAS_iter, forget_factor = 200, 800, 0.999
key = _random.key(12345678)
M = 95.11


def central_difference_quotient(x, t):
    x, t = np.asarray(x, dtype=np.float64), np.asarray(t, dtype=np.float64)
    d = np.empty_like(x)
    d[1:-1] = (x[2:] - x[:-2]) / (t[2:] - t[:-2])
    d[0] = (x[1] - x[0]) / (t[1] - t[0])
    d[-1] = (x[-1] - x[-2]) / (t[-1] - t[-2])
    return d


def dx(x, tau, F):
    """x = [q, dq]"""
    return np.hstack([x[1], (tau - F) / M])


def dx_linModel(x, tau):
    return np.hstack([x[1], (tau - 203.5 * x[1] - 20.39 * np.sign(x[1]) + 3.16) / 95.11])


def f_x_linModel(x, tau, dt):
    return rk4_step(lambda s: dx_linModel(s, tau), x, dt)


def f_y(x):
    return x[0]


def _synthetic_measurements(n=24841, fs=1000.0, seed=12345678):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    tau = np.zeros(n)
    for f in (0.13, 0.31, 0.57, 0.83, 1.21):
        tau += rng.uniform(10.0, 25.0) * np.sin(2 * np.pi * f * t + rng.uniform(0, 2 * np.pi))
    tau = np.clip(tau, -100.0, 100.0)
    x = np.zeros((n, 2))
    x[0] = [0.1, 0.0]
    for i in range(1, n):
        x[i] = f_x_linModel(x[i - 1], tau[i - 1], 1.0 / fs)
    return dict(t=t, qm=x[:, 0] + 1e-5 * rng.normal(size=n), vir=tau, gtau=np.ones(1))


_path = os.environ.get("EMPS_DATA", "src/Measurements/DATA_EMPS.mat")
if os.path.exists(_path):
    import scipy.io
    data = scipy.io.loadmat(_path)
    data_is_synthetic = False
else:
    data = _synthetic_measurements()
    data_is_synthetic = True

# reference data: low-pass + differentiate + decimate by 10 (src/EMPS.py:52-65)
import scipy.signal as _sig
q_ref = _sig.sosfiltfilt(_sig.butter(4, 100 / 500, btype="lowpass", output="sos"), np.asarray(data["qm"]).flatten())
dq_ref = central_difference_quotient(q_ref, np.asarray(data["t"]).flatten())
X = np.vstack([q_ref, dq_ref]).T[0:-1:10]
time = np.asarray(data["t"]).flatten()[0:-1:10]
Y = np.asarray(data["qm"]).flatten()[0:-1:10]
steps = time.shape[0]
dt = time[1] - time[0]
x0 = np.array([Y[0], 0])
P0 = np.diag([1e-5, 1e-6])
P0_F = np.diag([1e-12])
R = np.diag([1e-4])
Q = np.diag([1e-6, 1e-7])
ctrl_input = (np.asarray(data["vir"]) * np.asarray(data["gtau"])).flatten()[0:-1:10]


def f_x(x, tau, F, dt=dt):
    return rk4_step(lambda s: dx(s, tau, F), x, dt)


# ---- friction GP (src/EMPS.py:84-99) and the 3-D baseline basis (:101-123)
N_basis_fcn = 9
basis_fcn, sd = generate_Hilbert_BasisFunction(N_basis_fcn, np.array([-0.2, 0.2]), 0.4 / N_basis_fcn, 20)


def basis_fcn_f(state, input):
    return basis_fcn(state[1])


GP_prior = list(prior_mniw_2naturalPara(np.zeros((1, N_basis_fcn)), np.diag(sd), np.eye(1) * 4, 2))

N_basis_fcn_baseline = N_basis_fcn ** 3
basis_fcn_baseline, sd_baseline = generate_Hilbert_BasisFunction(N_basis_fcn_baseline, np.array([[-1, 1], [-1, 1], [-1, 1]]),
                                                                 0.5 / N_basis_fcn_baseline, 20)


def basis_fcn_f_PGAS(state, input):
    return basis_fcn_baseline(_models.hstack([state, input]) / np.array([0.4, 0.4, 160]))


GP_prior_PGAS = list(prior_mniw_2naturalPara(np.zeros((2, N_basis_fcn_baseline)), np.diag(sd_baseline), np.eye(2), 2))


def EMPS_Validation_Simulation(GP_Mean_Alg2, GP_mean_PGAS, data=None):
    """free-run validation of both learned models (src/EMPS.py:129-152); needs DATA_EMPS_PULSES.mat or `data`"""
    if data is None:
        import scipy.io
        data = scipy.io.loadmat("src/Measurements/DATA_EMPS_PULSES.mat")
    tv = np.asarray(data["t"]).flatten()[0:-1:10]
    Yv = np.asarray(data["qm"]).flatten()[0:-1:10]
    Tau = (np.asarray(data["vir"]) * np.asarray(data["gtau"])).flatten()[0:-1:10]
    n, h = tv.shape[0], tv[1] - tv[0]
    Xa, Xp = np.zeros((n, 2)), np.zeros((n, 2))
    Xa[0] = Xp[0] = [Yv[0], 0]
    phi_f = np.asarray(basis_fcn(Xa[:1, 1]))      # warm the evaluator
    for i in range(1, n):
        F = (GP_Mean_Alg2 @ np.asarray(basis_fcn(Xa[i - 1, 1])))[0]
        Xa[i] = f_x(x=Xa[i - 1], tau=Tau[i - 1], F=F, dt=h)
        Xp[i] = GP_mean_PGAS @ np.asarray(basis_fcn_baseline(np.hstack([Xp[i - 1], Tau[i - 1]]) / np.array([0.4, 0.4, 160])))
    return np.sqrt(np.mean((Xa[:, 0] - Yv) ** 2)), np.sqrt(np.mean((Xp[:, 0] - Yv) ** 2))


EMPS_SSM = StateSpaceModel(process_noise=Q, output_noise=R,
                           transition_model=lambda state, input, *int_var: f_x(state, input, int_var[0], dt),
                           output_model=lambda state, input, *int_var: f_y(state))

_common = dict(observations=Y, inputs=ctrl_input, SSM=EMPS_SSM, init_state_mean=x0, init_state_cov=P0,
               init_int_var_mean=[np.array([0.0])], init_int_var_cov=[P0_F], GP_prior=[GP_prior], basis_fcn=[basis_fcn_f])
EMPS_Algorithm1 = Algorithm1(N_samples=N_particles, forgetting_factor=forget_factor, **_common)
EMPS_Algorithm2 = Algorithm2(N_samples=N_particles, N_iterations=N_PGAS_iter, **_common)

EMPS_PGAS_baseline = PGAS(N_samples=N_particles, N_iterations=N_PGAS_iter * 3, observations=Y, inputs=ctrl_input,
                          init_state_mean=x0, init_state_cov=P0,
                          # the reference's own lambda (src/EMPS.py:250-252) with this package's stats module in place of jax.scipy.stats
                          likelihood_fcn=lambda obs, state, input: np.squeeze(_stats.multivariate_normal.logpdf(obs, mean=f_y(state), cov=R)),
                          GP_prior=GP_prior_PGAS, basis_fcn=basis_fcn_f_PGAS)
