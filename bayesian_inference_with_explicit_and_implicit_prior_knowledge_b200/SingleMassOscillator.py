"""Single-mass oscillator example (reference src/SingleMassOscillator.py): a mass on a nonlinear
spring-damper whose force F_sd(s, v) is the unknown function, learned with a 2-D Hilbert-space GP.
Module-level names match the reference (the drivers import them by name).
Importing runs the ground-truth simulation, like the reference does (:136-137)."""
import numpy as np

from . import random as _random
from ._examples import rk4_step, simulate
from .Algorithm1 import Algorithm1
from .Algorithm2 import Algorithm2
from .BasisFunctions import generate_Hilbert_BasisFunction
from .BayesianInferrence import prior_mniw_2naturalPara
from .StateSpaceModel import StateSpaceModel

# ---- plant (src/SingleMassOscillator.py:17-48)
m, c1, c2, d1, d2 = 0.2, 5.0, 2.0, 0.4, 0.4


def F_spring(x):
    return c1 * x + c2 * x ** 3


def F_damper(dx):
    return d1 * dx * (1 / (1 + d2 * dx * np.tanh(dx)))


def dx(x, F, F_sd, m=m):
    """time derivative of [position, velocity]"""
    return np.hstack([x[1], (F - F_sd) / m])


def f_x(x, F, F_sd, dt):
    return rk4_step(lambda s: dx(s, F, F_sd), x, dt)


def f_y(x):
    return x[0]


# ---- GP basis and prior (:54-69)
N_basis_fcn = 41
basis_fcn, sd = generate_Hilbert_BasisFunction(num_fcn=N_basis_fcn, domain_boundary=np.array([[-7.5, 7.5], [-7.5, 7.5]]),
                                               lengthscale=7.5 * 2 / N_basis_fcn, scale=100)
GP_prior = prior_mniw_2naturalPara(np.zeros((1, N_basis_fcn)), np.diag(sd), np.eye(1), 3)

# ---- simulation set-up (:75-97)
N_particles, N_PGAS_iter = 200, 800
t_end, dt, forget_factor = 15.0, 0.02, 0.999
time = np.arange(0.0, t_end, dt)
steps = len(time)
key = _random.key(12345678)
x0 = np.array([0.0, 0.0])
P0 = np.diag([1e-4, 1e-4])
P0_F = np.diag([1e-12])
R = np.array([[1e-3]])
Q = np.diag([5e-8, 5e-9])
F_ext = np.ones((steps,)) * 9.81 * m
F_ext[int(t_end / (3 * dt)):] = 0
F_ext[int(2 * t_end / (3 * dt)):] = -9.81 * m

SMO_SSM = StateSpaceModel(process_noise=Q, output_noise=R,
                          transition_model=lambda state, input, *int_var: f_x(state, input, int_var[0], dt),
                          output_model=lambda state, input, *int_var: f_y(state))


def SingleMassOscillator_simulation(key):
    X, Y, H = simulate(SMO_SSM, key, x0, F_ext, lambda x, u: [F_spring(x[0]) + F_damper(x[1])], np.sqrt(np.diag(R)))
    return X, Y[:, 0], H[0]


key, key_sim = _random.split(key)
X, Y, F_sd = SingleMassOscillator_simulation(key_sim)

_common = dict(observations=Y, inputs=F_ext, SSM=SMO_SSM, init_state_mean=x0, init_state_cov=P0,
               init_int_var_mean=[np.array([0.0])], init_int_var_cov=[P0_F], GP_prior=[GP_prior],
               basis_fcn=[lambda state, input: basis_fcn(state)])
SMO_Algorithm1 = Algorithm1(N_samples=N_particles, forgetting_factor=forget_factor, **_common)
SMO_Algorithm2 = Algorithm2(N_samples=N_particles, N_iterations=N_PGAS_iter, **_common)
