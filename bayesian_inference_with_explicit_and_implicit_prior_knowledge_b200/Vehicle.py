"""Vehicle lateral-dynamics example (reference src/Vehicle.py): single-track model whose front / rear
tyre friction coefficients mu_f(alpha_f), mu_r(alpha_r) are the unknown functions, each a 1-D
Hilbert-space GP of its slip angle.  Module-level names match the reference."""
import numpy as np

from . import random as _random
from ._examples import rk4_step, simulate
from .Algorithm1 import Algorithm1
from .Algorithm2 import Algorithm2
from .BasisFunctions import generate_Hilbert_BasisFunction
from .BayesianInferrence import prior_mniw_2naturalPara
from .StateSpaceModel import StateSpaceModel

# ---- parameters (src/Vehicle.py:17-26)
m, I_zz, l_f, l_r, g, mu_x = 1720.0, 1827.5, 1.16, 1.47, 9.81, 0.9
mu, B, C, E = 0.9, 10.0, 1.9, 0.97


def f_Fz(m, l_f, l_r, g):
    """static tyre loads front / rear"""
    w = m * g / (l_f + l_r)
    return w * l_r, w * l_f


def mu_y(alpha, mu=mu, B=B, C=C, E=E):
    """magic-formula friction curve (ground truth only)"""
    ta = np.tan(alpha)
    return mu * np.sin(C * np.arctan(B * (1 - E) * ta + E * np.arctan(B * ta)))


def f_alpha(x, u, l_f=l_f, l_r=l_r):
    """slip angles; x = [yaw rate, lateral velocity], u = [steering angle, longitudinal velocity]"""
    front = (x[1] + x[0] * l_f) / u[1]
    rear = (x[1] - x[0] * l_r) / u[1]
    return u[0] - np.arctan(front), -np.arctan(rear)


def _lateral_acc(x, u, mu_yf, mu_yr, m, l_f, l_r, g, mu_x):
    F_zf, F_zr = f_Fz(m, l_f, l_r, g)
    return 1 / m * (F_zf * mu_yf * np.cos(u[0]) + F_zr * mu_yr + F_zf * mu_x * np.sin(u[0])) - u[1] * x[0]


def dx(x, u, mu_yf, mu_yr, m, I_zz, l_f, l_r, g, mu_x):
    F_zf, F_zr = f_Fz(m, l_f, l_r, g)
    ddpsi = 1 / I_zz * (l_f * F_zf * mu_yf * np.cos(u[0]) - l_r * F_zr * mu_yr + l_f * F_zf * mu_x * np.sin(u[0]))
    return np.hstack([ddpsi, _lateral_acc(x, u, mu_yf, mu_yr, m, l_f, l_r, g, mu_x)])


def f_x(x, u, mu_yf, mu_yr, dt, m=m, I_zz=I_zz, l_f=l_f, l_r=l_r, g=g, mu_x=mu_x):
    return rk4_step(lambda s: dx(s, u, mu_yf, mu_yr, m, I_zz, l_f, l_r, g, mu_x), x, dt)


def f_y(x, u, mu_yf, mu_yr, m=m, l_f=l_f, l_r=l_r, g=g, mu_x=mu_x, mu=mu, B=B, C=C, E=E):
    return np.tanh(np.hstack([x[0], _lateral_acc(x, u, mu_yf, mu_yr, m, l_f, l_r, g, mu_x)]))


# ---- GP bases and priors (:134-174): even frequencies 2,4,..,40 on [-30 deg, 30 deg]
N_basis_fcn = 20
lengthscale = 2 / 180 * np.pi
basis_fcn, spectral_density = generate_Hilbert_BasisFunction(N_basis_fcn, np.array([-30 / 180 * np.pi, 30 / 180 * np.pi]),
                                                             lengthscale, 50, idx_start=2, idx_step=2)


def basis_fcn_f(state, input):
    return basis_fcn(f_alpha(state, input)[0])


def basis_fcn_r(state, input):
    return basis_fcn(f_alpha(state, input)[1])


GP_prior_f = list(prior_mniw_2naturalPara(np.zeros((1, N_basis_fcn)), np.diag(spectral_density), np.eye(1), 0))
GP_prior_r = list(prior_mniw_2naturalPara(np.zeros((1, N_basis_fcn)), np.diag(spectral_density), np.eye(1), 0))

# ---- simulation set-up (:180-208)
N_particles, N_PGAS_iter, forget_factor = 200, 800, 0.999
dt, t_end = 0.02, 30.0
time = np.arange(0.0, t_end, dt)
steps = len(time)
key = _random.key(12345678)
x0 = np.array([0.0, 0.0])
P0 = np.diag([1e-4, 1e-4])
P0_mu = np.diag([1e-4])
R = np.diag([0.001 / 180 * np.pi, 1e-3])
Q = np.diag([1e-8, 1e-8])
ctrl_input = np.zeros((steps, 2))
ctrl_input[:, 0] = 10 / 180 * np.pi * np.sin(2 * np.pi * time / 5) * np.exp(-0.5 * (time - t_end / 2) ** 2 / (t_end / 5) ** 2)
ctrl_input[:, 1] = 11.0

Vehicle_SSM = StateSpaceModel(process_noise=Q, output_noise=R,
                              transition_model=lambda state, input, *int_var: f_x(state, input, int_var[0], int_var[1], dt),
                              output_model=lambda state, input, *int_var: f_y(state, input, int_var[0], int_var[1]))


def Vehicle_simulation(key):
    X, Y, H = simulate(Vehicle_SSM, key, x0, ctrl_input, lambda x, u: [mu_y(a) for a in f_alpha(x, u)], np.sqrt(np.diag(R)))
    return X, Y, H[0], H[1]


key, key_sim = _random.split(key)
X, Y, mu_f, mu_r = Vehicle_simulation(key_sim)

_common = dict(observations=Y, inputs=ctrl_input, SSM=Vehicle_SSM, init_state_mean=x0, init_state_cov=P0,
               init_int_var_mean=[np.array([0.0]), np.array([0.0])], init_int_var_cov=[P0_mu, P0_mu],
               GP_prior=[GP_prior_f, GP_prior_r], basis_fcn=[basis_fcn_f, basis_fcn_r])
Vehicle_Algorithm1 = Algorithm1(N_samples=N_particles, forgetting_factor=forget_factor, **_common)
Vehicle_Algorithm2 = Algorithm2(N_samples=N_particles, N_iterations=N_PGAS_iter, **_common)
