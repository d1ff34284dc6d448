"""Helpers shared by the example model modules (SingleMassOscillator, Vehicle, EMPS)."""
import numpy as np

from . import random as _random


def rk4_step(rhs, x, dt):
    """One classical Runge-Kutta step of x' = rhs(x); works on numbers and on traced values."""
    s1 = rhs(x)
    s2 = rhs(x + (dt / 2.0) * s1)
    s3 = rhs(x + (dt / 2.0) * s2)
    s4 = rhs(x + dt * s3)
    return x + (dt / 6.0) * (s1 + 2 * s2 + 2 * s3 + s4)


def simulate(ssm, key, x0, inputs, hidden_fcn, output_std):
    """Ground-truth roll-out used by the example modules at import time: x_{i} ~ SSM.draw_state(x_{i-1}, u_{i-1},
    hidden(x_{i-1}, u_{i-1})), y_i = output(x_i, u_i, hidden(x_i, u_i)) + noise.  Returns X, Y, hidden (list of arrays)."""
    steps = inputs.shape[0]
    X = np.zeros((steps, len(x0)))
    X[0] = x0
    h0 = hidden_fcn(X[0], inputs[0])
    H = [np.zeros(steps) for _ in h0]
    for j, v in enumerate(h0):
        H[j][0] = v
    y0 = np.atleast_1d(ssm.output_mdl(X[0], inputs[0], *h0))
    Y = np.zeros((steps, len(y0)))
    for i in range(1, steps):
        key, k_state, k_obs = _random.split(key, 3)
        prev = [H[j][i - 1] for j in range(len(H))]
        X[i] = ssm.draw_state(k_state, X[i - 1], inputs[i - 1], *prev)
        cur = hidden_fcn(X[i], inputs[i])
        for j, v in enumerate(cur):
            H[j][i] = v
        Y[i] = np.atleast_1d(ssm.output_mdl(X[i], inputs[i], *cur)) + _random.normal(k_obs, (Y.shape[1],)) * output_std
    return X, Y, H
