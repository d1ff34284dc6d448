"""State-space-model wrapper (reference src/StateSpaceModel.py:8-87): the seam through which user
models reach the marginalised filters.

Constructor and method names follow the reference.  `transition_model(state, input, *int_vars)` and
`output_model(state, input, *int_vars)` stay ordinary Python callables written with numpy; the
filters never call them per particle — `tables()` traces them once per time step (tracing.py) and
hands per-step coefficient tables to the CUDA kernels (csrc/marginal.cu).  A callable outside the
compiled-in family (affine in state / interface variables given the input, optional tanh output
link) is traced once more by `programs()`, with the input symbolic as well, into a postfix expression
program that the kernels interpret per particle (the model plug-in); what is
outside that instruction set too raises at construction of the algorithm object; there is no host
fallback.

`draw_state` / `log_likelihood` / `transition_mdl` / `output_mdl` on CONCRETE numbers evaluate the
user callable directly; the example modules use them once at import time to synthesise their
ground-truth data (src/SingleMassOscillator.py:117-137), which is set-up, not the hot path.
"""
import numpy as np

from . import random as _random
from . import tracing as _tr


class StateSpaceModel:
    def __init__(self, process_noise, output_noise, transition_model, output_model):
        self.process_noise = np.atleast_2d(np.asarray(process_noise, dtype=np.float64))
        self.output_noise = np.atleast_2d(np.asarray(output_noise, dtype=np.float64))
        self.transition_model = transition_model
        self.output_model = output_model
        self.is_deterministic = bool(np.all(self.process_noise == 0))          # src/StateSpaceModel.py:30

    def transition_mdl(self, state, input, *int_variables):
        return self.transition_model(state, input, *int_variables)

    def output_mdl(self, state, input, *int_variables):
        return self.output_model(state, input, *int_variables)

    def draw_state(self, key, state, input, *int_variables):
        """src/StateSpaceModel.py:56-73 for one concrete state (data synthesis at import time)."""
        new_state = np.asarray(self.transition_mdl(np.asarray(state, dtype=np.float64), input, *int_variables), dtype=np.float64)
        if self.is_deterministic:
            return new_state
        z = _random.normal(key, np.shape(state))
        return new_state + np.linalg.cholesky(self.process_noise) @ z

    def log_likelihood(self, observation, state, input, *int_variables):
        """src/StateSpaceModel.py:75-87 for one concrete state."""
        out = np.atleast_1d(np.asarray(self.output_mdl(state, input, *int_variables), dtype=np.float64))
        L = np.linalg.cholesky(self.output_noise)
        e = np.linalg.solve(L, np.atleast_1d(observation) - out)
        return float(-0.5 * e @ e - 0.5 * len(out) * np.log(2 * np.pi) - np.sum(np.log(np.diag(L))))

    # ------------------------------------------------------------------ tracing
    def tables(self, inputs, n_x, n_xi):
        """Per-time-step coefficient tables: trans (T, n_x, n_x+G+1), outp (T, n_y, n_x+G+1), out_link."""
        T = inputs.shape[0]
        G = len(n_xi)
        if any(k != 1 for k in n_xi):
            raise NotImplementedError("interface variables must be scalar (n_xi = 1), as in every reference example")
        state, xis = _tr.variables(n_x, n_xi)
        trans = np.zeros((T, n_x, n_x + G + 1))
        outp = None
        link = None
        for t in range(T):
            f = self.transition_model(state, inputs[t], *xis)
            if not isinstance(f, _tr.Expr):
                f = _tr.hstack([f]) if isinstance(f, (list, tuple)) else f
            if not isinstance(f, _tr.Expr) or f.link is not None or f.b.shape[0] != n_x:
                raise TypeError("transition_model must be affine in (state, interface variables) and return n_x values")
            trans[t, :, :n_x + G] = f.A
            trans[t, :, n_x + G] = f.b
            g = self.output_model(state, inputs[t], *xis)
            if not isinstance(g, _tr.Expr):
                raise TypeError("output_model must depend on the state (got a constant)")
            if g.link not in (None, "tanh") or not (np.all(g.p == 1.0) and np.all(g.q == 0.0)):
                raise TypeError("output_model must be an affine map with an optional tanh link")
            if outp is None:
                outp = np.zeros((T, g.b.shape[0], n_x + G + 1))
                link = g.link
            if g.link != link:
                raise TypeError("output_model changes its link over time")
            outp[t, :, :n_x + G] = g.A
            outp[t, :, n_x + G] = g.b
        return trans, outp, link

    def programs(self, inputs, n_x, n_xi):
        """Model plug-in (include/pgas_b200.h: pgas_marg_program): transition_model and output_model as postfix
        expression programs over ([state; xi_1 .. xi_G], inputs[t]).  Returns ((ops, consts), (ops, consts), n_y);
        raises TypeError when a callable uses something outside the instruction set."""
        from . import models as _md
        if any(k != 1 for k in n_xi):
            raise NotImplementedError("interface variables must be scalar (n_xi = 1), as in every reference example")
        state, xis, u = program_variables(n_x, len(n_xi), inputs)
        out = []
        for name, fn, n_out in (("transition_model", self.transition_model, n_x), ("output_model", self.output_model, None)):
            try:
                val = fn(state, u, *xis)
                if isinstance(val, (list, tuple)):
                    val = _md.hstack(list(val))
                if not isinstance(val, _md.Sym):
                    raise TypeError("the result does not depend on the state")
                prog = _md.compile_program(val)
            except TypeError as e:
                raise TypeError(f"{name} is outside the supported model families (affine in state and interface variables with an "
                                f"optional tanh output link, or an expression of numpy arithmetic / elementary functions): {e}") from e
            if n_out is not None and len(val) != n_out:
                raise TypeError(f"{name} returns {len(val)} values, expected {n_out}")
            out.append((prog, len(val)))
        return out[0][0], out[1][0], out[1][1]


def program_variables(n_x, G, inputs):
    """Symbolic (state, [xi_g], input) of the model plug-in: operand k of PUSH_X is state component k for k < n_x and interface
    variable k - n_x otherwise; a one-dimensional `inputs` array hands the callables a scalar input, as inputs[t] would."""
    from . import models as _md
    inputs = np.asarray(inputs)
    state = _md.Sym([("x", k) for k in range(n_x)])
    xis = [_md.Sym([("x", n_x + g)]) for g in range(G)]
    u = _md.Sym([("u", 0)], scalar=True) if inputs.ndim == 1 else _md.Sym([("u", k) for k in range(inputs.shape[1])])
    return state, xis, u
