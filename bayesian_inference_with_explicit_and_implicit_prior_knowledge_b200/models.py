"""Model descriptors: how the reference's user callables reach the CUDA kernels.

The reference passes arbitrary Python callables `basis_fcn(state, input)` and
`likelihood_fcn(obs, state, input)` (src/PGAS.py:24-43) and lets JAX trace them.  A persistent
CUDA kernel cannot call Python, so the callables are TRACED ONCE on the host with symbolic
affine values and reduced to a parameter block of a compiled-in family:

  basis_fcn      -> Hilbert-space GP basis of an affine map of (state, input)
                    (every shipped Theta-conditioned model: src/EMPS.py:110-113,
                    src/Toy_Example.py:146) or of the vehicle slip angles (src/Vehicle.py:50-57);
  likelihood_fcn -> Gaussian log-density of obs around an affine map of the state
                    (src/EMPS.py:250-252, src/Toy_Example.py:142-144).

A `basis_fcn` whose GP-input map is not affine (np.sin(state[0]), products of state components, ...) is traced a second
time with symbolic EXPRESSIONS (`Sym`) and shipped to the kernels as a postfix program they interpret per particle
(include/pgas_b200.h: PGAS_MAP_PROGRAM) — the model plug-in of SURVEY.md 8f item 2: arithmetic, powers and the elementary
functions of numpy, no data-dependent Python control flow.  Anything else raises at construction time; nothing falls back to
the host.
"""
import ctypes as C

import numpy as np

from . import _lib


# ----------------------------------------------------------------------------- affine tracer
class Affine:
    """Symbolic vector  A @ v + b  over the variables v = [state; input]."""
    __array_priority__ = 1000

    def __init__(self, A, b, scalar=False):
        self.A = np.atleast_2d(np.asarray(A, dtype=np.float64))
        self.b = np.atleast_1d(np.asarray(b, dtype=np.float64))
        self.scalar = scalar

    @property
    def shape(self):
        return () if self.scalar else (self.A.shape[0],)

    def __len__(self):
        return self.A.shape[0]

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            return Affine(self.A[idx:idx + 1] if idx != -1 else self.A[-1:], self.b[[idx]], scalar=True)
        return Affine(self.A[idx], self.b[idx])

    @staticmethod
    def _const(x, n):
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 0:
            return np.full(n, float(x))
        return x.ravel()

    def __add__(self, o):
        if isinstance(o, Affine):
            return Affine(self.A + o.A, self.b + o.b, self.scalar and o.scalar)
        return Affine(self.A, self.b + self._const(o, len(self)), self.scalar and np.ndim(o) == 0)

    __radd__ = __add__

    def __neg__(self):
        return Affine(-self.A, -self.b, self.scalar)

    def __sub__(self, o):
        return self + (-o if isinstance(o, Affine) else -np.asarray(o, dtype=np.float64))

    def __rsub__(self, o):
        return (-self) + o

    def __mul__(self, o):
        if isinstance(o, Affine):
            raise TypeError("product of two state-dependent values is not affine")
        c = self._const(o, len(self))
        return Affine(self.A * c[:, None], self.b * c, self.scalar and np.ndim(o) == 0)

    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Affine):
            raise TypeError("division by a state-dependent value is not affine")
        c = self._const(o, len(self))
        return Affine(self.A / c[:, None], self.b / c, self.scalar and np.ndim(o) == 0)

    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        if method != "__call__":
            return NotImplemented
        a, b = (inputs + (None,))[:2]
        if ufunc is np.add:
            return a + b if isinstance(a, Affine) else b + a
        if ufunc is np.subtract:
            return a - b if isinstance(a, Affine) else (-b) + a
        if ufunc is np.multiply:
            return a * b if isinstance(a, Affine) else b * a
        if ufunc in (np.divide, np.true_divide) and isinstance(a, Affine):
            return a / b
        if ufunc is np.negative:
            return -a
        raise TypeError(f"{ufunc.__name__} of a state-dependent value is outside the compiled-in (affine) model families")

    def __array_function__(self, func, types, args, kwargs):
        if func in (np.hstack, np.concatenate):
            parts = [p if isinstance(p, Affine) else Affine(np.zeros((np.size(p), self.A.shape[1])), np.ravel(p))
                     for p in args[0]]
            return Affine(np.vstack([p.A for p in parts]), np.concatenate([p.b for p in parts]))
        if func in (np.atleast_1d, np.ravel, np.asarray, np.squeeze):
            return Affine(self.A, self.b, scalar=(func is np.squeeze and len(self) == 1))
        raise TypeError(f"numpy.{func.__name__} of a state-dependent value is not supported by the model tracer")


# ----------------------------------------------------------------------------- expression tracer (model plug-in)
_UNARY = {np.negative: "NEG", np.sin: "SIN", np.cos: "COS", np.tan: "TAN", np.tanh: "TANH", np.arctan: "ATAN", np.exp: "EXP",
          np.log: "LOG", np.sqrt: "SQRT", np.abs: "ABS", np.absolute: "ABS", np.fabs: "ABS"}
_BINARY = {np.add: "ADD", np.subtract: "SUB", np.multiply: "MUL", np.divide: "DIV", np.true_divide: "DIV", np.power: "POW",
           np.arctan2: "ATAN2"}


class Sym:
    """Symbolic vector of expression trees over (state, input[, observation]).  A node is ("x", k) | ("u", k) | ("y", k) |
    ("c", value) | (opcode name, child[, child])."""
    __array_priority__ = 1000

    def __init__(self, nodes, scalar=False):
        self.nodes, self.scalar = list(nodes), scalar

    @property
    def shape(self):
        return () if self.scalar else (len(self.nodes),)

    def __len__(self):
        return len(self.nodes)

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            return Sym([self.nodes[idx]], scalar=True)
        return Sym(list(np.asarray(self.nodes, dtype=object)[idx]) if not isinstance(idx, slice) else self.nodes[idx])

    def __iter__(self):
        return (self[i] for i in range(len(self.nodes)))

    @staticmethod
    def _lift(o, n):
        if isinstance(o, Sym):
            nodes = o.nodes
        elif isinstance(o, Affine):
            raise TypeError("mixing affine and expression tracers")
        else:
            nodes = [("c", float(v)) for v in np.ravel(np.asarray(o, dtype=np.float64))]
        if len(nodes) == n:
            return nodes
        if len(nodes) == 1:
            return nodes * n
        raise TypeError(f"shape mismatch: {len(nodes)} vs {n}")

    def _binary(self, name, o, swap=False):
        n = max(len(self), len(o) if isinstance(o, Sym) else np.size(o))
        a, b = self._lift(self, n), self._lift(o, n)
        if swap:
            a, b = b, a
        out = []
        for x, y in zip(a, b):
            if x[0] == "c" and y[0] == "c":                          # constant folding
                fold = {"ADD": np.add, "SUB": np.subtract, "MUL": np.multiply, "DIV": np.divide, "POW": np.power, "ATAN2": np.arctan2}[name]
                out.append(("c", float(fold(x[1], y[1]))))
            else:
                out.append((name, x, y))
        return Sym(out, self.scalar and (o.scalar if isinstance(o, Sym) else np.ndim(o) == 0))

    def _unary(self, name):
        return Sym([(name, x) for x in self.nodes], self.scalar)

    def __add__(self, o): return self._binary("ADD", o)
    def __radd__(self, o): return self._binary("ADD", o, swap=True)
    def __sub__(self, o): return self._binary("SUB", o)
    def __rsub__(self, o): return self._binary("SUB", o, swap=True)
    def __mul__(self, o): return self._binary("MUL", o)
    def __rmul__(self, o): return self._binary("MUL", o, swap=True)
    def __truediv__(self, o): return self._binary("DIV", o)
    def __rtruediv__(self, o): return self._binary("DIV", o, swap=True)
    def __pow__(self, o): return self._binary("POW", o)
    def __rpow__(self, o): return self._binary("POW", o, swap=True)
    def __neg__(self): return self._unary("NEG")
    def __pos__(self): return self
    def __abs__(self): return self._unary("ABS")

    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        if method != "__call__":
            return NotImplemented
        if ufunc in _UNARY:
            return inputs[0]._unary(_UNARY[ufunc])
        if ufunc is np.square:
            return inputs[0]._binary("MUL", inputs[0])
        if ufunc in _BINARY:
            a, b = inputs
            return a._binary(_BINARY[ufunc], b) if isinstance(a, Sym) else b._binary(_BINARY[ufunc], a, swap=True)
        raise TypeError(f"numpy.{ufunc.__name__} of a state-dependent value is not in the expression-program instruction set "
                        "(arithmetic, power, sin cos tan tanh arctan arctan2 exp log sqrt abs)")

    def __array_function__(self, func, types, args, kwargs):
        if func in (np.hstack, np.concatenate):
            nodes = []
            for part in args[0]:
                nodes += part.nodes if isinstance(part, Sym) else [("c", float(v)) for v in np.ravel(part)]
            return Sym(nodes)
        if func in (np.atleast_1d, np.ravel, np.asarray, np.squeeze):
            return Sym(self.nodes, scalar=(func is np.squeeze and len(self) == 1))
        raise TypeError(f"numpy.{func.__name__} of a state-dependent value is not supported by the model tracer")


def compile_program(sym):
    """Postfix program (ops, consts) of a Sym vector; after the last instruction the stack holds the components in order."""
    ops, consts = [], []

    def depth_of(node):
        if node[0] in ("x", "u", "y", "c"):
            return 1
        ds = [depth_of(c) for c in node[1:]]
        return ds[0] if len(ds) == 1 else max(ds[0], ds[1] + 1)

    def emit(node):
        kind = node[0]
        if kind == "x":
            ops.append(_lib.OPS["PUSH_X"] | (node[1] << 8))
        elif kind == "u":
            ops.append(_lib.OPS["PUSH_U"] | (node[1] << 8))
        elif kind == "y":
            ops.append(_lib.OPS["PUSH_Y"] | (node[1] << 8))
        elif kind == "c":
            if node[1] not in consts:
                consts.append(node[1])
            ops.append(_lib.OPS["PUSH_C"] | (consts.index(node[1]) << 8))
        else:
            for child in node[1:]:
                emit(child)
            ops.append(_lib.OPS[kind])
    for i, node in enumerate(sym.nodes):
        if i + depth_of(node) > _lib.PGAS_PROG_STACK:
            raise TypeError(f"GP-input expression {i} needs more than {_lib.PGAS_PROG_STACK} operands on the stack")
        emit(node)
    if len(ops) > _lib.PGAS_MAX_PROG or len(consts) > _lib.PGAS_MAX_PROG:
        raise TypeError(f"GP-input map compiles to {len(ops)} instructions / {len(consts)} constants (limit {_lib.PGAS_MAX_PROG})")
    return ops, consts


def run_program(ops, consts, x, u, y=None):
    """Host mirror of the device interpreter (csrc/basis_eval.cuh: pgas_run_program), for tests and debugging.  x, u, y may be
    NumPy arrays or torch tensors with a trailing component axis: the operations broadcast (the final log-likelihood table of
    PGAS.__call__ evaluates a likelihood program on whole trajectories this way)."""
    if hasattr(x, "is_cuda") or hasattr(y, "is_cuda"):
        return _run_program_torch(ops, consts, x, u, y)
    names = {v: k for k, v in _lib.OPS.items()}
    una = {"NEG": np.negative, "SIN": np.sin, "COS": np.cos, "TAN": np.tan, "TANH": np.tanh, "ATAN": np.arctan, "EXP": np.exp, "LOG": np.log,
           "SQRT": np.sqrt, "ABS": np.abs}
    bina = {"ADD": np.add, "SUB": np.subtract, "MUL": np.multiply, "DIV": np.divide, "POW": np.power, "ATAN2": np.arctan2}
    st = []
    for ins in ops:
        name, arg = names[ins & 0xFF], ins >> 8
        if name == "PUSH_X":
            st.append(float(x[arg]))
        elif name == "PUSH_U":
            st.append(float(u[arg]))
        elif name == "PUSH_Y":
            st.append(float(y[arg]))
        elif name == "PUSH_C":
            st.append(consts[arg])
        elif name in bina:
            b = st.pop(); a = st.pop(); st.append(float(bina[name](a, b)))
        else:
            st.append(float(una[name](st.pop())))
    return np.array(st)


def _run_program_torch(ops, consts, x, u, y):
    import torch
    names = {v: k for k, v in _lib.OPS.items()}
    una = {"NEG": torch.neg, "SIN": torch.sin, "COS": torch.cos, "TAN": torch.tan, "TANH": torch.tanh, "ATAN": torch.atan, "EXP": torch.exp,
           "LOG": torch.log, "SQRT": torch.sqrt, "ABS": torch.abs}
    bina = {"ADD": torch.add, "SUB": torch.sub, "MUL": torch.mul, "DIV": torch.div, "POW": torch.pow, "ATAN2": torch.atan2}
    ref = x if hasattr(x, "is_cuda") else y
    st = []
    for ins in ops:
        name, arg = names[ins & 0xFF], ins >> 8
        if name == "PUSH_X":
            st.append(x[..., arg])
        elif name == "PUSH_U":
            st.append(u[..., arg])
        elif name == "PUSH_Y":
            st.append(y[..., arg])
        elif name == "PUSH_C":
            st.append(torch.tensor(consts[arg], dtype=ref.dtype, device=ref.device))
        elif name in bina:
            b = st.pop(); a = st.pop(); st.append(bina[name](a, b))
        else:
            st.append(una[name](st.pop()))
    return st


def hstack(parts):
    """stand-in for jnp.hstack usable on traced values"""
    if any(isinstance(p, (Affine, Sym)) for p in parts):
        ref = next(p for p in parts if isinstance(p, (Affine, Sym)))
        return ref.__array_function__(np.hstack, (), (parts,), {})
    return np.hstack(parts)


def _tracers(n_x, n_u):
    eye = np.eye(n_x + n_u)
    return Affine(eye[:n_x], np.zeros(n_x)), Affine(eye[n_x:], np.zeros(n_u))


# ----------------------------------------------------------------------------- basis
class BasisExpr:
    """Result of applying a HilbertBasis to a traced value: hgp(Az v + bz)."""

    def __init__(self, hgp, Az, bz):
        self.hgp, self.Az, self.bz = hgp, np.atleast_2d(Az), np.atleast_1d(bz)
        self.map_kind = _lib.MAP_AFFINE
        self.slip = (0.0, 0.0)

    def __len__(self):
        return self.hgp.M


class VehicleSlipBasis:
    """basis over the slip angles (alpha_f, alpha_r) of src/Vehicle.py:50-57 (D = 2)."""

    def __init__(self, hgp, l_f, l_r):
        assert hgp.D == 2, "slip-angle basis is two-dimensional (front, rear)"
        self.hgp, self.slip = hgp, (float(l_f), float(l_r))
        self.map_kind = _lib.MAP_VEHICLE_SLIP
        self.Az, self.bz = np.zeros((2, 4)), np.zeros(2)

    def __len__(self):
        return self.hgp.M


class ProgramBasis:
    """basis over a GP-input map given as an expression program (model plug-in)."""

    def __init__(self, hgp, sym):
        if len(sym) != hgp.D:
            raise ValueError(f"basis expects {hgp.D} inputs, traced value has {len(sym)}")
        self.hgp, self.sym = hgp, sym
        self.ops, self.consts = compile_program(sym)
        self.map_kind = _lib.MAP_PROGRAM
        self.slip = (0.0, 0.0)
        self.Az, self.bz = np.zeros((hgp.D, _lib.PGAS_MAX_NX + _lib.PGAS_MAX_NU)), np.zeros(hgp.D)

    def __len__(self):
        return self.hgp.M


def trace_basis(basis_fcn, n_x, n_u):
    """Reduce `basis_fcn(state, input)` to a descriptor (BasisExpr / VehicleSlipBasis)."""
    if isinstance(basis_fcn, (BasisExpr, VehicleSlipBasis, ProgramBasis)):
        return basis_fcn
    from .BasisFunctions import HilbertBasis
    if isinstance(basis_fcn, HilbertBasis):
        if basis_fcn.D != n_x:
            raise ValueError("a bare HilbertBasis as basis_fcn must have D == n_x")
        return BasisExpr(basis_fcn, np.hstack([np.eye(n_x), np.zeros((n_x, n_u))]), np.zeros(n_x))
    s, u = _tracers(n_x, n_u)
    try:
        out = basis_fcn(s, u if n_u > 0 else np.zeros(0))
    except TypeError:
        # not affine: trace again with symbolic expressions -> expression program interpreted by the kernels (model plug-in)
        xs = Sym([("x", k) for k in range(n_x)])
        us = Sym([("u", k) for k in range(n_u)]) if n_u > 0 else np.zeros(0)
        try:
            out = basis_fcn(xs, us)
        except TypeError as e:
            raise TypeError("basis_fcn is outside the supported model families (Hilbert-space GP basis of an affine map, of the "
                            f"vehicle slip angles, or of an expression of numpy arithmetic / elementary functions): {e}") from e
    if not isinstance(out, (BasisExpr, VehicleSlipBasis, ProgramBasis)):
        raise TypeError("basis_fcn must return the value of a generate_Hilbert_BasisFunction basis")
    return out


# ----------------------------------------------------------------------------- likelihood
class GaussianLikelihood:
    """likelihood_fcn(obs, state, input) = log N(obs; H state + h0, R)."""

    def __init__(self, H, h0, R):
        self.H = np.atleast_2d(np.asarray(H, dtype=np.float64))
        self.h0 = np.atleast_1d(np.asarray(h0, dtype=np.float64))
        self.R = np.atleast_2d(np.asarray(R, dtype=np.float64))
        assert self.R.shape == (self.H.shape[0],) * 2

    def logpdf_torch(self, obs, states):
        """obs (..., n_y), states (..., n_x) CUDA tensors -> log-density (...); library ops, used only for
        the final log-likelihood table of PGAS.__call__ (src/PGAS.py:383-392), not in the sweep."""
        import torch
        H = torch.as_tensor(self.H, device=states.device)
        h0 = torch.as_tensor(self.h0, device=states.device)
        Lr = torch.linalg.cholesky(torch.as_tensor(self.R, device=states.device))
        d = obs - (states @ H.T + h0)
        e = torch.linalg.solve_triangular(Lr, d.unsqueeze(-1), upper=False).squeeze(-1)
        n = self.R.shape[0]
        return -0.5 * (e * e).sum(-1) - 0.5 * n * np.log(2 * np.pi) - torch.log(torch.diagonal(Lr)).sum()


def gaussian_likelihood(f_y, R, n_x=None):
    """Build the likelihood descriptor from an (affine) output map f_y(state) and covariance R —
    the shape every shipped `likelihood_fcn` lambda has (src/EMPS.py:250-252)."""
    R = np.atleast_2d(np.asarray(R, dtype=np.float64))

    def build(nx):
        s, _ = _tracers(nx, 0)
        out = f_y(s)
        if not isinstance(out, Affine):
            raise TypeError("output map must be affine in the state")
        return GaussianLikelihood(out.A[:, :nx], out.b, R)
    if n_x is not None:
        return build(n_x)
    lazy = _LazyLikelihood(build)
    return lazy


class _LazyLikelihood:
    def __init__(self, build):
        self._build = build

    def resolve(self, n_x):
        return self._build(n_x)


class ProgramLikelihood:
    """likelihood_fcn(obs, state, input) as an expression program that leaves the log-density (model plug-in; pgas_b200.h:
    lik_prog_*).  H / h0 / R are placeholders of the right shape for the Gaussian fields of pgas_model_params."""

    def __init__(self, sym, n_x, n_y):
        if len(sym.nodes) != 1:
            raise TypeError(f"likelihood_fcn must return one log-density, got {len(sym.nodes)} values")
        self.sym = sym
        self.ops, self.consts = compile_program(sym)
        self.H, self.h0, self.R = np.zeros((n_y, n_x)), np.zeros(n_y), np.eye(n_y)

    def logpdf_torch(self, obs, states, inputs=None):
        (out,) = run_program(self.ops, self.consts, states, inputs, obs)
        import torch
        return torch.broadcast_to(out, torch.broadcast_shapes(states.shape[:-1], obs.shape[:-1]))


class ObservationMarker:
    """stands for `obs` while a reference-style likelihood_fcn(obs, state, input) is traced"""


class GaussianLogpdfTrace:
    """value of stats.multivariate_normal.logpdf(obs, mean=<affine in the state>, cov=R) under tracing; np.squeeze passes it through"""

    def __init__(self, mean, cov):
        self.mean, self.cov = mean, np.atleast_2d(np.asarray(cov, dtype=np.float64))

    def __array_function__(self, func, types, args, kwargs):
        if func in (np.squeeze, np.asarray, np.atleast_1d, np.ravel):
            return self
        raise TypeError(f"numpy.{func.__name__} of a traced log-density is not supported")

    def squeeze(self, *a, **k):
        return self


def resolve_likelihood(lik, n_x, n_u=0, n_y=None):
    if isinstance(lik, (GaussianLikelihood, ProgramLikelihood)):
        return lik
    if isinstance(lik, _LazyLikelihood):
        return lik.resolve(n_x)
    if callable(lik):
        # the reference's own form (src/EMPS.py:250-252): lambda obs, state, input: squeeze(multivariate_normal.logpdf(obs, mean=f_y(state), cov=R))
        # written against this package's `stats` module; traced once with a symbolic state
        s, u = _tracers(n_x, n_u)
        try:
            out = lik(ObservationMarker(), s, u if n_u > 0 else np.zeros(0))
        except TypeError:
            out = None
        if isinstance(out, GaussianLogpdfTrace) and not np.any(out.mean.A[:, n_x:] != 0):
            return GaussianLikelihood(out.mean.A[:, :n_x], out.mean.b, out.cov)
        if n_y is not None:
            # outside the Gaussian-of-an-affine-map family: the whole log-density as an expression program over
            # (state, input, observation) — the model plug-in; such a model runs the fused sweep kernel
            xs = Sym([("x", k) for k in range(n_x)])
            us = Sym([("u", k) for k in range(n_u)]) if n_u > 0 else np.zeros(0)
            ys = Sym([("y", k) for k in range(n_y)])
            try:
                out = lik(ys, xs, us)
            except TypeError as e:
                raise TypeError(f"likelihood_fcn is outside the supported families (Gaussian observation of an affine output map, or an "
                                f"expression of numpy arithmetic / elementary functions of observation, state and input): {e}") from e
            if isinstance(out, Sym):
                return ProgramLikelihood(out, n_x, n_y)
    raise TypeError("likelihood_fcn must be a Gaussian observation of an affine output map: either models.gaussian_likelihood(f_y, R) "
                    "or the reference's lambda written with this package's stats.multivariate_normal.logpdf(obs, mean=f_y(state), cov=R); "
                    "other Python callables cannot run inside the CUDA sweep")


# ----------------------------------------------------------------------------- device model
class DeviceModel:
    """Owns a pgas_model handle (include/pgas_b200.h: pgas_model_create)."""

    def __init__(self, observations, inputs, m0, P0, likelihood, basis, flags=0):
        _lib.require_cuda()
        obs = np.ascontiguousarray(np.asarray(observations, dtype=np.float64))
        self.T = obs.shape[0]
        obs = obs.reshape(self.T, -1)
        inp = np.asarray(inputs, dtype=np.float64)
        inp = inp.reshape(self.T, -1) if inp.size else np.zeros((self.T, 0))
        inp = np.ascontiguousarray(inp)
        m0 = np.atleast_1d(np.asarray(m0, dtype=np.float64))
        P0 = np.atleast_2d(np.asarray(P0, dtype=np.float64))
        self.n_x, self.n_y, self.n_u = m0.shape[0], obs.shape[1], inp.shape[1]
        self.basis = trace_basis(basis, self.n_x, self.n_u)
        self.likelihood = resolve_likelihood(likelihood, self.n_x, self.n_u, self.n_y)
        hgp = self.basis.hgp
        self.M, self.D = hgp.M, hgp.D
        self.flags = int(flags)
        if self.likelihood.H.shape != (self.n_y, self.n_x):
            raise ValueError(f"likelihood maps to {self.likelihood.H.shape[0]} outputs, observations have {self.n_y}")
        p = _lib.ModelParams()
        p.n_x, p.n_y, p.n_u, p.D, p.M, p.T = self.n_x, self.n_y, self.n_u, self.D, self.M, self.T
        self._freq = np.ascontiguousarray(hgp.freq, dtype=np.int32)
        p.freq = self._freq.ctypes.data_as(C.POINTER(C.c_int32))
        p.idx_start, p.idx_step = hgp.idx_start, hgp.idx_step
        for d in range(self.D):
            p.center[d] = hgp.center[d]
            p.half_width[d] = hgp.half_width[d]
            p.bz[d] = self.basis.bz[d]
            for k in range(self.n_x + self.n_u):
                p.Az[d][k] = self.basis.Az[d, k] if k < self.basis.Az.shape[1] else 0.0
        p.map_kind = self.basis.map_kind
        p.slip_lf, p.slip_lr = self.basis.slip
        if isinstance(self.basis, ProgramBasis):
            p.prog_len = len(self.basis.ops)
            for i, ins in enumerate(self.basis.ops):
                p.prog_op[i] = ins
            for i, c in enumerate(self.basis.consts):
                p.prog_const[i] = c
        if isinstance(self.likelihood, ProgramLikelihood):
            p.lik_prog_len = len(self.likelihood.ops)
            for i, ins in enumerate(self.likelihood.ops):
                p.lik_prog_op[i] = ins
            for i, c in enumerate(self.likelihood.consts):
                p.lik_prog_const[i] = c
        for r in range(self.n_y):
            p.h0[r] = self.likelihood.h0[r]
            for k in range(self.n_x):
                p.H[r][k] = self.likelihood.H[r, k]
            for c in range(self.n_y):
                p.R[r][c] = self.likelihood.R[r, c]
        self._obs, self._inp = obs, inp
        p.observations = obs.ctypes.data_as(C.POINTER(C.c_double))
        p.inputs = inp.ctypes.data_as(C.POINTER(C.c_double)) if self.n_u else None
        for i in range(self.n_x):
            p.m0[i] = m0[i]
            for j in range(self.n_x):
                p.P0[i][j] = P0[i, j]
        p.flags = self.flags
        h = C.c_void_p()
        _lib.check(_lib.lib().pgas_model_create(C.byref(p), C.byref(h)))
        self.handle = h
        self.jmax = _lib.lib().pgas_model_jmax(h)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().pgas_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
