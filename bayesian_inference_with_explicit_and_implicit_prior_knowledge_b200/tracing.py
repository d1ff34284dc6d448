"""Symbolic tracer for the StateSpaceModel callables of the marginalised filters.

The reference hands arbitrary Python callables to `StateSpaceModel` / `Algorithm1`
(src/StateSpaceModel.py:19-30, src/Algorithm1.py:27-40) and lets JAX trace them.  A persistent CUDA
kernel cannot call Python, so each callable is traced HERE, once per time step (the input u_t is a
concrete number there), with symbolic values for the state and the interface variables, and reduced
to the coefficient tables of the compiled-in family (include/pgas_b200.h, group B):

    transition   x' = A_t x + B_t xi + c_t
    output       y  = link(C_t x + D_t xi + e_t)          link in {identity, tanh}
    GP input     z  = p_t * link(a_t . x + b_t) + q_t     link in {identity, atan}

An `Expr` is a vector  p * link(A v + b) + q  over the variables v = [state; xi_1 .. xi_G].  Without
a link it supports affine arithmetic; `np.arctan` / `np.tanh` attach a link, after which only
element-wise scaling and shifting by constants is allowed.  Anything else raises TypeError — there is
no host fallback.
"""
import numpy as np


class Expr:
    __array_priority__ = 1000

    def __init__(self, A, b, link=None, p=None, q=None, scalar=False):
        self.A = np.atleast_2d(np.asarray(A, dtype=np.float64))
        self.b = np.atleast_1d(np.asarray(b, dtype=np.float64))
        self.link = link
        k = self.b.shape[0]
        self.p = np.ones(k) if p is None else np.atleast_1d(np.asarray(p, dtype=np.float64))
        self.q = np.zeros(k) if q is None else np.atleast_1d(np.asarray(q, dtype=np.float64))
        self.scalar = bool(scalar)

    # ---- shape protocol
    def __len__(self):
        if self.scalar:
            raise TypeError("len() of a scalar traced value")
        return self.b.shape[0]

    @property
    def shape(self):
        return () if self.scalar else (self.b.shape[0],)

    @property
    def ndim(self):
        return 0 if self.scalar else 1

    @property
    def size(self):
        return self.b.shape[0]

    @property
    def T(self):
        return self

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            if len(idx) == 1 or all(i is Ellipsis for i in idx[1:]):
                idx = idx[0]
            else:
                raise TypeError("traced values are one-dimensional")
        if idx is Ellipsis:
            return self
        if isinstance(idx, (int, np.integer)):
            i = int(idx) % self.b.shape[0]
            return Expr(self.A[i:i + 1], self.b[i:i + 1], self.link, self.p[i:i + 1], self.q[i:i + 1], scalar=True)
        return Expr(self.A[idx], self.b[idx], self.link, self.p[idx], self.q[idx])

    def __iter__(self):
        if self.scalar:
            raise TypeError("iteration over a scalar traced value")
        return (self[i] for i in range(self.b.shape[0]))

    # ---- helpers
    def _const(self, o):
        o = np.asarray(o, dtype=np.float64)
        if o.ndim == 0:
            return np.full(self.b.shape[0], float(o)), True
        o = o.ravel()
        k = self.b.shape[0]
        if o.shape[0] == k:
            return o, False
        if k == 1:
            return o, False        # broadcast a length-1 traced value against a vector constant
        if o.shape[0] == 1:
            return np.full(k, float(o[0])), False
        raise TypeError(f"shape mismatch between a traced value of length {k} and a constant of length {o.shape[0]}")

    def _bcast(self, k):
        if self.b.shape[0] == k:
            return self
        if self.b.shape[0] != 1:
            raise TypeError("shape mismatch between traced values")
        rep = np.zeros(k, dtype=int)
        return Expr(self.A[rep], self.b[rep], self.link, self.p[rep], self.q[rep])

    # ---- arithmetic
    def __neg__(self):
        if self.link is None:
            return Expr(-self.A, -self.b, scalar=self.scalar)
        return Expr(self.A, self.b, self.link, -self.p, -self.q, self.scalar)

    def __pos__(self):
        return self

    def __add__(self, o):
        if isinstance(o, Expr):
            if self.link is not None or o.link is not None:
                raise TypeError("sum of two values behind a nonlinear link is outside the compiled-in model families")
            k = max(self.b.shape[0], o.b.shape[0])
            a, c = self._bcast(k), o._bcast(k)
            return Expr(a.A + c.A, a.b + c.b, scalar=self.scalar and o.scalar)
        c, sc = self._const(o)
        me = self._bcast(c.shape[0])
        if self.link is None:
            return Expr(me.A, me.b + c, scalar=self.scalar and sc)
        return Expr(me.A, me.b, self.link, me.p, me.q + c, self.scalar and sc)

    __radd__ = __add__

    def __sub__(self, o):
        return self + (-o if isinstance(o, Expr) else -np.asarray(o, dtype=np.float64))

    def __rsub__(self, o):
        return (-self) + o

    def __mul__(self, o):
        if isinstance(o, Expr):
            raise TypeError("product of two state-dependent values is outside the compiled-in (affine) model families")
        c, sc = self._const(o)
        me = self._bcast(c.shape[0])
        if self.link is None:
            return Expr(me.A * c[:, None], me.b * c, scalar=self.scalar and sc)
        return Expr(me.A, me.b, self.link, me.p * c, me.q * c, self.scalar and sc)

    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Expr):
            raise TypeError("division by a state-dependent value is outside the compiled-in (affine) model families")
        c, sc = self._const(o)
        me = self._bcast(c.shape[0])
        if self.link is None:
            return Expr(me.A / c[:, None], me.b / c, scalar=self.scalar and sc)
        return Expr(me.A, me.b, self.link, me.p / c, me.q / c, self.scalar and sc)

    def _apply_link(self, name):
        if self.link is not None:
            raise TypeError(f"{name} of a value that already sits behind {self.link} is outside the compiled-in model families")
        return Expr(self.A, self.b, name, scalar=self.scalar)

    # ---- numpy protocols (so the model modules can be written with plain numpy calls)
    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        if method != "__call__":
            return NotImplemented
        a, b = (inputs + (None,))[:2]
        if ufunc is np.add:
            return a + b if isinstance(a, Expr) else b + a
        if ufunc is np.subtract:
            return a - b if isinstance(a, Expr) else (-b) + a
        if ufunc is np.multiply:
            return a * b if isinstance(a, Expr) else b * a
        if ufunc in (np.divide, np.true_divide) and isinstance(a, Expr):
            return a / b
        if ufunc is np.negative:
            return -a
        if ufunc is np.positive:
            return a
        if ufunc is np.arctan:
            return a._apply_link("atan")
        if ufunc is np.tanh:
            return a._apply_link("tanh")
        raise TypeError(f"numpy.{ufunc.__name__} of a state-dependent value is outside the compiled-in model families "
                        "(affine maps with an optional arctan / tanh link)")

    def __array_function__(self, func, types, args, kwargs):
        if func in (np.hstack, np.concatenate):
            return hstack(args[0])
        if func in (np.atleast_1d, np.ravel, np.asarray, np.array):
            return Expr(self.A, self.b, self.link, self.p, self.q)
        if func is np.squeeze:
            return Expr(self.A, self.b, self.link, self.p, self.q, scalar=self.b.shape[0] == 1)
        if func is np.atleast_2d:
            return self
        raise TypeError(f"numpy.{func.__name__} of a state-dependent value is not supported by the model tracer")


def hstack(parts):
    """np.hstack / jnp.hstack on a mix of traced values and constants."""
    parts = list(parts)
    ref = next((p for p in parts if isinstance(p, Expr)), None)
    if ref is None:
        return np.hstack(parts)
    nv = ref.A.shape[1]
    links = {p.link for p in parts if isinstance(p, Expr)}
    if len(links) > 1:
        raise TypeError("hstack of values behind different links is outside the compiled-in model families")
    link = links.pop()
    As, bs, ps, qs = [], [], [], []
    for p in parts:
        if isinstance(p, Expr):
            As.append(p.A); bs.append(p.b); ps.append(p.p); qs.append(p.q)
        else:
            c = np.atleast_1d(np.asarray(p, dtype=np.float64)).ravel()
            As.append(np.zeros((c.shape[0], nv)))
            if link is None:
                bs.append(c); ps.append(np.ones_like(c)); qs.append(np.zeros_like(c))
            else:                       # a constant next to linked values: p = 0, q = c
                bs.append(np.zeros_like(c)); ps.append(np.zeros_like(c)); qs.append(c)
    return Expr(np.vstack(As), np.concatenate(bs), link, np.concatenate(ps), np.concatenate(qs))


def variables(n_x, n_xi):
    """Symbolic state (n_x,) and interface variables [(n_xi_g,), ...] over v = [state; xi_1; ..]."""
    nv = n_x + int(sum(n_xi))
    eye = np.eye(nv)
    state = Expr(eye[:n_x], np.zeros(n_x))
    xis, o = [], n_x
    for k in n_xi:
        xis.append(Expr(eye[o:o + k], np.zeros(k)))
        o += k
    return state, xis


class BasisCall:
    """Value of a Hilbert-space basis applied to a traced GP input: hgp(z), z an Expr of length D."""

    def __init__(self, hgp, z):
        if z.b.shape[0] != hgp.D:
            raise ValueError(f"basis expects {hgp.D} inputs, traced value has {z.b.shape[0]}")
        self.hgp, self.z = hgp, z

    def __len__(self):
        return self.hgp.M

    @property
    def shape(self):
        return (self.hgp.M,)
