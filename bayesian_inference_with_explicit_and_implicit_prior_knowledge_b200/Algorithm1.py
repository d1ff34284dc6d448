"""Online marginalised auxiliary particle filter (reference src/Algorithm1.py), host side.

Class name, constructor keywords, call signature and the 8-tuple returned by `__call__` follow the
reference (src/Algorithm1.py:27-40, :399-492).  The T time steps run inside ONE persistent CUDA
kernel (csrc/marginal.cu: marg_sweep_kernel<0>) instead of T jitted dispatches; the user callables
are traced once per time step (tracing.py).  There is no CPU fallback.

`filter(...)` is the batched / injected-variates form used by the parity tests and the benchmark.
"""
import ctypes as C

import numpy as np

from . import _lib
from . import random as _random
from . import tracing as _tr

_LINK = {None: _lib.LINK_IDENTITY, "atan": _lib.LINK_ATAN, "tanh": _lib.LINK_TANH}


class MargDeviceModel:
    """Owns a pgas_marg_model handle (include/pgas_b200.h: pgas_marg_model_create)."""

    def __init__(self, observations, inputs, SSM, m0, P0, xi_mean, xi_cov, GP_prior, basis_fcn):
        _lib.require_cuda()
        obs = np.asarray(observations, dtype=np.float64)
        self.T = T = obs.shape[0]
        obs = np.ascontiguousarray(obs.reshape(T, -1))
        inp = np.asarray(inputs, dtype=np.float64)
        self.inputs = inp
        m0 = np.atleast_1d(np.asarray(m0, dtype=np.float64))
        P0 = np.atleast_2d(np.asarray(P0, dtype=np.float64))
        self.n_x, self.n_y, self.G = m0.shape[0], obs.shape[1], len(basis_fcn)
        if not (1 <= self.G <= _lib.PGAS_MAX_GP):
            raise ValueError(f"{self.G} GPs; this build supports 1..{_lib.PGAS_MAX_GP}")
        n_xi = [np.atleast_1d(np.asarray(m)).shape[0] for m in xi_mean]
        p = _lib.MargParams()
        dptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        keep = [obs]
        self.programs = {}             # model plug-in: which callables run as interpreted expression programs

        def set_program(q, name, prog):
            ops, consts = np.asarray(prog[0], dtype=np.int32), np.asarray(prog[1], dtype=np.float64)
            q.len, q.n_const = ops.shape[0], consts.shape[0]
            q.ops, q.consts = ops.ctypes.data_as(C.POINTER(C.c_int32)), dptr(consts)
            keep.extend([ops, consts])
            self.programs[name] = prog

        try:
            trans, outp, out_link = SSM.tables(inp, self.n_x, n_xi)
            n_out = outp.shape[1]
            keep += [trans, outp]
            p.trans, p.outp = dptr(trans), dptr(outp)
        except TypeError:
            # outside the coefficient-table families: expression programs (raises TypeError when outside those as well)
            tprog, oprog, n_out = SSM.programs(inp, self.n_x, n_xi)
            out_link = None
            set_program(p.trans_prog, "transition_model", tprog)
            set_program(p.outp_prog, "output_model", oprog)
        if n_out != self.n_y:
            raise ValueError(f"output_model returns {n_out} values, observations have {self.n_y}")
        inp2 = np.ascontiguousarray(inp.reshape(T, -1))
        p.n_u, p.inputs = inp2.shape[1], dptr(inp2)
        keep.append(inp2)
        state, _ = _tr.variables(self.n_x, n_xi)
        self.hgp, self.M = [], []
        p.n_x, p.n_y, p.n_gp, p.T = self.n_x, self.n_y, self.G, T
        for g in range(self.G):
            q = p.gp[g]
            try:
                calls = []
                for t in range(T):
                    c = basis_fcn[g](state, inp[t])
                    if not isinstance(c, _tr.BasisCall):
                        raise TypeError("basis_fcn must return the value of a generate_Hilbert_BasisFunction basis applied to "
                                        "a map of the state")
                    calls.append(c)
                hgp = calls[0].hgp
                link = calls[0].z.link
                if link not in (None, "atan") or any(c.hgp is not hgp or c.z.link != link for c in calls):
                    raise TypeError("basis_fcn: GP input must be affine or arctan-linked, with the same basis at every time step")
                if any(np.any(c.z.A[:, self.n_x:] != 0) for c in calls):
                    raise ValueError("basis_fcn must not depend on the interface variables")
            except TypeError as first:
                # GP-input map outside the affine / arctan family: one expression program over (state, inputs[t])
                from . import StateSpaceModel as _ssm
                from . import models as _md
                s_sym, _, u_sym = _ssm.program_variables(self.n_x, 0, inp)
                try:
                    c = basis_fcn[g](s_sym, u_sym)
                except TypeError as e:
                    raise TypeError(f"basis_fcn[{g}] is outside the supported families ({first}; as an expression program: {e})") from e
                if not isinstance(c, _md.ProgramBasis):
                    raise TypeError("basis_fcn must return the value of a generate_Hilbert_BasisFunction basis applied to a map of the state")
                hgp, link, calls = c.hgp, None, []
                set_program(q.prog, f"basis_fcn[{g}]", (c.ops, c.consts))
            D, M = hgp.D, hgp.M
            gp_in = np.zeros((T, D, self.n_x + 1))
            gp_post = np.zeros((T, D, 2))
            for t, c in enumerate(calls):
                gp_in[t, :, :self.n_x] = c.z.A[:, :self.n_x]
                gp_in[t, :, self.n_x] = c.z.b
                gp_post[t, :, 0] = c.z.p
                gp_post[t, :, 1] = c.z.q
            pr = GP_prior[g]
            eta0 = np.ascontiguousarray(np.asarray(pr[0], dtype=np.float64).reshape(M, -1))
            if eta0.shape[1] != 1:
                raise NotImplementedError("interface variables must be scalar (n_xi = 1)")
            eta1 = np.ascontiguousarray(np.asarray(pr[1], dtype=np.float64).reshape(M, M))
            sqrt_eig = np.ascontiguousarray(np.sqrt(hgp.eigen_val))                # src/BasisFunctions.py:79
            q.M, q.D, q.link = M, D, _LINK[link]
            q.sqrt_eig, q.gp_in, q.gp_post, q.eta0, q.eta1 = dptr(sqrt_eig), dptr(gp_in), dptr(gp_post), dptr(eta0), dptr(eta1)
            for d in range(D):
                q.center[d], q.half_width[d] = hgp.center[d], hgp.half_width[d]
            q.eta2 = float(np.asarray(pr[2], dtype=np.float64).reshape(-1)[0])
            q.eta3 = float(np.asarray(pr[3], dtype=np.float64))
            q.xi_mean = float(np.asarray(xi_mean[g], dtype=np.float64).reshape(-1)[0])
            q.xi_var = float(np.asarray(xi_cov[g], dtype=np.float64).reshape(-1)[0])
            keep += [sqrt_eig, gp_in, gp_post, eta0, eta1]
            self.hgp.append(hgp)
            self.M.append(M)
        p.observations = dptr(obs)
        p.out_link = _LINK[out_link]
        Q, R = SSM.process_noise, SSM.output_noise
        for i in range(self.n_x):
            p.m0[i] = m0[i]
            for j in range(self.n_x):
                p.Q[i][j] = Q[i, j]
                p.P0[i][j] = P0[i, j]
        for i in range(self.n_y):
            for j in range(self.n_y):
                p.R[i][j] = R[i, j]
        self._keep = keep
        self.prior_df = [float(np.asarray(GP_prior[g][3])) for g in range(self.G)]
        h = C.c_void_p()
        _lib.check(_lib.lib().pgas_marg_model_create(C.byref(p), C.byref(h)))
        self.handle = h
        self._ws = {}

    def workspace(self, N, n_chains, run=False):
        import torch
        k = (N, n_chains, run)
        if k not in self._ws:
            fn = _lib.lib().pgas_marg_run_workspace_bytes if run else _lib.lib().pgas_marg_workspace_bytes
            self._ws[k] = torch.empty(int(fn(self.handle, N, n_chains)), dtype=torch.uint8, device="cuda")
        return self._ws[k]

    def stat_shapes(self, lead):
        """shapes of the 4*G statistic arrays with leading dims `lead`"""
        out = []
        for M in self.M:
            out += [lead + (M,), lead + (M, M), lead, lead]
        return out

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().pgas_marg_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def make_marg_rng(key=None, chain_base=0, iteration=0, variates=None):
    """pgas_marg_rng for Philox mode (key) or injected mode (dict of CUDA tensors Z, ZXI0, U, TS)."""
    r = _lib.MargRng()
    if variates is not None:
        r.mode = 1
        for name in ("Z", "ZXI0", "U", "TS"):
            t = variates[name]
            assert t.is_cuda and t.is_contiguous() and t.dtype.is_floating_point
            setattr(r, name, t.data_ptr())
        r._keep = variates
    else:
        r.mode = 0
        r.seed = _random.as_key(key).seed
    r.chain_base, r.iteration = int(chain_base), int(iteration)
    return r


class Algorithm1:
    """Reference: src/Algorithm1.py:13-492."""

    def __init__(self, N_samples, observations, inputs, SSM, forgetting_factor, init_state_mean, init_state_cov,
                 init_int_var_mean, init_int_var_cov, GP_prior, basis_fcn, cluster_size=0):
        self.N_samples = int(N_samples)
        self.observations = np.asarray(observations, dtype=np.float64)
        self.inputs = np.asarray(inputs, dtype=np.float64)
        self.SSM = SSM
        self.forgetting_factor = float(forgetting_factor)
        self.init_state_mean = np.atleast_1d(np.asarray(init_state_mean, dtype=np.float64))
        self.init_state_cov = np.atleast_2d(np.asarray(init_state_cov, dtype=np.float64))
        self.init_int_var_mean = [np.atleast_1d(np.asarray(m, dtype=np.float64)) for m in init_int_var_mean]
        self.init_int_var_cov = [np.atleast_2d(np.asarray(c, dtype=np.float64)) for c in init_int_var_cov]
        self.basis_fcn = list(basis_fcn)
        self.GP_prior = [[np.asarray(GP_prior[i][j], dtype=np.float64) for j in range(4)] for i in range(len(GP_prior))]
        self.cluster_size = int(cluster_size)
        self._model = None
        self._dim_basis = None

    @property
    def model(self):
        if self._model is None:
            self._model = MargDeviceModel(self.observations, self.inputs, self.SSM, self.init_state_mean, self.init_state_cov,
                                          self.init_int_var_mean, self.init_int_var_cov, self.GP_prior, self.basis_fcn)
        return self._model

    @property
    def dim_basis(self):
        if self._dim_basis is None:                      # src/Algorithm1.py:57-63
            self._dim_basis = np.array([np.asarray(p[1]).shape[0] for p in self.GP_prior], dtype=np.int32)
        return self._dim_basis

    # ---- batched device API
    def _alloc_traces(self, nc):
        import torch
        m, N = self.model, self.N_samples
        f64 = dict(dtype=torch.float64, device="cuda")
        return dict(state_trace=torch.empty((nc, m.T, N, m.n_x), **f64), xi_trace=torch.empty((nc, m.G, m.T, N), **f64),
                    logw_trace=torch.empty((nc, m.T, N), **f64), anc_trace=torch.empty((nc, m.T - 1, N), dtype=torch.int32, device="cuda"),
                    status=torch.zeros((nc,), dtype=torch.int32, device="cuda"))

    def filter(self, key=None, variates=None, n_chains=1, chain_base=0, iteration=0, want_sst=True, want_final=True):
        """n_chains independent filters; returns a dict of CUDA tensors."""
        torch = _lib.require_cuda()
        m, N = self.model, self.N_samples
        out = self._alloc_traces(n_chains)
        f64 = dict(dtype=torch.float64, device="cuda")
        sst = [torch.empty(s, **f64) for s in m.stat_shapes((n_chains, m.T))] if want_sst else None
        fin = [torch.empty(s, **f64) for s in m.stat_shapes((n_chains, N))] if want_final else None
        ws = m.workspace(N, n_chains)
        rng = make_marg_rng(key, chain_base, iteration, variates)
        _lib.check(_lib.lib().pgas_marg_filter_f64(m.handle, N, n_chains, self.forgetting_factor, C.byref(rng), _lib.ptr(out["state_trace"]),
                                                   _lib.ptr(out["xi_trace"]), _lib.ptr(out["logw_trace"]), _lib.ptr(out["anc_trace"]),
                                                   _lib.ptr_array(sst), _lib.ptr_array(fin), _lib.ptr(out["status"]), self.cluster_size,
                                                   _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        out["sst_trace"], out["final_stats"] = sst, fin
        return out

    def outputs(self, states, xi):
        """vmap(vmap(output_mdl)) / vmap(vmap(log_likelihood)) (src/Algorithm1.py:463-480): states (T,n,n_x), xi (G,T,n)."""
        torch = _lib.require_cuda()
        m = self.model
        n = states.shape[1]
        obs = torch.empty((m.T, n, m.n_y), dtype=torch.float64, device="cuda")
        ll = torch.empty((m.T, n), dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().pgas_marg_outputs_f64(m.handle, _lib.ptr(states.contiguous()), _lib.ptr(xi.contiguous()), n, _lib.ptr(obs),
                                                    _lib.ptr(ll), _lib.stream_ptr()))
        return obs, ll

    def _squeeze_obs(self, obs):
        # output_mdl returning a scalar (x[0]) gives (T, n); a vector gives (T, n, n_y)
        if "output_model" in self.model.programs:                     # model plug-in: probe with the expression tracer
            from . import StateSpaceModel as _ssm
            s_sym, xi_sym, u_sym = _ssm.program_variables(self.model.n_x, self.model.G, self.inputs)
            probe = self.SSM.output_model(s_sym, u_sym, *xi_sym)
        else:
            probe = self.SSM.output_model(*_tr.variables(self.model.n_x, [1] * self.model.G)[:1], self.inputs[0],
                                          *_tr.variables(self.model.n_x, [1] * self.model.G)[1])
        return obs[..., 0] if getattr(probe, "scalar", False) else obs

    # ---- reference API
    def __call__(self, key):
        import torch
        m = self.model
        r = self.filter(key=key)
        st = r["state_trace"][0]
        obs, ll = self.outputs(st, r["xi_trace"][0])
        G = m.G
        int_var_trace = [r["xi_trace"][0, g].unsqueeze(-1).cpu().numpy() for g in range(G)]
        sst = [[r["sst_trace"][4 * g].cpu().numpy()[0][..., None], r["sst_trace"][4 * g + 1].cpu().numpy()[0],
                r["sst_trace"][4 * g + 2].cpu().numpy()[0][..., None, None], r["sst_trace"][4 * g + 3].cpu().numpy()[0]] for g in range(G)]
        fin = tuple((r["final_stats"][4 * g].cpu().numpy()[0][..., None], r["final_stats"][4 * g + 1].cpu().numpy()[0],
                     r["final_stats"][4 * g + 2].cpu().numpy()[0][..., None, None], r["final_stats"][4 * g + 3].cpu().numpy()[0])
                    for g in range(G))
        weights = torch.softmax(r["logw_trace"][0], dim=1).cpu().numpy()          # src/Algorithm1.py:460
        if int(r["status"][0]) != 0:
            raise _lib.PgasError("a per-particle eta1 lost positive definiteness (the reference would return NaN)")
        return (st.cpu().numpy(), int_var_trace, sst, weights, r["anc_trace"][0].cpu().numpy(), fin,
                self._squeeze_obs(obs).cpu().numpy(), ll.cpu().numpy())
