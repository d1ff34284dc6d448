"""Marginalised PGAS outer loop (reference src/Algorithm2.py), host side.

`Algorithm2(...)` keeps the reference constructor keywords (src/Algorithm2.py:12-25) and
`__call__(key, init_ref_state, init_ref_int_var)` -> the reference's 6-tuple (:106-187).  The K
iterations (reference statistics -> conditional sweep -> pick -> backward trace) are enqueued
stream-ordered by one C call, pgas_marg_run_f64 (csrc/marginal.cu), with no host synchronisation
inside the loop — the reference crosses the Python/XLA boundary K*(T-1) times.
"""
import ctypes as C

import numpy as np

from . import _lib
from .Algorithm1 import make_marg_rng
from .Algorithm3 import Algorithm3


class Algorithm2:
    def __init__(self, N_samples, N_iterations, observations, inputs, SSM, init_state_mean, init_state_cov,
                 init_int_var_mean, init_int_var_cov, GP_prior, basis_fcn, cluster_size=0):
        self.N_iterations = int(N_iterations)
        self.N_steps = np.asarray(observations).shape[0]
        self.cSMC = Algorithm3(N_samples=N_samples, observations=observations, inputs=inputs, SSM=SSM,
                               init_state_mean=init_state_mean, init_state_cov=init_state_cov,
                               init_int_var_mean=init_int_var_mean, init_int_var_cov=init_int_var_cov,
                               GP_prior=GP_prior, basis_fcn=basis_fcn, cluster_size=cluster_size)

    def run(self, init_x, init_xi, key=None, variates=None, K=None, chain_base=0, want_sst=True):
        """n_chains independent Gibbs chains.  init_x (n_chains,T,n_x), init_xi (n_chains,G,T) CUDA tensors ->
        dict(x_trace (n_chains,K,T,n_x), xi_trace (n_chains,G,K,T), sst [4*G tensors (n_chains,K,...)], status)."""
        torch = _lib.require_cuda()
        c = self.cSMC
        m, N = c.model, c.N_samples
        K = self.N_iterations if K is None else int(K)
        init_x = init_x.reshape(-1, m.T, m.n_x).contiguous()
        nc = init_x.shape[0]
        init_xi = init_xi.reshape(nc, m.G, m.T).contiguous()
        f64 = dict(dtype=torch.float64, device="cuda")
        xtr = torch.empty((nc, K, m.T, m.n_x), **f64)
        xitr = torch.empty((nc, m.G, K, m.T), **f64)
        sst = [torch.empty(s, **f64) for s in m.stat_shapes((nc, K))] if want_sst else None
        status = torch.zeros((nc,), dtype=torch.int32, device="cuda")
        ws = m.workspace(N, nc, run=True)
        rng = make_marg_rng(key, chain_base, 0, variates)
        _lib.check(_lib.lib().pgas_marg_run_f64(m.handle, N, K, nc, _lib.ptr(init_x), _lib.ptr(init_xi), C.byref(rng), _lib.ptr(xtr),
                                                _lib.ptr(xitr), _lib.ptr_array(sst), _lib.ptr(status), c.cluster_size, _lib.ptr(ws),
                                                ws.numel(), _lib.stream_ptr()))
        return dict(x_trace=xtr, xi_trace=xitr, sst=sst, status=status)

    def __call__(self, key, init_ref_state, init_ref_int_var):
        torch = _lib.require_cuda()
        c = self.cSMC
        m = c.model
        T, K, G = m.T, self.N_iterations, m.G
        f64 = dict(dtype=torch.float64, device="cuda")
        x0 = torch.as_tensor(np.asarray(init_ref_state, dtype=np.float64).reshape(1, T, m.n_x), **f64)
        xi0 = torch.as_tensor(np.stack([np.asarray(v, dtype=np.float64).reshape(T) for v in init_ref_int_var])[None], **f64)
        r = self.run(x0, xi0, key=key)
        if int(r["status"][0]) != 0:
            raise _lib.PgasError("a per-particle eta1 lost positive definiteness (the reference would return NaN)")
        states = r["x_trace"][0].transpose(0, 1).contiguous()                      # (T, K, n_x)  src/Algorithm2.py:153
        xis = r["xi_trace"][0].transpose(1, 2).contiguous()                        # (G, T, K)
        obs, ll = c.outputs(states, xis)
        state_trace = states.cpu().numpy()
        int_var_trace = [xis[g].unsqueeze(-1).cpu().numpy() for g in range(G)]
        sst = [[r["sst"][4 * g].cpu().numpy()[0][..., None], r["sst"][4 * g + 1].cpu().numpy()[0],
                r["sst"][4 * g + 2].cpu().numpy()[0][..., None, None], r["sst"][4 * g + 3].cpu().numpy()[0]] for g in range(G)]
        return (state_trace, int_var_trace, np.ones((self.N_steps, K)) / K, sst, c._squeeze_obs(obs).cpu().numpy(), ll.cpu().numpy())
