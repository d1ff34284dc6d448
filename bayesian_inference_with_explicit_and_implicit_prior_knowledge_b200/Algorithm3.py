"""Marginalised conditional SMC with ancestor sampling (reference src/Algorithm3.py), host side.

`Algorithm3(Algorithm1)` keeps the reference constructor (no forgetting factor: fixed 1.0,
src/Algorithm3.py:34) and `__call__(key, ref_state, ref_int_var, ref_suff_stats)` ->
`(state_traj, int_var_traj)` (:199-303).  All T steps, the final pick and the backward trace run on
the device (csrc/marginal.cu: marg_refstats_kernel, marg_sweep_kernel<1>, marg_pick_trace_kernel).
"""
import ctypes as C

import numpy as np

from . import _lib
from .Algorithm1 import Algorithm1, make_marg_rng


class Algorithm3(Algorithm1):
    def __init__(self, N_samples, observations, inputs, SSM, init_state_mean, init_state_cov, init_int_var_mean,
                 init_int_var_cov, GP_prior, basis_fcn, cluster_size=0):
        super().__init__(N_samples, observations, inputs, SSM, 1.0, init_state_mean, init_state_cov, init_int_var_mean,
                         init_int_var_cov, GP_prior, basis_fcn, cluster_size)

    def reference_stats(self, x_traj, xi_traj):
        """Sum over all T steps of calcStatistics(xi_t, basis(x_t, u_t)) (src/Algorithm2.py:83-96, :139-152).
        x_traj (n_chains,T,n_x), xi_traj (n_chains,G,T) CUDA tensors -> list of 4*G CUDA tensors."""
        torch = _lib.require_cuda()
        m = self.model
        nc = x_traj.shape[0]
        out = [torch.empty(s, dtype=torch.float64, device="cuda") for s in m.stat_shapes((nc,))]
        _lib.check(_lib.lib().pgas_marg_refstats_f64(m.handle, _lib.ptr(x_traj.contiguous()), m.T * m.n_x, _lib.ptr(xi_traj.contiguous()),
                                                     m.G * m.T, m.T, nc, _lib.ptr_array(out), _lib.stream_ptr()))
        return out

    def csmc(self, ref_x, ref_xi, ref_stats=None, key=None, variates=None, chain_base=0, iteration=0):
        """n_chains conditional sweeps.  ref_x (n_chains,T,n_x), ref_xi (n_chains,G,T), ref_stats list of 4*G tensors
        (n_chains,...) or None (= statistics of the reference trajectory itself).  Returns a dict of CUDA tensors."""
        torch = _lib.require_cuda()
        m, N = self.model, self.N_samples
        ref_x = ref_x.reshape(-1, m.T, m.n_x).contiguous()
        nc = ref_x.shape[0]
        ref_xi = ref_xi.reshape(nc, m.G, m.T).contiguous()
        out = self._alloc_traces(nc)
        out["idx"] = torch.empty((nc,), dtype=torch.int32, device="cuda")
        out["traj"] = torch.empty((nc, m.T, m.n_x), dtype=torch.float64, device="cuda")
        out["xi_traj"] = torch.empty((nc, m.G, m.T), dtype=torch.float64, device="cuda")
        ws = m.workspace(N, nc)
        rng = make_marg_rng(key, chain_base, iteration, variates)
        if ref_stats is not None:
            ref_stats = [t.contiguous() for t in ref_stats]
        _lib.check(_lib.lib().pgas_marg_csmc_f64(m.handle, N, nc, _lib.ptr(ref_x), _lib.ptr(ref_xi), _lib.ptr_array(ref_stats), C.byref(rng),
                                                 _lib.ptr(out["state_trace"]), _lib.ptr(out["xi_trace"]), _lib.ptr(out["logw_trace"]),
                                                 _lib.ptr(out["anc_trace"]), _lib.ptr(out["idx"]), _lib.ptr(out["traj"]),
                                                 _lib.ptr(out["xi_traj"]), _lib.ptr(out["status"]), self.cluster_size, _lib.ptr(ws),
                                                 ws.numel(), _lib.stream_ptr()))
        return out

    def __call__(self, key, ref_state, ref_int_var, ref_suff_stats):
        torch = _lib.require_cuda()
        m = self.model
        T = m.T
        f64 = dict(dtype=torch.float64, device="cuda")
        rx = torch.as_tensor(np.asarray(ref_state, dtype=np.float64).reshape(1, T, m.n_x), **f64)
        rxi = torch.as_tensor(np.stack([np.asarray(v, dtype=np.float64).reshape(T) for v in ref_int_var])[None], **f64)
        stats = []
        for g in range(m.G):
            M = m.M[g]
            s = ref_suff_stats[g]
            stats += [torch.as_tensor(np.asarray(s[0], dtype=np.float64).reshape(1, M), **f64),
                      torch.as_tensor(np.asarray(s[1], dtype=np.float64).reshape(1, M, M), **f64),
                      torch.as_tensor(np.asarray(s[2], dtype=np.float64).reshape(1), **f64),
                      torch.as_tensor(np.asarray(s[3], dtype=np.float64).reshape(1), **f64)]
        r = self.csmc(rx, rxi, stats, key=key)
        if int(r["status"][0]) != 0:
            raise _lib.PgasError("a per-particle eta1 lost positive definiteness (the reference would return NaN)")
        state_traj = np.squeeze(r["traj"][0].cpu().numpy())                       # reconstruct_trajectory squeezes
        int_var_traj = tuple(r["xi_traj"][0, g].cpu().numpy() for g in range(m.G))
        return state_traj, int_var_traj
