"""Theta-conditioned conditional SMC with ancestor sampling and the PGAS outer loop
(reference src/PGAS.py), host side.  Class names, constructor keywords, call signatures and
return shapes follow the reference; the work is done by libpgas_b200.so (sweep.cu, suffstats.cu,
mniw_draw.cu, chains.cu).  Keys are `random.PhiloxKey`s (see random.py).

Beyond the reference API, `sweep`, `step` and `run_chains` expose the batched / injected-variates
forms the parity tests, the benchmark and the multi-GPU layer use.
"""
import ctypes as C

import numpy as np

from . import _lib
from . import models as _models
from . import random as _random
from . import BayesianInferrence as BI


def _make_rng(key=None, chain_base=0, iteration=0, variates=None):
    """pgas_rng for Philox mode (key) or injected mode (dict of CUDA tensors Z,U / chi2,G,Nrm)."""
    r = _lib.Rng()
    if variates is not None:
        r.mode = 1
        for name in ("Z", "U", "chi2", "G", "Nrm"):
            t = variates.get(name)
            setattr(r, name, t.data_ptr() if t is not None else None)
        r._keep = variates                      # keep the tensors alive
    else:
        r.mode = 0
        r.seed = _random.as_key(key).seed
    r.chain_base = int(chain_base)
    r.iteration = int(iteration)
    return r


class condSequentialMonteCarlo:
    """Reference: src/PGAS.py:14-228."""

    def __init__(self, N_samples, observations, inputs, init_state_mean, init_state_cov, likelihood_fcn, basis_fcn,
                 flags=0, cluster_size=0):
        self.N_samples = int(N_samples)
        self.observations = np.asarray(observations, dtype=np.float64)
        self.inputs = np.asarray(inputs, dtype=np.float64)
        self.init_state_mean = np.atleast_1d(np.asarray(init_state_mean, dtype=np.float64))
        self.init_state_cov = np.atleast_2d(np.asarray(init_state_cov, dtype=np.float64))
        self.likelihood_fcn = likelihood_fcn
        self.basis_fcn = basis_fcn
        self.flags = int(flags)
        self.cluster_size = int(cluster_size)
        self._model = None
        n_u = self.inputs.reshape(self.observations.shape[0], -1).shape[1] if self.inputs.size else 0
        self.dim_basis = len(_models.trace_basis(basis_fcn, self.init_state_mean.shape[0], n_u))   # src/PGAS.py:41-43

    @property
    def model(self):
        if self._model is None:
            self._model = _models.DeviceModel(self.observations, self.inputs, self.init_state_mean, self.init_state_cov,
                                              self.likelihood_fcn, self.basis_fcn, self.flags)
        return self._model

    # ---- batched device API -------------------------------------------------------------
    def sweep(self, ref, Theta, Sigma, key=None, variates=None, chain_base=0, iteration=0, want_traces=True):
        """n_chains sweeps.  ref (n_chains,T,n_x), Theta (n_chains,n_x,M), Sigma (n_chains,n_x,n_x) CUDA
        float64 tensors.  Returns dict(traj, state_trace, anc_trace, logw_last, idx) of CUDA tensors.
        anc_trace[..., N-1] is the reference particle's ancestor as src/PGAS.py:122-127 stores it (unclipped: it is N when
        rounding leaves cumsum(w)[-1] below u_anc); clamp to N-1 before indexing with it (Filtering.reconstruct_trajectory does)."""
        torch = _lib.require_cuda()
        m = self.model
        ref = ref.reshape(-1, m.T, m.n_x).contiguous()
        nc = ref.shape[0]
        Theta = Theta.reshape(nc, m.n_x, m.M).contiguous()
        Sigma = Sigma.reshape(nc, m.n_x, m.n_x).contiguous()
        N = self.N_samples
        st = torch.empty((nc, m.T, N, m.n_x), dtype=torch.float64, device="cuda")
        anc = torch.empty((nc, m.T - 1, N), dtype=torch.int32, device="cuda")
        lw = torch.empty((nc, N), dtype=torch.float64, device="cuda")
        idx = torch.empty((nc,), dtype=torch.int32, device="cuda")
        traj = torch.empty((nc, m.T, m.n_x), dtype=torch.float64, device="cuda")
        rng = _make_rng(key, chain_base, iteration, variates)
        nbytes = int(_lib.lib().pgas_csmc_sweep_workspace_bytes(m.handle, N, nc))
        if getattr(self, "_sweep_ws", None) is None or self._sweep_ws.numel() < nbytes:
            self._sweep_ws = torch.empty((nbytes,), dtype=torch.uint8, device="cuda")
        _lib.check(_lib.lib().pgas_csmc_sweep_f64(m.handle, N, nc, _lib.ptr(ref), _lib.ptr(Theta), _lib.ptr(Sigma),
                                                  C.byref(rng), _lib.ptr(st), _lib.ptr(anc), _lib.ptr(lw), _lib.ptr(idx),
                                                  _lib.ptr(traj), self.cluster_size, _lib.ptr(self._sweep_ws), nbytes, _lib.stream_ptr()))
        return dict(traj=traj, state_trace=st, anc_trace=anc, logw_last=lw, idx=idx)

    def step(self, time, log_weights, state, coeff_mat, error_cov, ref_state, u2, z):
        """One `condSequentialMonteCarlo.step` (src/PGAS.py:79-153) under injected variates u2=(u_res,u_anc),
        z (N,n_x).  CUDA tensors in, (new_log_weights, new_state, a_indices) CUDA tensors out."""
        torch = _lib.require_cuda()
        m = self.model
        N = self.N_samples
        f = lambda t: torch.as_tensor(t, dtype=torch.float64).contiguous().cuda()
        lw, x, Th, Sg, rf, uu, zz = map(f, (log_weights, state, coeff_mat, error_cov, ref_state, u2, z))
        lw_o = torch.empty((N,), dtype=torch.float64, device="cuda")
        x_o = torch.empty((N, m.n_x), dtype=torch.float64, device="cuda")
        a_o = torch.empty((N,), dtype=torch.int32, device="cuda")
        _lib.check(_lib.lib().pgas_csmc_step_f64(m.handle, N, int(time), _lib.ptr(lw), _lib.ptr(x), _lib.ptr(Th), _lib.ptr(Sg),
                                                 _lib.ptr(rf), _lib.ptr(uu), _lib.ptr(zz), _lib.ptr(lw_o), _lib.ptr(x_o),
                                                 _lib.ptr(a_o), self.cluster_size, _lib.stream_ptr()))
        return lw_o, x_o, a_o

    # ---- reference call ------------------------------------------------------------------
    def __call__(self, key, ref_state, coeff_mat, error_cov):
        """Reference signature (src/PGAS.py:176-182): returns the sampled trajectory (T,n_x) (or (T,))."""
        torch = _lib.require_cuda()
        f = lambda t: torch.as_tensor(np.asarray(t, dtype=np.float64)).cuda()
        out = self.sweep(f(ref_state), f(np.atleast_2d(coeff_mat)), f(np.atleast_2d(error_cov)), key=key)
        return np.squeeze(out["traj"][0].cpu().numpy())


class PGAS:
    """Reference: src/PGAS.py:231-397."""

    def __init__(self, N_samples, N_iterations, observations, inputs, init_state_mean, init_state_cov, likelihood_fcn,
                 GP_prior, basis_fcn, flags=0, cluster_size=0):
        self.N_iterations = int(N_iterations)
        self.N_steps = np.asarray(observations).shape[0]
        self.GP_prior = tuple(np.asarray(g, dtype=np.float64) for g in GP_prior[:3]) + (float(GP_prior[3]),)
        self.cSMC = condSequentialMonteCarlo(N_samples=N_samples, observations=observations, inputs=inputs,
                                             init_state_mean=init_state_mean, init_state_cov=init_state_cov,
                                             likelihood_fcn=likelihood_fcn, basis_fcn=basis_fcn, flags=flags,
                                             cluster_size=cluster_size)
        self._prior_dev = None

    def _prior(self):
        if self._prior_dev is None:
            torch = _lib.require_cuda()
            self._prior_dev = tuple(torch.as_tensor(np.ascontiguousarray(g)).cuda() for g in self.GP_prior[:3])
        return self._prior_dev

    def sample_params(self, key, state_trajectory, variates=None, chain_base=0, iteration=0):
        """Reference signature `sample_params(key, state_trajectory)` (src/PGAS.py:288-343) -> (A, S).
        Accepts a (T,n_x) array (returns numpy) or a (n_chains,T,n_x) CUDA tensor (returns CUDA tensors)."""
        torch = _lib.require_cuda()
        m = self.cSMC.model
        batched = torch.is_tensor(state_trajectory) and state_trajectory.ndim == 3
        traj = state_trajectory if batched else torch.as_tensor(
            np.asarray(state_trajectory, dtype=np.float64).reshape(1, m.T, m.n_x)).cuda()
        T0, T1, T2, T3 = BI.trajectory_statistics(m, traj)
        p0, p1, p2 = self._prior()
        rng = _make_rng(key, chain_base, iteration, variates)
        A, S, status = BI.mniw_posterior_draw(p0 + T0, p1 + T1, p2 + T2, self.GP_prior[3] + T3, rng, self.cSMC.flags)
        if batched:
            return A, S
        return A[0].cpu().numpy(), S[0].cpu().numpy()

    def run_chains(self, key, init_ref_state, n_chains=1, chain_base=0, variates=None, want_params=True, iteration=0, K=None):
        """K iterations for n_chains independent chains, entirely stream-ordered on the device
        (pgas_run_chains_f64).  init_ref_state (T,n_x) (shared) or (n_chains,T,n_x).
        Returns dict(state_trace (n_chains,K,T,n_x), A_trace, S_trace) of CUDA tensors.

        Checkpoint / resume: the whole Gibbs state is (reference trajectory, key, iteration index) — the Philox
        counters carry (chain, iteration, time, particle), nothing else persists.  A run of K iterations equals a
        run of K1 followed by `run_chains(key, state_trace[:, K1-1], iteration=K1-1, K=K-K1+1)` bit for bit."""
        torch = _lib.require_cuda()
        m = self.cSMC.model
        K, N = (self.N_iterations if K is None else int(K)), self.cSMC.N_samples
        ref = torch.as_tensor(init_ref_state, dtype=torch.float64)
        ref = ref.reshape(-1, m.T, m.n_x)
        if ref.shape[0] == 1 and n_chains > 1:
            ref = ref.expand(n_chains, m.T, m.n_x)
        ref = ref.contiguous().cuda()
        n_chains = ref.shape[0]
        out = torch.empty((n_chains, K, m.T, m.n_x), dtype=torch.float64, device="cuda")
        A_tr = torch.empty((n_chains, K, m.n_x, m.M), dtype=torch.float64, device="cuda") if want_params else None
        S_tr = torch.empty((n_chains, K, m.n_x, m.n_x), dtype=torch.float64, device="cuda") if want_params else None
        nbytes = _lib.lib().pgas_run_chains_workspace_bytes(m.handle, N, n_chains)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device="cuda")
        p0, p1, p2 = self._prior()
        rng = _make_rng(key, chain_base, iteration, variates)
        _lib.check(_lib.lib().pgas_run_chains_f64(m.handle, N, K, n_chains, _lib.ptr(p0), _lib.ptr(p1), _lib.ptr(p2),
                                                  self.GP_prior[3], _lib.ptr(ref), C.byref(rng), _lib.ptr(out), _lib.ptr(A_tr),
                                                  _lib.ptr(S_tr), self.cSMC.cluster_size, _lib.ptr(ws), nbytes, _lib.stream_ptr()))
        return dict(state_trace=out, A_trace=A_tr, S_trace=S_tr, workspace=ws)

    def __call__(self, key, init_ref_state):
        """Reference signature (src/PGAS.py:345-349): returns (state_trace (T,K,n_x), log_likelihood (T,K))."""
        torch = _lib.require_cuda()
        m = self.cSMC.model
        res = self.run_chains(key, np.atleast_2d(np.asarray(init_ref_state, dtype=np.float64).T).T, n_chains=1,
                              want_params=False)
        st = res["state_trace"][0].permute(1, 0, 2).contiguous()            # (T,K,n_x)  (src/PGAS.py:380)
        obs = torch.as_tensor(m._obs).cuda()                                # (T,n_y)
        if isinstance(m.likelihood, _models.ProgramLikelihood):             # model plug-in: the program on whole trajectories
            ll = m.likelihood.logpdf_torch(obs[:, None, :], st, torch.as_tensor(m._inp).cuda()[:, None, :])
        else:
            ll = m.likelihood.logpdf_torch(obs[:, None, :], st)             # (T,K)      (src/PGAS.py:383-392)
        return st.cpu().numpy(), ll.cpu().numpy()
