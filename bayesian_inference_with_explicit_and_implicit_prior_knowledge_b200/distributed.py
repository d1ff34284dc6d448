"""Multi-GPU layer: independent chains / Monte-Carlo replicates are partitioned across ranks (one
process per GPU, torch.distributed).  A sweep is sequential in time, so nothing is exchanged while
the chains run; the only collective is one final gather of the per-chain traces.
The Philox counter carries the GLOBAL chain id, so a chain's result does not depend on how many
ranks there are.
"""
import numpy as np


def shard_chains(n_chains, rank, world_size):
    """contiguous block of global chain ids owned by `rank`: (first, count)"""
    base, rem = divmod(int(n_chains), int(world_size))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def gather_chain_outputs(local, n_chains, group=None):
    """all-gather a per-chain tensor (first axis = this rank's chains) into (n_chains, ...) on every rank.
    Ranks may hold different chain counts: pad to the maximum, gather, trim."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    counts = [shard_chains(n_chains, r, world)[1] for r in range(world)]
    cmax = max(counts)
    pad = torch.zeros((cmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def run_chains_distributed(pgas, key, init_ref_state, n_chains, group=None, want_params=False):
    """PGAS.run_chains for `n_chains` global chains sharded over the ranks of `group`; returns the
    gathered state trace (n_chains, K, T, n_x) (and parameter traces) on every rank."""
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    first, count = shard_chains(n_chains, rank, world)
    ref = np.asarray(init_ref_state, dtype=np.float64)
    if ref.ndim == 3:
        if ref.shape[0] != n_chains:
            raise ValueError(f"init_ref_state holds {ref.shape[0]} trajectories for {n_chains} chains")
        ref = ref[first:first + count]
    if count > 0:
        out = pgas.run_chains(key, ref, n_chains=count, chain_base=first, want_params=want_params)
    else:
        # more ranks than chains: this rank owns none.  It still takes part in the gathers, with zero-length blocks of the
        # right trailing shape (running a dummy chain would reuse another rank's chain id).
        import torch
        m, K = pgas.cSMC.model, pgas.N_iterations
        z = lambda *shape: torch.zeros((0,) + shape, dtype=torch.float64, device="cuda")   # noqa: E731
        out = dict(state_trace=z(K, m.T, m.n_x), A_trace=z(K, m.n_x, m.M), S_trace=z(K, m.n_x, m.n_x))
    res = dict(state_trace=gather_chain_outputs(out["state_trace"], n_chains, group))
    if want_params:
        res["A_trace"] = gather_chain_outputs(out["A_trace"], n_chains, group)
        res["S_trace"] = gather_chain_outputs(out["S_trace"], n_chains, group)
    return res


def run_replicas_distributed(alg2, key, init_ref_state, init_ref_int_var, n_replicas, group=None, K=None):
    """Monte-Carlo replicates of the marginalised PGAS (Algorithm2): the shipped examples run ONE chain, so the
    multi-GPU form is `n_replicas` independent chains with global chain ids 0..n_replicas-1 sharded over the ranks
    ("replicas only"), pooled at the end.  init_ref_state (T,n_x) and init_ref_int_var [G x (T,)] are
    shared by all replicas.  Returns dict(x_trace (n_replicas,K,T,n_x), xi_trace (n_replicas,G,K,T)) on every rank."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    first, count = shard_chains(n_replicas, rank, world)
    m = alg2.cSMC.model
    f64 = dict(dtype=torch.float64, device="cuda")
    x0 = torch.as_tensor(np.asarray(init_ref_state, dtype=np.float64).reshape(1, m.T, m.n_x), **f64).repeat(max(count, 1), 1, 1)
    xi0 = torch.as_tensor(np.stack([np.asarray(v, dtype=np.float64).reshape(m.T) for v in init_ref_int_var])[None], **f64).repeat(max(count, 1), 1, 1)
    out = alg2.run(x0, xi0, key=key, K=K, chain_base=first, want_sst=False)
    return dict(x_trace=gather_chain_outputs(out["x_trace"][:count], n_replicas, group),
                xi_trace=gather_chain_outputs(out["xi_trace"][:count], n_replicas, group))
