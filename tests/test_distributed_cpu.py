"""world_size-2 gloo tests of the multi-GPU layer's host logic (chain sharding + final gather)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_chains, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    DI = helpers.pkg("distributed")
    first, count = DI.shard_chains(n_chains, rank, world)
    # stand-in for the per-chain device result: value = global chain id (what the Philox counter keys on)
    local = torch.arange(first, first + count, dtype=torch.float64)[:, None, None] * torch.ones((count, 3, 2), dtype=torch.float64)
    full = DI.gather_chain_outputs(local, n_chains)
    q.put((rank, first, count, full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_chains", [8, 5, 2, 1])      # 1: the second rank owns no chain (zero-length block in the gather)
def test_shard_and_gather_world2(n_chains):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_chains, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.arange(n_chains, dtype=np.float64)[:, None, None] * np.ones((n_chains, 3, 2))
    covered = []
    for rank, first, count, full in res:
        assert full.shape == (n_chains, 3, 2)
        assert np.array_equal(full, expect)                  # every rank holds all chains, in global chain order
        covered += list(range(first, first + count))
    assert sorted(covered) == list(range(n_chains))          # a partition: no chain lost or duplicated


def test_shard_chains_is_a_balanced_partition():
    DI = helpers.pkg("distributed")
    for n in (1, 7, 64, 65):
        for world in (1, 2, 4, 8):
            blocks = [DI.shard_chains(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            assert all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
