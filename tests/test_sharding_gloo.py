"""CPU-only, world_size 2 over gloo: the ensemble-sharding rules.
 * each rank's Philox noise for members [first, first+B_local) equals the unsharded stream's slice;
 * the PC sampler's batch-mean gradient norm from all-gathered per-member sums equals the unsharded value."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import philox_ref


def shard_range(rank: int, world: int, members: int):
    per = members // world
    return rank * per, per


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    members, per_member, seed = 8, 64, 31
    first, local = shard_range(rank, world, members)
    mine = philox_ref.normal(local * per_member, seed, 3, first * per_member)
    full = philox_ref.normal(members * per_member, seed, 3)
    ok_noise = np.array_equal(mine, full[first * per_member:(first + local) * per_member])
    score = torch.from_numpy(full.reshape(members, per_member))
    local_sumsq = (score[first:first + local] ** 2).sum(1)
    gathered = torch.empty(members)
    dist.all_gather_into_tensor(gathered, local_sumsq)
    gn_sharded = gathered.sqrt().mean()
    gn_full = torch.norm(score, dim=-1).mean()
    ret[rank] = bool(ok_noise) and bool(torch.allclose(gn_sharded, gn_full, rtol=1e-6))
    dist.destroy_process_group()


def test_two_rank_sharding_rules():
    world = 2
    from conftest import free_port
    port = free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
