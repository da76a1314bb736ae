"""N>1 path on CPU: world_size-2 gloo run of the shard + final all-gather logic (no solver involved)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from uneven_planner_b200 import distributed as D


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, B, stride, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cost = (np.arange(B) * 7919) % 23 + 5
    shards = D.shard_indices(cost, world)
    mine = shards[rank]
    local = torch.zeros((len(mine), stride), dtype=torch.float64)
    for k, g in enumerate(mine):
        local[k] = torch.arange(stride, dtype=torch.float64) + 1000.0 * g  # record content identifies its global index
    full = D.all_gather_records(local, shards, rank, world)
    ok = all(bool(torch.equal(full[g], torch.arange(stride, dtype=torch.float64) + 1000.0 * g)) for g in range(B))
    q.put((rank, ok, [len(s) for s in shards]))
    dist.destroy_process_group()


def test_shards_cover_batch_once_and_balance():
    cost = np.array([5, 50, 7, 9, 31, 2, 44, 8, 8, 8, 21])
    for world in (1, 2, 4, 8):
        sh = D.shard_indices(cost, world)
        allidx = np.sort(np.concatenate(sh))
        assert np.array_equal(allidx, np.arange(len(cost)))
        loads = [cost[s].sum() for s in sh if len(s)]
        assert max(len(s) for s in sh) - min(len(s) for s in sh) <= 1
        if world == 2:
            assert max(loads) - min(loads) <= cost.max()


def test_two_rank_gloo_all_gather_restores_global_order():
    world, B, stride = 2, 13, D.record_stride(4, 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, stride, q)) for r in range(world)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs: p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    assert sorted(res[0][2]) == [6, 7]
