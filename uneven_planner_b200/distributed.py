"""Multi-GPU sharding of a batch of independent optimizeSE2Traj problems.

The problems never interact, so there is NO collective on the data path: every rank solves its own shard on its own
GPU with the map replicated, and one final all-gather of fixed-stride result records (NCCL over NVLink on the GPU box;
gloo in the CPU tests) leaves every rank with all results.
"""
import numpy as np
import torch
import torch.distributed as dist

RECORD_HEADER = 12  # [ret, outer, evals, iters, cost, jerk, T, res_h, res_g, N, M, pad]


def record_stride(n_max, m_max):
    return RECORD_HEADER + 12 * int(n_max) + 6 * int(m_max)


def shard_indices(cost, world):
    """Deal problems to ranks: sort by cost (samples per evaluation) descending and deal round-robin in a snake, so every
    rank gets the same mix of long and short problems.  Deterministic, identical on every rank.
    Returns a list of index arrays (ascending within a rank)."""
    cost = np.asarray(cost)
    order = np.argsort(-cost, kind="stable")
    shards = [[] for _ in range(world)]
    for pos, idx in enumerate(order):
        rnd, k = divmod(pos, world)
        r = k if rnd % 2 == 0 else world - 1 - k
        shards[r].append(int(idx))
    return [np.array(sorted(s), dtype=np.int64) for s in shards]


def all_gather_records(local, shards, rank, world, group=None):
    """local: [len(shards[rank]), stride] tensor on this rank's device.  Returns [B, stride] in global problem order
    on every rank.  One all_gather_into_tensor of a padded block per rank."""
    stride = local.shape[1]
    cnt = max(len(s) for s in shards)
    pad = torch.zeros((cnt, stride), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * cnt, stride), dtype=local.dtype, device=local.device)
    if world > 1:
        dist.all_gather_into_tensor(out, pad, group=group)
    else:
        out.copy_(pad)
    B = sum(len(s) for s in shards)
    full = torch.empty((B, stride), dtype=local.dtype, device=local.device)
    for r, s in enumerate(shards):
        if len(s):
            full[torch.as_tensor(s, device=local.device)] = out[r * cnt: r * cnt + len(s)]
    return full
