"""Data-parallel plumbing: env shards + Q-replica averaging.

Envs are independent, so a batch shards by contiguous env-id ranges with no data-path
collective; the RNG is keyed by GLOBAL env id, so a shard computes exactly what the same envs
would compute inside one big batch.  The only exchange step is the periodic averaging of the
per-GPU Q replicas: one all-reduce over NCCL/NVLink (AVG), or SUM + scale where the backend has
no AVG (gloo, used by the CPU tests).  The reference has no counterpart (single process).
"""
import torch
import torch.distributed as dist


def shard_env_ids(envs_per_rank, rank, world):
    """(env_id_base, n_envs) of `rank`: weak scaling, contiguous global ids."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return rank * envs_per_rank, envs_per_rank


def split_env_ids(total_envs, rank, world):
    """Strong-scaling split of `total_envs` ids: first (total % world) ranks get one extra."""
    base, extra = divmod(total_envs, world)
    n = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, n


def average_(tensor, group=None):
    """In-place mean over the ranks of `group`; a no-op without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()):
        return tensor
    world = dist.get_world_size(group)
    if world == 1:
        return tensor
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(tensor, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
        tensor.div_(world)
    return tensor


class ReplicaSync:
    """Calls average_ on the Q replica every `interval` rounds (0 / None = never)."""

    def __init__(self, interval, group=None):
        self.interval = int(interval or 0)
        self.group = group
        self.syncs = 0

    def due(self, rounds_done):
        return self.interval > 0 and rounds_done > 0 and rounds_done % self.interval == 0

    def step(self, q, rounds_done):
        if self.due(rounds_done):
            average_(q, self.group)
            self.syncs += 1
            return True
        return False
