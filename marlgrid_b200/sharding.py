"""Env-index sharding of one batched env family across ranks (one process per GPU).

Env instances never interact (the reference's MultiGridEnv objects share only read-mostly caches,
marlgrid/base.py:85), so the batch partitions into contiguous global index ranges with NO collective on
the data path.  The Philox stream is keyed by the GLOBAL env index (DESIGN.md "RNG contract"), which makes
results independent of how the batch is cut.  The only cross-rank traffic is optional scalar statistics.
"""


def shard_range(total_envs, rank, world_size):
    """Contiguous [offset, offset+count) slice of `total_envs` owned by `rank` (remainder to the low ranks)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def make_sharded(env_factory, total_envs, rank, world_size, **kwargs):
    """Build this rank's shard: env_factory(num_envs=count, env_offset=offset, **kwargs)."""
    offset, count = shard_range(total_envs, rank, world_size)
    return env_factory(num_envs=count, env_offset=offset, **kwargs)


def global_stats(local_sum, local_count, dist=None):
    """Mean of a per-env statistic over all shards: one scalar all-reduce, off the hot path."""
    import torch

    t = torch.tensor([float(local_sum), float(local_count)], dtype=torch.float64)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t)
    return float(t[0] / max(t[1], 1.0))
