"""Multi-GPU plumbing (one process per GPU, torch.distributed): environments are sharded across ranks with
no data-path collective inside the rollout; the update needs
  * once per iteration: global advantage moments (n, mean, M2), global exps count, logger / ZFilter sums
  * once per PPO epoch and net: one sum all-reduce of the flat gradient buffer
These helpers work on CPU (gloo) and CUDA (nccl) tensors alike."""
import torch


def group():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def world_size():
    d = group()
    return d.get_world_size() if d is not None else 1


def allreduce_sum_(t):
    d = group()
    if d is not None:
        d.all_reduce(t)
    return t


def merge_moments_(stats):
    """stats = (n, mean, M2) of the local advantages -> the same for the union over ranks (Chan et al.),
    computed identically on every rank so replicas stay bit-identical."""
    d = group()
    if d is None:
        return stats
    gathered = [torch.empty_like(stats) for _ in range(d.get_world_size())]
    d.all_gather(gathered, stats.contiguous())
    # pooled form of the Chan merge, evaluated on the device (no host round trip) from the same gathered rows in
    # the same order on every rank: n = sum n_i, mean = sum n_i mean_i / n, M2 = sum (M2_i + n_i (mean_i - mean)^2)
    g = torch.stack(gathered).to(torch.float64)
    nb, mb, sb = g[:, 0], g[:, 1], g[:, 2]
    n = nb.sum()
    mean = (nb * mb).sum() / torch.clamp(n, min=1.0)
    m2 = (sb + nb * (mb - mean) ** 2).sum()
    stats.copy_(torch.stack([n, mean, m2]).to(stats.dtype))
    return stats


def reduce_logger_(lg, min_slots, max_slots):
    """sum every slot of the rollout logger vector except the min / max slots"""
    d = group()
    if d is None:
        return lg
    mins = lg[list(min_slots)].clone()
    maxs = lg[list(max_slots)].clone()
    d.all_reduce(lg)
    d.all_reduce(mins, op=d.ReduceOp.MIN)
    d.all_reduce(maxs, op=d.ReduceOp.MAX)
    lg[list(min_slots)] = mins
    lg[list(max_slots)] = maxs
    return lg


def shard_envs(n_env, rank=None, world=None):
    """contiguous shard of the global environment index range for this rank -> (first, count)"""
    d = group()
    if world is None:
        world = d.get_world_size() if d is not None else 1
    if rank is None:
        rank = d.get_rank() if d is not None else 0
    base, rem = divmod(n_env, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count
