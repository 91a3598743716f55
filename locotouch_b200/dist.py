"""Env-sharded data parallelism helpers: one process per GPU, ``torch.distributed`` for the plumbing.

The path shards by environment; the only exchange steps are (1) the flat PPO gradient, (2) the KL statistic that drives
the adaptive learning rate, (3) sum / sum-of-squares / count of the advantages.  These helpers are device agnostic
(NCCL on GPUs, gloo in the CPU tests); the kernels themselves never communicate.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_info() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_envs(total_envs: int, rank: int | None = None, world: int | None = None) -> tuple[int, int]:
    """Rank r owns envs [start, start + count): equal contiguous shards (total must divide evenly so that
    mean-of-rank-means == global mean, SURVEY.md 8e 'Parity statement')."""
    r, w = world_info()
    rank = r if rank is None else rank
    world = w if world is None else world
    if total_envs % world != 0:
        raise ValueError(f"{total_envs} envs do not shard evenly over {world} ranks")
    count = total_envs // world
    return rank * count, count


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    _, world = world_info()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def average_gradients_(flat_grads: torch.Tensor) -> float:
    """SUM all-reduce of the flat gradient buffer; returns the scale (1/W) the fused clip+Adam kernel applies, so the
    division costs no extra pass."""
    _, world = world_info()
    allreduce_sum_(flat_grads)
    return 1.0 / world


def reduce_adv_stats_(stats: torch.Tensor) -> torch.Tensor:
    """stats = (sum, sum of squares, count, -) of the local advantages (float64) -> global."""
    return allreduce_sum_(stats)


def mean_and_unbiased_std(stats: torch.Tensor) -> tuple[float, float]:
    s, ss, n = float(stats[0]), float(stats[1]), float(stats[2])
    mean = s / n
    var = max((ss - s * mean) / (n - 1.0), 0.0)
    return mean, var ** 0.5


def global_kl_mean_(local_kl_mean: torch.Tensor) -> torch.Tensor:
    """Mean over ranks of the per-rank KL means (equal shard sizes) so that every rank takes the same lr decision."""
    _, world = world_info()
    allreduce_sum_(local_kl_mean)
    local_kl_mean /= world
    return local_kl_mean


def assert_same_on_all_ranks(t: torch.Tensor, what: str = "tensor"):
    _, world = world_info()
    if world == 1:
        return
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if not torch.equal(lo, hi):
        raise AssertionError(f"{what} differs across ranks")
