"""Multi-GPU data parallelism for the hector hot path (the reference is single-device).

One process per GPU (torchrun), `torch.distributed` over NCCL/NVLink for the plumbing:

  * env stage and GAE scan: rank r owns envs [r*N/G, (r+1)*N/G) with its own gym tensors, history and
    RolloutStorage — no data-path collective at all (weak scaling);
  * advantage normalisation is global in the reference (rollout_storage.py:136): (sum, sum of squares) of
    the raw advantages, 2 doubles, are all-reduced between the two GAE passes;
  * PPO update: every rank runs minibatches of mb/G local samples; gradients are scaled by 1/(global
    minibatch) in the loss head, so ONE all-reduce(sum) of the flat packed gradient buffer (1.52 M floats,
    6.1 MB) per optimizer step gives the global-batch gradient; clip + Adam are then identical on all ranks;
  * the KL sum that drives the adaptive learning rate (ppo.py:140-148) is all-reduced with the loss
    statistics (4 doubles), so every rank takes the same branch.

The reducers only need `dist.all_reduce`, so the same code is exercised on CPU tensors with the gloo
backend in tests/test_parallel.py.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradReducer:
    """Sums the flat gradient buffer and the fp64 loss statistics over all ranks (in place)."""

    def __init__(self, group=None):
        self.group = group
        self.world_size = dist.get_world_size(group)

    def __call__(self, flat_grad: torch.Tensor, stats: torch.Tensor, wire: torch.Tensor = None) -> None:
        """`wire`: optional fp32 buffer whose head IS flat_grad and which has stats.numel() spare floats behind it
        (ActorCritic._grad_wire): the statistics then ride in the gradient's all-reduce as fp32 - one collective
        per optimizer step instead of two (12 us per step at 2 GPUs).  Every rank reads back the same reduced
        values, so the adaptive-KL branch stays identical on all ranks."""
        if self.world_size == 1:
            return
        if wire is not None and wire.data_ptr() == flat_grad.data_ptr() and wire.numel() >= flat_grad.numel() + stats.numel():
            tail = wire[flat_grad.numel():flat_grad.numel() + stats.numel()]
            tail.copy_(stats)
            dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=self.group)
            stats.copy_(tail)
            return
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.group)


class AdvantageStatsReducer:
    """Global (sum, sum sq) of the raw advantages and the global sample count for `hb_gae_normalize_n`."""

    def __init__(self, group=None):
        self.group = group
        self.world_size = dist.get_world_size(group)

    def __call__(self, stats: torch.Tensor, local_count: int) -> int:
        if self.world_size == 1:
            return local_count
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.group)
        return local_count * self.world_size       # every rank holds the same [T, N/G] shard


def broadcast_parameters(actor_critic, src: int = 0, group=None) -> None:
    """Replicas must start from identical weights (rank `src`'s)."""
    dist.broadcast(actor_critic.flat, src=src, group=group)


_GOLDEN = 0x9E3779B97F4A7C15


def rank_seed(base: int, rank: int) -> int:
    """A distinct 64-bit generator key per rank (rank 0 keeps `base`): replicas that drew from the same key would
    sample identical action noise, observation noise, command resamples and reset poses for env i of every shard, and
    the all-reduced gradient would carry 1/G of the intended information."""
    return (int(base) ^ ((int(rank) * _GOLDEN) & 0xFFFFFFFFFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF


def attach_data_parallel(alg, group=None, env=None):
    """Turn a single-GPU `PPO` into one data-parallel replica (call after init_storage).  The action-sample generator
    of the replica is re-keyed with `rank_seed`; pass the replica's env as well (or call `env.seed(rank_seed(cfg.seed,
    rank))` yourself) so that its in-kernel draws differ between shards too."""
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
    alg.world_size = dist.get_world_size(group)
    rank = dist.get_rank(group)
    alg.grad_allreduce = GradReducer(group)
    if alg.storage is not None:
        alg.storage.reduce_stats = AdvantageStatsReducer(group)
    broadcast_parameters(alg.actor_critic, 0, group)
    if hasattr(alg, "seed") and hasattr(alg, "_eps_seed"):
        alg.seed(rank_seed(alg._eps_seed, rank))
    if env is not None:
        env.seed(rank_seed(env._rng_seed, rank))
    return alg


def shard_range(num_envs: int, rank: int, world: int):
    """Contiguous env shard of a rank; num_envs must divide evenly (hector configs: 4096..65536 over 1..8)."""
    if num_envs % world:
        raise ValueError(f"{num_envs} envs do not shard evenly over {world} GPUs")
    per = num_envs // world
    return rank * per, (rank + 1) * per
