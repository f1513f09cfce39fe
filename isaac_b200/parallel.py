"""Multi-GPU data parallelism for the hector hot path (the reference is single-device).

One process per GPU (torchrun), `torch.distributed` over NCCL/NVLink for the plumbing:

  * env stage and GAE scan: rank r owns envs [r*N/G, (r+1)*N/G) with its own gym tensors, history and
    RolloutStorage — no data-path collective at all (weak scaling);
  * advantage normalisation is global in the reference (rollout_storage.py:136): (sum, sum of squares) of
    the raw advantages, 2 doubles, are all-reduced between the two GAE passes;
  * PPO update: every rank runs minibatches of mb/G local samples; gradients are scaled by 1/(global
    minibatch) in the loss head, so the SUM over ranks of the flat packed gradient buffer (1.52 M floats,
    6.1 MB) is the global-batch gradient.  Default on CUDA (`PeerOptimizer`): the parameter and gradient buffers
    live in symmetric memory and ONE kernel per rank (hb_dp_optimizer_step) reduces the rank's 1/G slice over
    NVLink (multimem.ld_reduce where the switch offers multicast, peer loads otherwise), exchanges the norm and
    the loss / KL sums through peer mailboxes, applies clip + Adam to the slice (optimizer state sharded) and
    writes the new parameters to every replica.  Fallback (`GradReducer`, also what gloo tests run): one NCCL
    all-reduce of the gradient with the four loss sums riding behind it, then the local optimizer step;
  * either way the KL sum that drives the adaptive learning rate (ppo.py:140-148) is global, so every rank takes
    the same branch, and every rank's generators are re-keyed (`rank_seed`) so that shards draw different noise.

The reducers only need `dist.all_reduce`, so the same code is exercised on CPU tensors with the gloo
backend in tests/test_parallel.py; the peer-memory kernel is covered by tests/test_multi_gpu.py (2 and 8 GPUs)
and, for its arithmetic, tests/test_dp_emulation.py on one GPU.
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


class PeerOptimizer:
    """The plumbing of `hb_dp_optimizer_step` (include/hector_b200.h): the replica's flat parameter and gradient buffers
    move into symmetric memory (torch.distributed._symmetric_memory: every rank's allocation mapped into every rank's
    address space over NVLink / NVSwitch, plus the switch's multicast address where NVLS exists), and the peer pointers are
    handed to the kernel that replaces all-reduce + clip + Adam + broadcast.  No NCCL call on the update path."""

    def __init__(self, alg, group=None, use_multicast=True):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        group = group or dist.group.WORLD
        ac = alg.actor_critic
        dev = ac.device
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world > _lib.HB_DP_MAX_RANKS:
            raise ValueError(f"at most {_lib.HB_DP_MAX_RANKS} ranks")
        n = ac.flat.numel()
        self.params = symm_mem.empty(n, dtype=torch.float32, device=dev)
        self.grads = symm_mem.empty(ac._grad_wire.numel(), dtype=torch.float32, device=dev)
        self.mail = symm_mem.empty(_lib.HB_DP_MAX_RANKS * 8, dtype=torch.float64, device=dev)
        self.flags = symm_mem.empty(3 * _lib.HB_DP_MAX_RANKS, dtype=torch.int32, device=dev)
        self.params.copy_(ac.flat)
        self.grads.zero_(), self.mail.zero_(), self.flags.zero_()
        self.handles = [symm_mem.rendezvous(t, group) for t in (self.params, self.grads, self.mail, self.flags)]
        hp, hg, hm, hf = self.handles
        comm = _lib.DpComm()
        comm.world, comm.rank = world, rank
        for p in range(world):
            comm.param[p], comm.grad[p] = hp.buffer_ptrs[p], hg.buffer_ptrs[p]
            comm.mail[p], comm.flag[p] = hm.buffer_ptrs[p], hf.buffer_ptrs[p]
        self.multicast = bool(use_multicast and hp.multicast_ptr and hg.multicast_ptr)
        comm.param_mc = hp.multicast_ptr if self.multicast else None
        comm.grad_mc = hg.multicast_ptr if self.multicast else None
        self.comm, self.comm_ref = comm, C.byref(comm)
        self.world, self.rank, self.group = world, rank, group
        ac.rebind(self.params, self.grads)           # the GEMMs, the loss head and Adam now work in the shared buffers
        torch.cuda.synchronize(dev)
        dist.barrier(group)                          # every rank's flags / mailboxes are zero before the first step

    def slice_range(self, n):
        per = ((n // 4 + self.world - 1) // self.world) * 4
        return self.rank * per, min(n, (self.rank + 1) * per), per

    def gather_sharded(self, t: torch.Tensor) -> torch.Tensor:
        """Full copy of an optimizer-state buffer whose slices live on their owner ranks (checkpoints)."""
        n = t.numel()
        lo, hi, per = self.slice_range(n)
        mine = torch.zeros(per, device=t.device)
        mine[:hi - lo] = t[lo:hi]
        full = torch.empty(per * self.world, device=t.device)
        dist.all_gather_into_tensor(full, mine, group=self.group)
        return full[:n]


def attach_data_parallel(alg, group=None, env=None, fused=None, use_multicast=True):
    """Turn a single-GPU `PPO` into one data-parallel replica (call after init_storage).  `fused` (default: on CUDA):
    gradients are reduced, clipped, applied and the parameters re-broadcast by ONE kernel over peer memory
    (`PeerOptimizer` / hb_dp_optimizer_step); fused=False keeps the NCCL all-reduce + local optimizer step.  The
    action-sample generator of the replica is re-keyed with `rank_seed`; pass the replica's env as well (or call
    `env.seed(rank_seed(cfg.seed, rank))` yourself) so that its in-kernel draws differ between shards too."""
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
    alg.world_size = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if alg.storage is not None:
        alg.storage.reduce_stats = AdvantageStatsReducer(group)
    broadcast_parameters(alg.actor_critic, 0, group)
    if fused is None:
        fused = alg.actor_critic.flat.is_cuda and alg.world_size > 1
    if fused:
        alg.grad_allreduce = None
        alg.attach_peer_optimizer(PeerOptimizer(alg, group, use_multicast))
    else:
        alg.grad_allreduce = GradReducer(group)
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
