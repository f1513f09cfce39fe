"""Host-side logic of the N>1 path on CPU: world_size-2 gloo (no GPU)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from isaac_b200.parallel import AdvantageStatsReducer, GradReducer, shard_range
        g = torch.Generator().manual_seed(0)
        # the "global" rollout: advantages [T, N] and per-sample gradient contributions [B, P]
        T, N, P = 6, 8, 5
        adv = torch.randn(T, N, generator=g, dtype=torch.float64)
        contrib = torch.randn(T * N, P, generator=g)
        lo, hi = shard_range(N, rank, world)
        local_adv = adv[:, lo:hi]
        # --- global advantage normalisation from sharded (sum, sum sq) ---
        stats = torch.stack((local_adv.sum(), (local_adv ** 2).sum()))
        count = AdvantageStatsReducer()(stats, local_adv.numel())
        mean = stats[0] / count
        std = ((stats[1] - count * mean * mean) / (count - 1)).sqrt()
        assert count == adv.numel()
        assert abs(mean - adv.mean()) < 1e-12 and abs(std - adv.std()) < 1e-12
        # --- gradient all-reduce: local grads carry 1/global_mb, the sum is the global-batch mean gradient ---
        rows = torch.arange(T * N).view(T, N)[:, lo:hi].flatten()
        flat_grad = contrib[rows].sum(0) / (T * N)
        loss_stats = torch.tensor([float(len(rows)), 1.0, 2.0, 3.0], dtype=torch.float64)
        GradReducer()(flat_grad, loss_stats)
        assert torch.allclose(flat_grad, contrib.mean(0), atol=1e-6)
        assert loss_stats[0] == T * N and loss_stats[2] == 2.0 * world
        # --- the one-collective form: the statistics ride behind the gradient as fp32 ---
        wire = torch.zeros(P + 8)
        wire[:P] = contrib[rows].sum(0) / (T * N)
        stats2 = torch.tensor([float(len(rows)), 1.0, 2.0, 3.0], dtype=torch.float64)
        GradReducer()(wire[:P], stats2, wire)
        assert torch.equal(wire[:P], flat_grad) and torch.equal(stats2, loss_stats)
        out[rank] = True
    finally:
        dist.destroy_process_group()


def test_reducers_world_size_2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_shard_range_errors():
    from isaac_b200.parallel import shard_range
    assert shard_range(65536, 3, 8) == (24576, 32768)
    with pytest.raises(ValueError):
        shard_range(100, 0, 8)
