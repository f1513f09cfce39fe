"""GAE kernel (rollout_storage.py:122-136) against the oracle; tolerance 1e-5 relative on returns,
1e-5/1e-5 on the normalised advantages (fp32; the scan re-associates the recurrence)."""
import numpy as np
import pytest
import torch

from _util import assert_close
from oracle.ppo_oracle import gae_returns

pytestmark = pytest.mark.gpu


def run_cuda_gae(lib, dev, rewards, values, dones, last_values, gamma, lam, fused=False):
    from isaac_b200.algo.rollout_storage import gae_compute_returns
    T, N = rewards.shape[:2]
    r, v, d, lv = (x.to(dev).contiguous() for x in (rewards, values, dones, last_values))
    ret, adv = torch.empty_like(r), torch.empty_like(r)
    gae_compute_returns(r, v, d, lv, ret, adv, gamma, lam, fused=fused)
    torch.cuda.synchronize()
    return ret.cpu(), adv.cpu()


@pytest.fixture
def gae_kernel(lib, request):
    """Force one of the two scan kernels (the library picks by shard width: serial-in-time from 8192 envs up)."""
    lib.hb_set_option(b"gae_serial_min_envs", 1 if request.param == "serial" else 1 << 30)
    yield request.param
    lib.hb_set_option(b"gae_serial_min_envs", 8192)


@pytest.mark.parametrize("T,N", [(24, 64), (60, 257), (1, 33), (24, 4096), (100, 40), (24, 8192 + 5), (7, 9000), (33, 16385),
                                 (24, 16384), (64, 31), (2, 1)])
def test_gae_single_launch_matches_oracle(lib, cuda_device, T, N):
    """hb_gae_fused (what single-GPU compute_returns launches): one cooperative launch; T > 64 takes the documented
    two-kernel fall-back through the same entry point.  Run twice: the launch re-arms its own scratch."""
    g = torch.Generator().manual_seed(T * 1000 + N)
    rewards = torch.rand(T, N, 1, generator=g)
    values = torch.randn(T, N, 1, generator=g)
    dones = (torch.rand(T, N, 1, generator=g) < 0.05).byte()
    last = torch.randn(N, 1, generator=g)
    want_ret, want_adv = gae_returns(rewards, values, dones, last, 0.994, 0.9)
    for rep in range(2):
        ret, adv = run_cuda_gae(lib, cuda_device, rewards, values, dones, last, 0.994, 0.9, fused=True)
        if T <= 100:
            assert torch.equal(ret, want_ret), "the reference's own loop and operation order, no FMA contraction: the same bits"
        assert_close("returns", ret.numpy(), want_ret.numpy(), rtol=1e-5, atol=1e-5)
        assert_close("advantages", adv.numpy(), want_adv.numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("gae_kernel", ["warp-scan", "serial"], indirect=True)
@pytest.mark.parametrize("T,N", [(24, 64), (60, 257), (1, 33), (24, 4096), (100, 40), (24, 8192 + 5), (7, 9000)])
def test_gae_matches_oracle(lib, cuda_device, T, N, gae_kernel):
    g = torch.Generator().manual_seed(T * 1000 + N)
    rewards = torch.rand(T, N, 1, generator=g)
    values = torch.randn(T, N, 1, generator=g)
    dones = (torch.rand(T, N, 1, generator=g) < 0.05).byte()
    last = torch.randn(N, 1, generator=g)
    want_ret, want_adv = gae_returns(rewards, values, dones, last, 0.994, 0.9)
    ret, adv = run_cuda_gae(lib, cuda_device, rewards, values, dones, last, 0.994, 0.9)
    assert_close("returns", ret.numpy(), want_ret.numpy(), rtol=1e-5, atol=1e-5)
    if gae_kernel == "serial":      # the reference's own loop and operation order, no FMA contraction: the same bits
        assert torch.equal(ret, want_ret)
    if T * N > 1:
        assert_close("advantages", adv.numpy(), want_adv.numpy(), rtol=1e-5, atol=1e-5)


def test_gae_properties_full_size(lib, cuda_device):
    """65536 x 24: normalised advantages have zero mean / unit unbiased std; a done cuts the recursion."""
    T, N = 24, 65536
    g = torch.Generator().manual_seed(1)
    rewards = torch.rand(T, N, 1, generator=g)
    values = torch.randn(T, N, 1, generator=g)
    dones = (torch.rand(T, N, 1, generator=g) < 0.01).byte()
    last = torch.randn(N, 1, generator=g)
    ret, adv = run_cuda_gae(lib, cuda_device, rewards, values, dones, last, 0.99, 0.95, fused=True)
    assert abs(float(adv.double().mean())) < 1e-5 and abs(float(adv.double().std()) - 1.0) < 1e-4
    term = dones.bool()
    assert_close("terminal step return", ret[term].numpy(), rewards[term].numpy(), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("threads", [32, 64, 128, 256])
def test_gae_block_widths_agree(lib, cuda_device, threads):
    """The single-launch kernel at every block width ("gae_threads"; the library picks 64 or 256 by shard width): returns
    are the same bits (one thread per env, the reference's operation order), the normalised advantages agree to fp32
    rounding of the statistics (fp64 atomics in a different order).  Ragged shard (N not a multiple of any width)."""
    T, N = 24, 4096 + 37
    g = torch.Generator().manual_seed(threads)
    rewards, values = torch.rand(T, N, 1, generator=g), torch.randn(T, N, 1, generator=g)
    dones, last = (torch.rand(T, N, 1, generator=g) < 0.05).byte(), torch.randn(N, 1, generator=g)
    want_ret, want_adv = gae_returns(rewards, values, dones, last, 0.994, 0.9)
    assert lib.hb_set_option(b"gae_threads", threads) == 0
    try:
        for rep in range(2):          # twice: the launch re-arms its own scratch (8 slot pairs + ticket)
            ret, adv = run_cuda_gae(lib, cuda_device, rewards, values, dones, last, 0.994, 0.9, fused=True)
            assert torch.equal(ret, want_ret)
            assert_close("advantages", adv.numpy(), want_adv.numpy(), rtol=1e-5, atol=1e-5)
    finally:
        lib.hb_set_option(b"gae_threads", 0)
