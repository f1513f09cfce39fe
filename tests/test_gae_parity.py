"""GAE kernel (rollout_storage.py:122-136) against the oracle; tolerance 1e-5 relative on returns,
1e-5/1e-5 on the normalised advantages (fp32; the scan re-associates the recurrence)."""
import numpy as np
import pytest
import torch

from _util import assert_close
from oracle.ppo_oracle import gae_returns

pytestmark = pytest.mark.gpu


def run_cuda_gae(lib, dev, rewards, values, dones, last_values, gamma, lam, fused=True):
    """fused: the single cooperative launch (one GPU); otherwise scan + normalise as two kernels, which is what a
    multi-GPU caller runs around its all-reduce of the statistics."""
    from isaac_b200.algo.rollout_storage import gae_compute_returns
    T, N = rewards.shape[:2]
    r, v, d, lv = (x.to(dev).contiguous() for x in (rewards, values, dones, last_values))
    ret, adv = torch.empty_like(r), torch.empty_like(r)
    gae_compute_returns(r, v, d, lv, ret, adv, gamma, lam, reduce_stats=None if fused else (lambda stats, count: count))
    torch.cuda.synchronize()
    return ret.cpu(), adv.cpu()


@pytest.mark.parametrize("fused", [True, False], ids=["one-launch", "two-kernels"])
@pytest.mark.parametrize("T,N", [(24, 64), (60, 257), (1, 33), (24, 4096), (100, 40), (24, 65536 + 7), (400, 20000)])
def test_gae_matches_oracle(lib, cuda_device, T, N, fused):
    """(24, 65543): more tiles than resident CTAs - several tiles per CTA in the one-launch kernel, ragged last tile;
    (400, 20000): does not fit one resident grid - the one-launch entry point declines and the wrapper falls back."""
    g = torch.Generator().manual_seed(T * 1000 + N)
    rewards = torch.rand(T, N, 1, generator=g)
    values = torch.randn(T, N, 1, generator=g)
    dones = (torch.rand(T, N, 1, generator=g) < 0.05).byte()
    last = torch.randn(N, 1, generator=g)
    want_ret, want_adv = gae_returns(rewards, values, dones, last, 0.994, 0.9)
    ret, adv = run_cuda_gae(lib, cuda_device, rewards, values, dones, last, 0.994, 0.9, fused)
    assert_close("returns", ret.numpy(), want_ret.numpy(), rtol=1e-5, atol=1e-5)
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
    ret, adv = run_cuda_gae(lib, cuda_device, rewards, values, dones, last, 0.99, 0.95)
    assert abs(float(adv.double().mean())) < 1e-5 and abs(float(adv.double().std()) - 1.0) < 1e-4
    term = dones.bool()
    assert_close("terminal step return", ret[term].numpy(), rewards[term].numpy(), rtol=1e-6, atol=1e-6)


def test_one_launch_gae_is_reproducible_and_declines_what_it_cannot_hold(lib, cuda_device):
    """The statistics are summed in a fixed order: two runs agree bit for bit.  A shape that cannot be resident at
    once comes back as HB_ERR_UNSUPPORTED (the two-kernel path takes it), never as a hang."""
    from isaac_b200 import _lib
    dev = cuda_device
    T, N = 24, 16384 + 5
    g = torch.Generator().manual_seed(7)
    r, v = torch.rand(T, N, generator=g).to(dev), torch.randn(T, N, generator=g).to(dev)
    d, lv = (torch.rand(T, N, generator=g) < 0.02).byte().to(dev), torch.randn(N, generator=g).to(dev)
    outs = []
    for _ in range(2):
        ret, adv = torch.empty_like(r), torch.empty_like(r)
        work = torch.full((_lib.HB_GAE_WORK_DOUBLES,), float("nan"), dtype=torch.float64, device=dev)   # needs no initialisation
        rc = lib.hb_gae_returns_normalized(r.data_ptr(), v.data_ptr(), d.data_ptr(), lv.data_ptr(), ret.data_ptr(),
                                           adv.data_ptr(), work.data_ptr(), T, N, 0.99, 0.95, None)
        torch.cuda.synchronize()
        assert rc == 0, lib.hb_last_error()
        outs.append((ret, adv, work[:2].clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    raw = (outs[0][0] - v).double()
    np.testing.assert_allclose(outs[0][2].cpu().numpy(), [float(raw.sum()), float((raw * raw).sum())], rtol=1e-6)
    rc = lib.hb_gae_returns_normalized(r.data_ptr(), v.data_ptr(), d.data_ptr(), lv.data_ptr(), ret.data_ptr(),
                                       adv.data_ptr(), work.data_ptr(), 2000, 8, 0.99, 0.95, None)
    assert rc == -3 and b"resident" in lib.hb_last_error()
