"""RolloutStorage's reference-facing surface on CPU tensors (no GPU, no kernels): shapes, the extra observation slot, and
the mini_batch_generator contract of rollout_storage.py:146-182."""
import torch

from isaac_b200.algo.rollout_storage import RolloutStorage


def test_layout_and_extra_slot():
    s = RolloutStorage(8, 4, [615], [1050], [10], device="cpu")
    assert s.observations.shape == (4, 8, 615) and s.privileged_observations.shape == (4, 8, 1050)
    assert s._observations.shape == (5, 8, 640) and s._privileged_observations.shape == (5, 8, 1056)   # T + 1 slots, 128-byte rows
    o, p = s.observation_slot(4)
    assert o.shape == (8, 615) and o.stride() == (640, 1) and p.shape == (8, 1050) and p.stride() == (1056, 1)
    o.fill_(3.0)
    assert float(s.observations.abs().sum()) == 0.0, "slot T lies outside the [T, N, *] view the update reads"
    for name in ("actions", "mu", "sigma"):
        assert getattr(s, name).shape == (4, 8, 10)
    for name in ("rewards", "values", "returns", "advantages", "actions_log_prob", "dones"):
        assert getattr(s, name).shape == (4, 8, 1)


def test_mini_batch_generator_contract():
    s = RolloutStorage(8, 4, [615], [1050], [10], device="cpu")
    for t in (s.observations, s.privileged_observations, s.actions, s.values, s.returns, s.advantages,
              s.actions_log_prob, s.mu, s.sigma):
        t.normal_()
    torch.manual_seed(0)
    batches = list(s.mini_batch_generator(4, num_epochs=3))
    torch.manual_seed(0)
    order = torch.randperm(32)
    assert len(batches) == 12
    flat = lambda t: t.flatten(0, 1)
    for e in range(3):                     # the SAME permutation in every epoch (reference quirk, :149)
        for i in range(4):
            idx = order[8 * i:8 * i + 8]
            b = batches[4 * e + i]
            want = (s.observations, s.privileged_observations, s.actions, s.values, s.advantages, s.returns,
                    s.actions_log_prob, s.mu, s.sigma)
            for got, src in zip(b[:9], want):
                assert torch.equal(got, flat(src)[idx])
            assert b[9] == (None, None) and b[10] is None
