"""Data-parallel PPO math on ONE GPU: G replicas on cuda:0, each holding its env shard, with an in-process reducer that
sums their wire buffers (flat gradient + loss statistics) where NCCL's all-reduce would - what runs on the driver's
1-GPU box.  The parity definition is SURVEY.md §8e's: the G-replica result equals the single-replica result on the
concatenated batch when replica r's minibatch k is the slice of the global minibatch k that falls in r's shard.
(The real 2-rank NCCL run of the same comparison is tests/test_multi_gpu.py.)"""
import numpy as np
import pytest
import torch

from _dp_worker import CFG, N, T, global_rollout, make_alg
from isaac_b200 import _lib
from isaac_b200.parallel import rank_seed, shard_range

pytestmark = pytest.mark.gpu


def load_shard(alg, R, lo, hi, dev, reduce_stats=None):
    s = alg.storage
    s.observations.copy_(R["obs"][:, lo:hi]), s.privileged_observations.copy_(R["priv"][:, lo:hi])
    s.actions.copy_(R["actions"][:, lo:hi]), s.mu.copy_(R["mu"][:, lo:hi]), s.sigma.fill_(1.0)
    s.rewards.copy_(R["rewards"][:, lo:hi]), s.values.copy_(R["values"][:, lo:hi]), s.dones.copy_(R["dones"][:, lo:hi])
    s.actions_log_prob.copy_(R["logp"][:, lo:hi])
    s.reduce_stats = reduce_stats
    s.compute_returns(R["last"][lo:hi].to(dev), CFG["gamma"], CFG["lam"])


@pytest.mark.parametrize("world", [2, 4])
def test_emulated_data_parallel_update_matches_single_replica(lib, cuda_device, world):
    dev = cuda_device
    R = global_rollout()
    nl = N // world
    reps = [make_alg(dev, nl) for _ in range(world)]
    for r, alg in enumerate(reps):
        alg.world_size = world                       # 1 / global minibatch in the loss head, global KL count
        assert torch.equal(alg.actor_critic.flat, reps[0].actor_critic.flat)
    # ---- advantage normalisation: (sum, sum sq) over all shards (rollout_storage.py:136 is global) ----
    stats = []
    for r, alg in enumerate(reps):
        lo, hi = shard_range(N, r, world)
        box = []

        def grab(s, c, box=box):         # the shard's own (sum, sum sq); no reduction yet
            box.append(s.clone())
            return c

        load_shard(alg, R, lo, hi, dev, reduce_stats=grab)
        stats.append(box[0])
    total = torch.stack(stats).sum(0)
    for r, alg in enumerate(reps):
        lo, hi = shard_range(N, r, world)
        load_shard(alg, R, lo, hi, dev, reduce_stats=lambda s, c: (s.copy_(total), c * world)[1])
    ref = make_alg(dev, N)
    init = {k: v.clone() for k, v in ref.actor_critic.state_dict().items()}
    load_shard(ref, R, 0, N, dev)
    for r, alg in enumerate(reps):
        lo, hi = shard_range(N, r, world)
        torch.testing.assert_close(alg.storage.advantages, ref.storage.advantages[:, lo:hi], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(alg.storage.returns, ref.storage.returns[:, lo:hi], rtol=1e-5, atol=1e-6)
    # ---- minibatch index tapes: global minibatch k = rank-wise concatenation of the local minibatches k ----
    mbl = (T * nl) // CFG["num_mini_batches"]
    perms = [torch.randperm(T * nl, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    chunks = []
    for k in range(CFG["num_mini_batches"]):
        for r in range(world):
            p = perms[r][k * mbl:(k + 1) * mbl]
            chunks.append((p // nl) * N + (p % nl + r * nl))
    global_perm = torch.cat(chunks)
    for alg, p in zip(reps, perms):
        alg.prepare_minibatches(p)
    ref.prepare_minibatches(global_perm)

    def reduce_wires():
        """What GradReducer does with NCCL: sum of every replica's [gradient | loss statistics] wire, back to all."""
        for alg in reps:
            n = alg.actor_critic.grad.numel()
            alg.actor_critic._grad_wire[n:n + 4].copy_(alg._stats)
        wire = torch.stack([alg.actor_critic._grad_wire for alg in reps]).sum(0)
        for alg in reps:
            alg.actor_critic._grad_wire.copy_(wire)
            n = alg.actor_critic.grad.numel()
            alg._stats.copy_(wire[n:n + 4])

    # (1) one minibatch: summed replica gradients == gradient of the concatenated minibatch
    for alg in reps:
        alg.minibatch_gradients(0)
    reduce_wires()
    ref.minibatch_gradients(0)
    g_dp, g_ref = reps[0].actor_critic.grad.double(), ref.actor_critic.grad.double()
    rel = float((g_dp - g_ref).norm() / g_ref.norm())
    assert rel < 1e-4, f"gradient DP vs single: rel L2 {rel:.3e}"
    np.testing.assert_allclose(reps[0]._stats.cpu().numpy(), ref._stats.cpu().numpy(), rtol=1e-5)
    for alg in reps + [ref]:
        alg.actor_critic.grad.zero_(), alg._stats.zero_()
    # (2) the whole update, adaptive schedule: every replica takes the same learning-rate branch as the single run
    for alg in reps + [ref]:
        alg.schedule, alg.desired_kl = "adaptive", 0.01
        alg._per_update.zero_()
        alg._opt_i64[_lib.OPT_STEPS_IN_UPDATE:_lib.OPT_STEPS_IN_UPDATE + 2].zero_()
    for _ in range(CFG["num_learning_epochs"]):
        for i in range(CFG["num_mini_batches"]):
            for alg in reps:
                alg.minibatch_gradients(i)
            reduce_wires()
            ref.minibatch_gradients(i)
            for alg in reps + [ref]:
                alg.optimizer_step(1)
    torch.cuda.synchronize()
    for alg in reps[1:]:
        assert torch.equal(alg.actor_critic.flat, reps[0].actor_critic.flat), "replicas must stay bit-identical"
    steps = CFG["num_learning_epochs"] * CFG["num_mini_batches"]
    tr = lambda a: a._opt[_lib.OPT_TRACE:_lib.OPT_TRACE + 2 * steps].view(steps, 2).cpu().numpy()
    np.testing.assert_allclose(tr(reps[0])[:, 1], tr(ref)[:, 1], rtol=1e-12)           # learning rate per step
    np.testing.assert_allclose(tr(reps[0])[:, 0], tr(ref)[:, 0], rtol=1e-4, atol=1e-8)  # global mean KL per step
    a, b = reps[0].actor_critic.state_dict(), ref.actor_critic.state_dict()
    worst = max(float((a[k] - b[k]).double().norm() / (b[k] - init[k]).double().norm().clamp_min(1e-30)) for k in a)
    print(f"worst relative update difference, emulated DP{world} vs single: {worst:.3e}")
    # Adam divides by sqrt(v): elements with near-zero gradient amplify the (1e-5-level) summation-order differences
    assert worst < 0.07


def test_rank_seeds_give_distinct_draws(lib, cuda_device):
    """replicas must not draw identical numbers.  attach_data_parallel re-keys each rank's generators with
    rank_seed(); two replicas keyed that way sample different action noise and different env draws, the same key
    reproduces."""
    from isaac_b200.synthetic import make_tape
    from test_env_parity import make_cuda_env
    dev = cuda_device
    assert len({rank_seed(5, r) for r in range(8)}) == 8 and rank_seed(5, 0) == 5
    n = 256
    draws = []
    for key in (rank_seed(5, 0), rank_seed(5, 1), rank_seed(5, 0)):
        alg = make_alg(dev, n)
        alg.seed(key)
        obs, priv = torch.zeros(n, 615, device=dev), torch.zeros(n, 1050, device=dev)
        a = alg.act(obs, priv).clone()
        tape = make_tape(n, 2, seed=3, fall_prob=0.3)
        env, phys = make_cuda_env(tape, dev)
        env.seed(key)
        phys.load_frame(tape.physics[1].to(dev))
        o = env.step(torch.zeros(n, 10, device=dev))[0].clone()        # observation noise + reset draws from the device generator
        draws.append((a, o, env.dof_pos.clone()))
    assert torch.equal(draws[0][0], draws[2][0]) and torch.equal(draws[0][1], draws[2][1])
    assert not torch.equal(draws[0][0], draws[1][0]), "action samples identical across ranks"
    assert not torch.equal(draws[0][1], draws[1][1]), "observation noise identical across ranks"
    assert not torch.equal(draws[0][2], draws[1][2]), "reset poses identical across ranks"


def test_seed_after_graph_capture_takes_effect(lib, cuda_device):
    """seed() used to be ignored by already-captured graphs (the key was a by-value kernel argument).  The keys
    now live in device memory: graphs captured under one key draw from the new key after seed()."""
    from isaac_b200.synthetic import make_tape
    from test_env_parity import make_cuda_env
    dev = cuda_device
    n, t = 256, 3
    # ---- PPO.act graphs ----
    obs, priv = torch.randn(n, 615, device=dev), torch.randn(n, 1050, device=dev)
    warm, fresh = make_alg(dev, n), make_alg(dev, n)
    for alg in (warm, fresh):
        alg.init_storage(n, t, [615], [1050], [10])
    warm.seed(111)
    for k in range(t):                  # captures one act graph per slot under key 111
        warm.act(obs, priv)
        warm.process_env_step(torch.zeros(n, device=dev), torch.zeros(n, dtype=torch.bool, device=dev), {})
    assert len(warm._act_graphs) == t
    warm.storage.clear()
    warm.seed(31), fresh.seed(31)
    for k in range(t):
        a, b = warm.act(obs, priv).clone(), fresh.act(obs, priv).clone()
        assert torch.equal(a, b), f"act graph of slot {k} kept its old key"
        for alg in (warm, fresh):
            alg.process_env_step(torch.zeros(n, device=dev), torch.zeros(n, dtype=torch.bool, device=dev), {})
    # ---- env step graphs ----
    tape = make_tape(n, 4, seed=8, fall_prob=0.05)
    env_a, phys_a = make_cuda_env(tape, dev)
    env_b, phys_b = make_cuda_env(tape, dev)
    env_a.seed(222)
    env_a.enable_cuda_graph()           # the step graphs are captured here, under key 222
    env_a.seed(77)
    env_b.seed(77)
    env_b.enable_cuda_graph()
    for k in range(1, 4):
        fr = tape.physics[k].to(dev)
        phys_a.load_frame(fr), phys_b.load_frame(fr)
        act = tape.noise[k].actions.to(dev)
        oa, ob = env_a.step(act), env_b.step(act)
        for x, y, name in zip(oa[:4], ob[:4], ("obs", "priv", "rew", "reset")):
            assert torch.equal(x, y), f"{name} differs at step {k}: a captured step graph kept its old key"
    assert torch.equal(env_a.dof_state, env_b.dof_state) and torch.equal(env_a.commands, env_b.commands)
