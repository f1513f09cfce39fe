"""Runs a few PPO updates on synthetic rollouts (for ncu / timing): python scripts/profile_ppo.py [envs] [updates]"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
import bench
from isaac_b200.algo.actor_critic import ActorCritic
from isaac_b200.algo.ppo import PPO

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
updates = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
torch.manual_seed(5)
ac = ActorCritic(615, 1050, 10, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128], device=dev)
alg = PPO(ac, device=dev, **bench.PPO_CFG)
alg.init_storage(n, 24, [615], [1050], [10])
last = torch.randn(n, 1050, device=dev)
for r in range(updates + 1):
    bench.fill_storage(alg.storage, 100, dev)
    alg.storage.step = 24
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    alg.compute_returns(last)
    t1 = time.perf_counter()
    alg.update()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"update {r}: host enqueue {1e3*(t2-t0):.1f} ms (returns {1e3*(t1-t0):.1f}), total {1e3*(t3-t0):.1f} ms")
