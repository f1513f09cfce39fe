"""A/B of a library option (hb_set_option) in ONE process: PPO act + record per rollout step and ms per update, interleaved.

    python scripts/ab_option.py <option> <value> <value> [...] [--envs N]
"""
import os
import sys
from types import SimpleNamespace

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch

import bench
from isaac_b200 import _lib
from isaac_b200.algo.actor_critic import ActorCritic
from isaac_b200.algo.ppo import PPO

argv = sys.argv[1:]
n = 4096
if "--envs" in argv:
    i = argv.index("--envs")
    n = int(argv[i + 1])
    del argv[i:i + 2]
option, values = argv[0], [int(v) for v in argv[1:]]
dev = torch.device("cuda:0")
lib = _lib.load(check_device=True)
torch.manual_seed(5)
ac = ActorCritic(615, 1050, 10, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128], device=dev)
alg = PPO(ac, device=dev, **bench.PPO_CFG)
alg.init_storage(n, bench.T_GAE, [615], [1050], [10])
last = torch.randn(n, 1050, device=dev)
rew, done = torch.rand(n, device=dev), torch.rand(n, device=dev) < 0.005
infos = {"time_outs": torch.rand(n, device=dev) < 0.0004}
stream = torch.cuda.current_stream(dev)
T = bench.T_GAE


def rollout_ms():
    best = 1e9
    for r in range(3):
        alg.storage.clear()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for k in range(T):
            alg.act(*alg.storage.observation_slot(k))
            alg.process_env_step(rew, done, infos)
        b.record(stream)
        b.synchronize()
        best = min(best, a.elapsed_time(b) / T)
    return best


res = {v: [] for v in values}
for rnd in range(3):
    for v in values:
        _lib.check(lib.hb_set_option(option.encode(), v), option)
        alg._update_graphs.clear(), alg._act_graphs.clear()
        alg.storage.observations.normal_(), alg.storage.privileged_observations.normal_()
        roll = rollout_ms()
        ms = bench.time_updates(alg, SimpleNamespace(), dev, n, 1, 0, last, 4)
        res[v].append((roll * 1e3, ms))
for v in values:
    print(f"{option}={v}: act+record us/step " + " ".join(f"{r:.1f}" for r, _ in res[v]) + " | ms/update " + " ".join(f"{m:.3f}" for _, m in res[v]), flush=True)
