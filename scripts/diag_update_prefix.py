"""In-situ duration of every launch of one PPO minibatch step: the step's launches are captured as graphs of their first
k launches (k = 1 .. all); time(k) - time(k - 1) is launch k's cost where it actually runs (its inputs as warm or cold in
L2 as the preceding launches leave them), unlike an ncu capture (cold, serialised).

    python scripts/diag_update_prefix.py [envs]
"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch

import bench
from isaac_b200 import _lib
from isaac_b200.algo import actor_critic as acm
from isaac_b200.algo.actor_critic import ActorCritic
from isaac_b200.algo.ppo import PPO

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
lib = _lib.load(check_device=True)
torch.manual_seed(5)
ac = ActorCritic(615, 1050, 10, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128], device=dev)
alg = PPO(ac, device=dev, **bench.PPO_CFG)
alg.init_storage(n, bench.T_GAE, [615], [1050], [10])
bench.fill_storage(alg.storage, 100, dev)
alg.storage.step = bench.T_GAE
alg.compute_returns(torch.randn(n, 1050, device=dev))
alg.update()                                  # eager warm-up (module load, attributes)
bench.fill_storage(alg.storage, 101, dev)
alg.storage.step = bench.T_GAE
alg.compute_returns(torch.randn(n, 1050, device=dev))
alg.prepare_minibatches(None)
alg._per_update.zero_()
torch.cuda.synchronize()

names, limit, count = [], [10 ** 9], [0]
orig_gemm2, orig_head, orig_opt = acm.gemm2, lib.hb_ppo_head_fused, lib.hb_optimizer_step


def gate(label, fn):
    def wrapped(*a, **kw):
        count[0] += 1
        if len(names) < count[0]:
            names.append(label(*a, **kw))
        if count[0] <= limit[0]:
            return fn(*a, **kw)
        return 0
    return wrapped


def gemm_label(lib_, st, s0, kw0, s1, kw1):
    kind = {2: "forward", 3: "data gradient", 4: "weight gradient"}.get(kw0.get("epilogue"), "gemm")
    return f"{kind:15s} critic {kw0['M']}x{kw0['N']}x{kw0['K']} + actor {kw1['M']}x{kw1['N']}x{kw1['K']}"


acm.gemm2 = gate(gemm_label, orig_gemm2)
lib.hb_ppo_head_fused = gate(lambda *a, **k: "head_fused", orig_head)
lib.hb_optimizer_step = gate(lambda *a, **k: "optimizer_step", orig_opt)


def step():
    count[0] = 0
    alg.minibatch_gradients(0)
    alg.optimizer_step(1)


step()                                          # eager once: fills `names`
torch.cuda.synchronize()
total = len(names)
stream = torch.cuda.current_stream(dev)
times = []
for k in range(1, total + 1):
    limit[0] = k
    g = _lib.LaunchGraph(dev).record(lambda st: step())
    for _ in range(3):
        g.replay(stream.cuda_stream)
    reps, best = 30, 1e9
    for _ in range(4):                          # best of 4 series: a series that starts on ramping clocks reads long
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            g.replay(stream.cuda_stream)
        b.record(stream)
        b.synchronize()
        best = min(best, a.elapsed_time(b) * 1e3 / reps)
    times.append(best)
prev = 0.0
print(f"one minibatch step at {n} envs x {bench.T_GAE} steps ({n * bench.T_GAE // 4} rows): marginal time of each launch in place")
for k, (nm, t) in enumerate(zip(names, times), 1):
    print(f"  {k:2d} {nm:70s} {t - prev:8.1f} us   (prefix {t:8.1f} us)")
    prev = t
