"""Per-GEMM timing of one PPO minibatch step (forward, data-gradient and weight-gradient GEMMs of both MLPs).

    python scripts/bench_gemm.py [minibatch_rows] [--opt name=value ...]

Every hb_gemm_tf32 call of `PPO.minibatch_gradients` is recorded, then replayed alone from a one-node CUDA graph
(L2 flushed before each replay) and timed with CUDA events."""
import ctypes as C
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

args = [a for a in sys.argv[1:] if not a.startswith("--")]
mb = int(args[0]) if args else 24576
dev = torch.device("cuda:0")
lib = _lib.load(check_device=True)
for i, a in enumerate(sys.argv):
    if a == "--opt":
        k, v = sys.argv[i + 1].split("=")
        _lib.check(lib.hb_set_option(k.encode(), int(v)), a)
if "--no-gemm-pdl" in sys.argv:
    lib.hb_set_option(b"gemm_pdl", 0)
if "--no-pair" in sys.argv:
    lib.hb_gemm_set_pair_mode(0)
torch.manual_seed(5)
ac = ActorCritic(615, 1050, 10, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128], device=dev)
alg = PPO(ac, device=dev, **bench.PPO_CFG)
n_envs = mb * 4 // 24
alg.init_storage(n_envs, 24, [615], [1050], [10])
bench.fill_storage(alg.storage, 100, dev)
alg.storage.step = 24
alg.compute_returns(torch.randn(n_envs, 1050, device=dev))
alg.prepare_minibatches(None)

calls = []
orig = acm.gemm


def rec(lib_, st, **kw):
    calls.append(dict(kw))
    return orig(lib_, st, **kw)


acm.gemm = rec
head_args = []
head_fn = lib.hb_ppo_head_fused
lib.hb_ppo_head_fused = lambda *a: (head_args.append(a), head_fn(*a))[1]
alg.minibatch_gradients(0)
acm.gemm = orig
lib.hb_ppo_head_fused = head_fn
torch.cuda.synchronize()

flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
sink = torch.zeros(1, device=dev)


def time_graph(fn, reps=5):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn(torch.cuda.current_stream(dev).cuda_stream)
    tot = 0.0
    for r in range(reps + 2):
        flush.fill_(float(r))
        sink.copy_(flush.sum())
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        b.synchronize()
        if r >= 2:
            tot += a.elapsed_time(b)
    return tot / reps * 1e3


total_us, total_flop = 0.0, 0.0
kinds = {(0, 0): "fwd  ", (0, 1): "dgrad", (1, 1): "wgrad"}
for kw in calls:
    us = time_graph(lambda st: orig(lib, st, **kw))
    flop = 2.0 * kw["M"] * kw["N"] * kw["K"]
    total_us += us
    total_flop += flop
    kind = kinds[(kw.get("a_mn_major", 0), kw.get("b_mn_major", 0))]
    print(f"{kind} M={kw['M']:6d} N={kw['N']:5d} K={kw['K']:6d} split={kw.get('split_k', 1):3d}  {us:8.1f} us  {flop / us / 1e6:7.1f} TFLOP/s")
if head_args:
    us = time_graph(lambda st: head_fn(*head_args[0][:-1], st))
    print(f"head_fused_kernel (output layers + loss head + their gradients): {us:.1f} us")
print(f"sum of GEMMs: {total_us:.1f} us, {total_flop / 1e9:.1f} GFLOP, {total_flop / total_us / 1e6:.1f} TFLOP/s")
us = time_graph(lambda st: alg.minibatch_gradients(0))
print(f"minibatch_gradients (fwd + loss head + bwd) as one graph: {us:.1f} us -> {total_flop / us / 1e6:.1f} TFLOP/s")
us = time_graph(lambda st: (alg.minibatch_gradients(0), alg.optimizer_step(1)))
print(f"full optimizer step (+ grad norm + Adam) as one graph: {us:.1f} us")
