"""A/B of the single-launch GAE (hb_gae_fused): block widths.  A graph of 8 launches on 8
cold buffer sets, / 8 (bench.py's protocol), best of 5.   python scripts/ab_gae.py [envs ...]"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch

from isaac_b200 import _lib
from isaac_b200.algo.rollout_storage import gae_compute_returns

dev = torch.device("cuda:0")
lib = _lib.load(check_device=True)
T = 24
sizes = [int(a) for a in sys.argv[1:]] or [4096, 16384, 65536]
stream = torch.cuda.Stream(dev)
flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
for n in sizes:
    g = torch.Generator().manual_seed(n)
    sets = []
    for _ in range(8):
        r_, v_ = torch.rand(T, n, 1, generator=g).to(dev), torch.randn(T, n, 1, generator=g).to(dev)
        d_, lv_ = (torch.rand(T, n, 1, generator=g) < 0.005).byte().to(dev), torch.randn(n, 1, generator=g).to(dev)
        sets.append((r_, v_, d_, lv_, torch.empty_like(r_), torch.empty_like(r_)))
    ref = None
    for staged in (0,):
        for threads in (0, 32, 64, 128, 256):
            lib.hb_set_option(b"gae_threads", threads)
            with torch.cuda.stream(stream):
                gae_compute_returns(*sets[0], 0.994, 0.9)
                stream.synchronize()
                out = (sets[0][4].clone(), sets[0][5].clone())
                if ref is None:
                    ref = out
                same_ret = torch.equal(out[0], ref[0])
                adv_err = float((out[1] - ref[1]).abs().max())
            graph = _lib.LaunchGraph(dev).record(lambda st: [gae_compute_returns(*s_, 0.994, 0.9) for s_ in sets])
            best = 1e9
            for rep in range(5):
                flush.zero_()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    a.record(stream)
                    graph.replay(stream.cuda_stream)
                    b.record(stream)
                b.synchronize()
                best = min(best, a.elapsed_time(b) / 8)
            print(f"envs {n} threads={threads or 'auto'}: {best * 1e3:.2f} us  returns bit-equal {same_ret}  adv max diff {adv_err:.2e}", flush=True)
lib.hb_set_option(b"gae_threads", 0)
