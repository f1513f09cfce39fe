"""Where a replayed step spends its time: events around graph A (prologue + first PD) and graph B (the rest), with the
bench.py protocol (L2 flushed between steps)."""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from isaac_b200.envs.hector_config import HectorCfg  # noqa: E402
from isaac_b200.envs.hector_env import HectorFreeEnvB200  # noqa: E402
from isaac_b200.physics import SyntheticPhysics  # noqa: E402
from isaac_b200.synthetic import make_tape  # noqa: E402


class Probe:
    def __init__(self, graph, log, tag, stream):
        self.graph, self.log, self.tag, self.stream = graph, log, tag, stream

    def replay(self, st):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(self.stream)
        self.graph.replay(st)
        b.record(self.stream)
        self.log.append((self.tag, a, b))


def main():
    dev = torch.device("cuda:0")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)
    tape = make_tape(n, 5, seed=1, fall_prob=0.005)
    frames = [f.to(dev) for f in tape.physics[1:]]
    acts = [f.actions.to(dev) for f in tape.noise[1:]]
    phys = SyntheticPhysics(n, device=dev)
    phys.load_frame(tape.physics[0].to(dev))
    env = HectorFreeEnvB200(HectorCfg(), sim_device=str(dev), physics=phys, statics=tape.statics)
    for i in range(3):
        phys.load_frame(frames[i % 4])
        env.step(acts[i % 4])
    env.enable_cuda_graph()
    for i in range(6):          # lazily captured graphs exist after this
        phys.load_frame(frames[i % 4])
        env.step(acts[i % 4])
    torch.cuda.synchronize()
    log = []
    env._graphs_a = {k: Probe(g, log, "A", stream) for k, g in env._graphs_a.items()}
    env._graphs = {k: (Probe(v[0], log, "B", stream),) + tuple(v[1:]) for k, v in env._graphs.items()}
    steps = []
    for i in range(60):
        phys.load_frame(frames[i % 4])
        flush.fill_(float(i))
        flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        env.step(acts[i % 4])
        b.record(stream)
        steps.append((a, b))
    torch.cuda.synchronize()
    tot = sum(a.elapsed_time(b) for a, b in steps[10:]) / 50 * 1e3
    ta = [a.elapsed_time(b) for t, a, b in log if t == "A"][10:]
    tb = [a.elapsed_time(b) for t, a, b in log if t == "B"][10:]
    gap = [log[i][2].elapsed_time(log[i + 1][1]) for i in range(0, len(log) - 1, 2)][10:]
    pre = [steps[i][0].elapsed_time(log[2 * i][1]) for i in range(10, 60)]
    post = [log[2 * i + 1][2].elapsed_time(steps[i][1]) for i in range(10, 60)]
    m = lambda v: sum(v) / len(v) * 1e3
    print(f"{os.getcwd()}: step {tot:.2f} us = before A {m(pre):.2f} + A {m(ta):.2f} + gap {m(gap):.2f} + B {m(tb):.2f} + after B {m(post):.2f}")


if __name__ == "__main__":
    main()
