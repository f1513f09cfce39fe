"""A/B of the replayed env step with dense observation rows (615 / 1050 floats) and with the 16-byte pitch the env
uses by default (616 / 1052): per-step CUDA events, L2 flushed between steps (the bench.py protocol)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isaac_b200.envs.hector_config import HectorCfg  # noqa: E402
from isaac_b200.envs.hector_env import HectorFreeEnvB200  # noqa: E402
from isaac_b200.physics import SyntheticPhysics  # noqa: E402
from isaac_b200.synthetic import make_tape  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, nargs="+", default=[4096, 16384])
    ap.add_argument("--steps", type=int, default=60)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    for n in args.envs:
        tape = make_tape(n, 5, seed=1, fall_prob=0.005)
        frames = [f.to(dev) for f in tape.physics[1:]]
        acts = [f.actions.to(dev) for f in tape.noise[1:]]
        for rnd in range(2):
            for dense in (True, False):
                phys = SyntheticPhysics(n, device=dev)
                phys.load_frame(tape.physics[0].to(dev))
                env = HectorFreeEnvB200(HectorCfg(), sim_device=str(dev), physics=phys, statics=tape.statics, dense_rows=dense)
                env.enable_cuda_graph()
                tot = 0.0
                for i in range(args.steps + 6):
                    phys.load_frame(frames[i % 4])
                    flush.fill_(i & 1)
                    flush.sum()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream)
                    env.step(acts[i % 4])
                    b.record(stream)
                    b.synchronize()
                    if i >= 6:
                        tot += a.elapsed_time(b)
                print(f"envs {n:6d} round {rnd} {'dense  ' if dense else 'pitched'} {tot / args.steps * 1e3:8.2f} us/step", flush=True)


if __name__ == "__main__":
    main()
