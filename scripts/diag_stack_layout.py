"""A/B of the frame-stack + finalize launch on dense rows (615 / 1050 floats) and on the 128-byte pitch (640 / 1056):
CUDA events around a graph of 4 launches on 4 cold buffer sets, L2 flushed before each replay."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isaac_b200 import _lib  # noqa: E402
from isaac_b200.envs.hector_config import HectorCfg  # noqa: E402
from isaac_b200.envs.hector_env import HectorFreeEnvB200  # noqa: E402
from isaac_b200.physics import SyntheticPhysics  # noqa: E402
from isaac_b200.synthetic import make_tape  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, nargs="+", default=[4096, 16384, 65536])
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _lib.load(check_device=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for n in args.envs:
        tape = make_tape(n, 2, seed=1)
        phys = SyntheticPhysics(n, device=dev)
        phys.load_frame(tape.physics[0].to(dev))
        env = HectorFreeEnvB200(HectorCfg(), sim_device=str(dev), physics=phys, statics=tape.statics)
        env.reset_buf.copy_(torch.rand(n, device=dev) < 0.005)
        for name, (ld_o, ld_p) in (("dense", (615, 1050)), ("pitched", (640, 1056))):
            env._p.obs_ld, env._p.priv_ld = (0, 0) if name == "dense" else (ld_o, ld_p)
            sets = [(torch.randn(n, ld_o, device=dev), torch.randn(n, ld_p, device=dev), torch.empty(n, ld_o, device=dev),
                     torch.empty(n, ld_p, device=dev)) for _ in range(4)]
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                st = torch.cuda.current_stream(dev).cuda_stream
                for s in sets:
                    _lib.check(lib.hb_env_stack_finalize(env._pp, env._pb, s[0].data_ptr(), s[1].data_ptr(), s[2].data_ptr(),
                                                         s[3].data_ptr(), env._host_count.data_ptr(), None, st), "stack")
            tot = 0.0
            stream = torch.cuda.current_stream(dev)
            for r in range(args.reps + 2):
                flush.fill_(r & 1)
                flush.sum()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                g.replay()
                b.record(stream)
                b.synchronize()
                if r >= 2:
                    tot += a.elapsed_time(b)
            us = tot / args.reps / 4 * 1e3
            gbs = n * (14 * 41 + 14 * 70) * 8 / (us * 1e-6) / 1e9
            print(f"envs {n:6d}  {name:8s} {us:8.2f} us/launch  {gbs:7.1f} GB/s algorithmic", flush=True)
            del sets, g


if __name__ == "__main__":
    main()
