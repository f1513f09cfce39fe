"""Rollout step anatomy (PPO.act -> env.step -> process_env_step, the loop of on_policy_runner.py:127-139): GPU time
between stream events against the host's wall clock, to tell launch bubbles from kernel time."""
import os
import sys
import time

import torch

sys.path.insert(0, os.getcwd())
from isaac_b200.algo import ActorCritic, PPO  # noqa: E402
from isaac_b200.envs.hector_config import HectorCfg  # noqa: E402
from isaac_b200.envs.hector_env import HectorFreeEnvB200  # noqa: E402
from isaac_b200.physics import SyntheticPhysics  # noqa: E402
from isaac_b200.synthetic import make_tape  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    T = 24
    stream = torch.cuda.current_stream(dev)
    tape = make_tape(n, 5, seed=1, fall_prob=0.005)
    frames = [f.to(dev) for f in tape.physics[1:]]
    phys = SyntheticPhysics(n, device=dev)
    phys.load_frame(tape.physics[0].to(dev))
    env = HectorFreeEnvB200(HectorCfg(), sim_device=str(dev), physics=phys, statics=tape.statics)
    env.enable_cuda_graph()
    ac = ActorCritic(615, 1050, 10, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128], device=dev)
    alg = PPO(ac, device=dev, num_learning_epochs=5, num_mini_batches=4, schedule="adaptive")
    alg.init_storage(n, T, [615], [1050], [10])
    alg.attach_env(env)
    obs, priv = env.get_observations(), env.get_privileged_observations()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for it in range(4):
        alg.storage.clear()
        marks = []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(T):
            e = [ev() for _ in range(4)]
            e[0].record(stream)
            a = alg.act(obs, priv)
            e[1].record(stream)
            obs, priv, rew, dones, infos = env.step(a)
            e[2].record(stream)
            alg.process_env_step(rew, dones, infos)
            e[3].record(stream)
            marks.append(e)
        host = time.perf_counter() - t0
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    act = sum(m[0].elapsed_time(m[1]) for m in marks) / T * 1e3
    step = sum(m[1].elapsed_time(m[2]) for m in marks) / T * 1e3
    rec = sum(m[2].elapsed_time(m[3]) for m in marks) / T * 1e3
    between = sum(marks[i][3].elapsed_time(marks[i + 1][0]) for i in range(T - 1)) / (T - 1) * 1e3
    print(f"envs {n}: wall {wall / T * 1e6:.1f} us/step, host enqueue {host / T * 1e6:.1f} us/step; on the stream: act {act:.1f} + "
          f"env.step {step:.1f} + record {rec:.1f} + between steps {between:.1f} us")


if __name__ == "__main__":
    main()
