"""One pass over every hot kernel at the benchmarked size (BASELINE configs[1]: 4096 envs, T = 24, minibatch 24576) inside a
cudaProfilerStart / Stop range, for `ncu --profile-from-start off --set full`:
    one env step launched eagerly (prologue + PD, 9 x PD, post-physics, frame stack + finalisation),
    compute_returns (one GAE launch), one minibatch step of PPO.update (3 grouped forward GEMMs, fused head, 5 grouped
    backward GEMMs, optimizer step), one PPO.act.
Usage: python scripts/profile_r02.py [envs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from isaac_b200.algo import ActorCritic, PPO
from isaac_b200.envs.hector_config import HectorCfg
from isaac_b200.envs.hector_env import HectorFreeEnvB200
from isaac_b200.physics import SyntheticPhysics
from isaac_b200.synthetic import make_tape

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = 24
dev = torch.device("cuda:0")
tape = make_tape(n, 3, seed=1234, fall_prob=0.005, randomize_gains=True)
phys = SyntheticPhysics(n, device=dev)
phys.load_frame(tape.physics[0].to(dev))
env = HectorFreeEnvB200(HectorCfg(), sim_device="cuda:0", physics=phys, statics=tape.statics, initial_noise=tape.noise[0].to(dev))
env.episode_length_buf.copy_(tape.statics.episode_length0)
torch.manual_seed(5)
ac = ActorCritic(615, 1050, 10, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128], device=dev)
alg = PPO(ac, device=dev, num_learning_epochs=1, num_mini_batches=4, clip_param=0.2, gamma=0.994, lam=0.9, value_loss_coef=1.0,
          entropy_coef=0.001, learning_rate=1e-5, max_grad_norm=1.0, schedule="adaptive", desired_kl=0.01)
alg.graph_rollout = alg.graph_update = False
alg.init_storage(n, T, [615], [1050], [10])
alg.attach_env(env)
frames = [f.to(dev) for f in tape.physics]
actions = tape.noise[1].actions.to(dev)


def rollout():
    obs, priv = env.get_observations(), env.get_privileged_observations()
    alg.storage.clear()
    for t in range(T):
        a = alg.act(obs, priv)
        phys.load_frame(frames[t % 3])
        obs, priv, rew, dn, infos = env.step(a)
        alg.process_env_step(rew, dn, infos)
    return priv


priv = rollout()
alg.compute_returns(priv)
alg.update()
priv = rollout()
torch.cuda.synchronize()
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
flush.fill_(1.0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
phys.load_frame(frames[1])
env.step(actions)                                   # eager: 12 launches
alg.compute_returns(priv)                           # critic forward (4 GEMMs) + hb_gae_fused
alg.prepare_minibatches()                           # gathers + pack
alg._per_update.zero_()
alg.minibatch_gradients(0)                          # 3 grouped fwd, head, 5 grouped bwd
alg.optimizer_step(1)
alg.storage.clear()
alg.act(env.get_observations(), env.get_privileged_observations())
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled pass done")
