"""`OnPolicyRunner` with the reference's interface (algo/ppo/on_policy_runner.py:45-307) over the B200 classes.

The reference's own runner also runs unmodified over `isaac_b200.algo.PPO` / `ActorCritic` and `HectorFreeEnvB200`
(tests/test_runner.py does exactly that); this one exists for the part of its loop that is on the hot path:

  * per-step bookkeeping (on_policy_runner.py:140-154) costs the reference two `nonzero()` + `.cpu()` round trips per env
    step.  Here it is one launch (`hb_runner_bookkeeping`): the running sums and the two `deque(maxlen=100)` buffers live
    on the device and are read once per iteration, by `log()`;
  * `PPO.attach_env`: the env writes its observations straight into the rollout slots.

Same constructor (`env, train_cfg, log_dir, device`), `learn`, `log`, `save`, `load`, `get_inference_policy`,
`get_inference_critic`, same checkpoint layout (`model_state_dict`, `optimizer_state_dict`, `iter`, `infos`) and the same
scalar names.  TensorBoard is used when it is importable and a log_dir is given; W&B is not contacted.
"""
from __future__ import annotations

import os
import statistics
import time
from typing import Optional

import torch

from .. import _lib
from .actor_critic import ActorCritic
from .ppo import PPO

_CLASSES = {"ActorCritic": ActorCritic, "PPO": PPO}
DEQUE_LEN = 100          # deque(maxlen=100), on_policy_runner.py:117-118


class EpisodeBuffers:
    """cur_reward_sum / cur_episode_length and the reward / length deques of the reference's loop, device-resident."""

    def __init__(self, num_envs: int, device, capacity: int = DEQUE_LEN):
        self._lib = _lib.load()
        self.device, self.capacity, self.num_envs = torch.device(device), capacity, num_envs
        z = lambda *s, **kw: torch.zeros(*s, device=self.device, **kw)
        self.cur_reward_sum, self.cur_episode_length = z(num_envs), z(num_envs)
        self.ring_rewards, self.ring_lengths = z(capacity), z(capacity)
        self.ring_state = z(1, dtype=torch.int64)

    def step(self, rewards: torch.Tensor, dones: torch.Tensor) -> None:
        rewards = rewards.to(self.device, dtype=torch.float32).contiguous()
        dones = dones.to(self.device).contiguous()
        if dones.dtype not in (torch.bool, torch.uint8):
            dones = dones > 0
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.hb_runner_bookkeeping(rewards.data_ptr(), dones.data_ptr(), self.num_envs,
                                                   self.cur_reward_sum.data_ptr(), self.cur_episode_length.data_ptr(),
                                                   self.ring_rewards.data_ptr(), self.ring_lengths.data_ptr(), self.capacity,
                                                   self.ring_state.data_ptr(), st), "hb_runner_bookkeeping")

    def deques(self):
        """(rewbuffer, lenbuffer) as lists, oldest first - what the reference's deques hold.  One host sync."""
        total = int(self.ring_state.item())
        count = min(total, self.capacity)
        order = [(total - count + j) % self.capacity for j in range(count)]
        rew, ln = self.ring_rewards.cpu(), self.ring_lengths.cpu()
        return [float(rew[k]) for k in order], [float(ln[k]) for k in order]


class OnPolicyRunner:
    def __init__(self, env, train_cfg, log_dir=None, device="cuda:0"):
        self.cfg, self.alg_cfg, self.policy_cfg = train_cfg["runner"], train_cfg["algorithm"], train_cfg["policy"]
        self.all_cfg = train_cfg
        self.device, self.env = device, env
        num_critic_obs = env.num_privileged_obs if env.num_privileged_obs is not None else env.num_obs
        policy_cls = _CLASSES[self.cfg.get("policy_class_name", "ActorCritic")]
        alg_cls = _CLASSES[self.cfg.get("algorithm_class_name", "PPO")]
        actor_critic = policy_cls(env.num_obs, num_critic_obs, env.num_actions, device=device, **self.policy_cfg).to(device)
        self.alg = alg_cls(actor_critic, device=device, **self.alg_cfg)
        self.num_steps_per_env, self.save_interval = self.cfg["num_steps_per_env"], self.cfg["save_interval"]
        self.alg.init_storage(env.num_envs, self.num_steps_per_env, [env.num_obs], [env.num_privileged_obs], [env.num_actions])
        if hasattr(env, "set_next_observation_buffers") and env.num_privileged_obs is not None:
            self.alg.attach_env(env)          # observations are written where the update reads them
        self.log_dir, self.writer = log_dir, None
        self.tot_timesteps, self.tot_time, self.current_learning_iteration = 0, 0.0, 0
        self.episodes = EpisodeBuffers(env.num_envs, device)
        self.last_log: Optional[dict] = None
        env.reset()

    # ------------------------------------------------------------------ on_policy_runner.py:91-177
    def learn(self, num_learning_iterations, init_at_random_ep_len=False):
        if self.log_dir is not None and self.writer is None:
            os.makedirs(self.log_dir, exist_ok=True)
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.writer = SummaryWriter(log_dir=self.log_dir, flush_secs=10)
            except Exception:          # no tensorboard in this environment: scalars stay in self.last_log
                self.writer = None
        env, alg = self.env, self.alg
        if init_at_random_ep_len:
            env.episode_length_buf = torch.randint_like(env.episode_length_buf, high=int(env.max_episode_length))
        obs = env.get_observations()
        priv = env.get_privileged_observations()
        critic_obs = priv if priv is not None else obs
        alg.actor_critic.train()
        ep_infos = []
        last = self.current_learning_iteration + num_learning_iterations
        for it in range(self.current_learning_iteration, last):
            t0 = time.time()
            for _ in range(self.num_steps_per_env):
                actions = alg.act(obs, critic_obs)
                obs, priv, rewards, dones, infos = env.step(actions)
                critic_obs = priv if priv is not None else obs
                alg.process_env_step(rewards, dones, infos)
                if self.log_dir is not None:
                    if "episode" in infos:
                        ep_infos.append(infos["episode"])
                    self.episodes.step(rewards, dones)
            torch.cuda.synchronize(self.device)
            t1 = time.time()
            alg.compute_returns(critic_obs)
            mean_value_loss, mean_surrogate_loss = alg.update()
            t2 = time.time()
            if self.log_dir is not None:
                self.log(dict(it=it, num_learning_iterations=num_learning_iterations, collection_time=t1 - t0, learn_time=t2 - t1,
                              mean_value_loss=mean_value_loss, mean_surrogate_loss=mean_surrogate_loss, ep_infos=ep_infos))
                if it % self.save_interval == 0:
                    self.save(os.path.join(self.log_dir, f"model_{it}.pt"))
            ep_infos.clear()
        self.current_learning_iteration += num_learning_iterations
        if self.log_dir is not None:
            self.save(os.path.join(self.log_dir, f"model_{self.current_learning_iteration}.pt"))

    # ------------------------------------------------------------------ on_policy_runner.py:179-273
    def log(self, locs, width=80, pad=35):
        steps = self.num_steps_per_env * self.env.num_envs
        iteration_time = locs["collection_time"] + locs["learn_time"]
        self.tot_timesteps += steps
        self.tot_time += iteration_time
        scalars, lines = {}, []
        if locs["ep_infos"]:          # extras["episode"]: one 0-dim device tensor per reward term and step
            for key in locs["ep_infos"][0]:
                vals = torch.stack([torch.as_tensor(info[key], dtype=torch.float32, device=self.device).reshape(())
                                    for info in locs["ep_infos"]])
                scalars["Episode/" + key] = float(vals.mean())
                lines.append(f"{f'Mean episode {key}:':>{pad}} {scalars['Episode/' + key]:.4f}")
        mean_std = float(self.alg.actor_critic.std.mean())
        fps = int(steps / iteration_time)
        scalars.update({"Loss/value_function": locs["mean_value_loss"], "Loss/surrogate": locs["mean_surrogate_loss"],
                        "Loss/learning_rate": self.alg.learning_rate, "Policy/mean_noise_std": mean_std, "Perf/total_fps": fps,
                        "Perf/collection time": locs["collection_time"], "Perf/learning_time": locs["learn_time"]})
        rewbuffer, lenbuffer = self.episodes.deques()
        if rewbuffer:
            scalars["Train/mean_reward"] = statistics.mean(rewbuffer)
            scalars["Train/mean_episode_length"] = statistics.mean(lenbuffer)
        if self.writer is not None:
            for k, v in scalars.items():
                self.writer.add_scalar(k, v, locs["it"])
            if rewbuffer:
                self.writer.add_scalar("Train/mean_reward/time", scalars["Train/mean_reward"], self.tot_time)
                self.writer.add_scalar("Train/mean_episode_length/time", scalars["Train/mean_episode_length"], self.tot_time)
        head = f" \033[1m Learning iteration {locs['it']}/{self.current_learning_iteration + locs['num_learning_iterations']} \033[0m "
        body = [f"{'Computation:':>{pad}} {fps:.0f} steps/s (collection: {locs['collection_time']:.3f}s, learning {locs['learn_time']:.3f}s)",
                f"{'Value function loss:':>{pad}} {locs['mean_value_loss']:.4f}", f"{'Surrogate loss:':>{pad}} {locs['mean_surrogate_loss']:.4f}",
                f"{'Mean action noise std:':>{pad}} {mean_std:.2f}"]
        if rewbuffer:
            body += [f"{'Mean reward:':>{pad}} {scalars['Train/mean_reward']:.2f}",
                     f"{'Mean episode length:':>{pad}} {scalars['Train/mean_episode_length']:.2f}"]
        eta = self.tot_time / (locs["it"] + 1) * (locs["num_learning_iterations"] - locs["it"])
        tail = [f"{'Total timesteps:':>{pad}} {self.tot_timesteps}", f"{'Iteration time:':>{pad}} {iteration_time:.2f}s",
                f"{'Total time:':>{pad}} {self.tot_time:.2f}s", f"{'ETA:':>{pad}} {eta:.1f}s"]
        print("\n".join(["#" * width, head.center(width, " "), ""] + body + lines + ["-" * width] + tail))
        self.last_log = scalars

    # ------------------------------------------------------------------ on_policy_runner.py:275-307
    def save(self, path, infos=None):
        torch.save({"model_state_dict": self.alg.actor_critic.state_dict(), "optimizer_state_dict": self.alg.optimizer.state_dict(),
                    "iter": self.current_learning_iteration, "infos": infos}, path)

    def load(self, path, load_optimizer=True):
        loaded = torch.load(path, map_location="cpu")
        self.alg.actor_critic.load_state_dict(loaded["model_state_dict"])
        if load_optimizer:
            self.alg.optimizer.load_state_dict(loaded["optimizer_state_dict"])
        self.current_learning_iteration = loaded["iter"]
        return loaded["infos"]

    def get_inference_policy(self, device=None):
        self.alg.actor_critic.eval()
        if device is not None:
            self.alg.actor_critic.to(device)
        return self.alg.actor_critic.act_inference

    def get_inference_critic(self, device=None):
        self.alg.actor_critic.eval()
        if device is not None:
            self.alg.actor_critic.to(device)
        return self.alg.actor_critic.evaluate
