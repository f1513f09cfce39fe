"""The opaque physics stage.

Isaac Gym's PhysX `gym.simulate` stays outside this project (BASELINE.json north_star).
The env talks to it through the handful of calls the reference makes on its hot path
(envs/base/legged_robot.py:93-100,123-125,370-372,394-396; hector_env.py:67-68):

    set_dof_actuation_force(torques) / simulate() / refresh_dof_state()
    refresh_post_physics()                  root, net contact force, rigid body tensors
    set_dof_state_indexed(ids_int32, n)     after reset wrote dof_state for `ids`
    set_root_state_indexed(ids_int32, n)    after reset wrote root_states for `ids`
    set_root_state()                        after a push wrote root_states

`SyntheticPhysics` owns torch tensors with the gym layouts and plays back supplied
frames (or leaves the state untouched), which is how tests and `bench.py` run the env
stage without the closed simulator.  `IsaacGymPhysics` is the adapter for the real thing
(needs `isaacgym`, which is not in this image).
"""
from __future__ import annotations

from typing import Optional

import torch


class SyntheticPhysics:
    """Gym-layout state tensors on `device`; `simulate()` is a no-op."""

    capturable = True      # no host work between decimation sub-steps: the env step may be replayed as a CUDA graph

    def __init__(self, num_envs: int, num_dof: int = 10, num_bodies: int = 11, device="cuda:0"):
        self.num_envs, self.num_dof, self.num_bodies = num_envs, num_dof, num_bodies
        self.device = torch.device(device)
        f32 = dict(dtype=torch.float32, device=self.device)
        self.root_states = torch.zeros(num_envs, 13, **f32)
        self.root_states[:, 6] = 1.0
        self.dof_state = torch.zeros(num_envs * num_dof, 2, **f32)
        self.contact_forces = torch.zeros(num_envs, num_bodies, 3, **f32)
        self.rigid_state = torch.zeros(num_envs, num_bodies, 13, **f32)
        self.calls = {"simulate": 0, "set_dof_state_indexed": 0, "set_root_state_indexed": 0, "set_root_state": 0}
        self.last_indexed_ids: Optional[torch.Tensor] = None

    def load_frame(self, frame) -> None:
        """Install what PhysX would have refreshed (isaac_b200.synthetic.PhysicsFrame)."""
        self.root_states.copy_(frame.root_states, non_blocking=True)
        self.dof_state.copy_(frame.dof_state, non_blocking=True)
        self.contact_forces.copy_(frame.contact_forces, non_blocking=True)
        self.rigid_state.copy_(frame.rigid_state, non_blocking=True)

    # --- the calls of the reference's hot path -------------------------------------------
    def set_dof_actuation_force(self, torques: torch.Tensor) -> None:
        pass

    def simulate(self) -> None:
        self.calls["simulate"] += 1

    def refresh_dof_state(self) -> None:
        pass

    def refresh_post_physics(self) -> None:
        pass

    def set_dof_state_indexed(self, env_ids_int32: torch.Tensor, count: int) -> None:
        self.calls["set_dof_state_indexed"] += 1
        self.last_indexed_ids = env_ids_int32[:count]

    def set_root_state_indexed(self, env_ids_int32: torch.Tensor, count: int) -> None:
        self.calls["set_root_state_indexed"] += 1

    def set_root_state(self) -> None:
        self.calls["set_root_state"] += 1


class IsaacGymPhysics:
    """Adapter over a live Isaac Gym sim (same calls as the reference makes)."""

    def __init__(self, gym, sim, num_envs: int, device: str):
        from isaacgym import gymtorch  # noqa: F401  (closed third-party package)
        self.gym, self.sim, self._gt = gym, sim, gymtorch
        self.device = torch.device(device)
        self.num_envs = num_envs
        self.root_states = gymtorch.wrap_tensor(gym.acquire_actor_root_state_tensor(sim))
        self.dof_state = gymtorch.wrap_tensor(gym.acquire_dof_state_tensor(sim))
        self.contact_forces = gymtorch.wrap_tensor(gym.acquire_net_contact_force_tensor(sim)).view(num_envs, -1, 3)
        self.rigid_state = gymtorch.wrap_tensor(gym.acquire_rigid_body_state_tensor(sim)).view(num_envs, -1, 13)
        self.num_bodies = self.rigid_state.shape[1]
        self.num_dof = self.dof_state.shape[0] // num_envs

    def set_dof_actuation_force(self, torques):
        self.gym.set_dof_actuation_force_tensor(self.sim, self._gt.unwrap_tensor(torques))

    def simulate(self):
        self.gym.simulate(self.sim)
        if self.device.type == "cpu":
            self.gym.fetch_results(self.sim, True)

    def refresh_dof_state(self):
        self.gym.refresh_dof_state_tensor(self.sim)

    def refresh_post_physics(self):
        self.gym.refresh_actor_root_state_tensor(self.sim)
        self.gym.refresh_net_contact_force_tensor(self.sim)
        self.gym.refresh_rigid_body_state_tensor(self.sim)

    def set_dof_state_indexed(self, env_ids_int32, count):
        self.gym.set_dof_state_tensor_indexed(self.sim, self._gt.unwrap_tensor(self.dof_state),
                                              self._gt.unwrap_tensor(env_ids_int32), count)

    def set_root_state_indexed(self, env_ids_int32, count):
        self.gym.set_actor_root_state_tensor_indexed(self.sim, self._gt.unwrap_tensor(self.root_states),
                                                     self._gt.unwrap_tensor(env_ids_int32), count)

    def set_root_state(self):
        self.gym.set_actor_root_state_tensor(self.sim, self._gt.unwrap_tensor(self.root_states))
