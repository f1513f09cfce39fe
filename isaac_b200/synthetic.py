"""Synthetic Isaac Gym state tensors for the hector hot path.

Isaac Gym's PhysX stage (`gym.simulate` + `refresh_*_tensor`) is opaque to this
project, so tests and benchmarks feed the env stage with
synthetic state tensors laid out exactly like the gym ones
(reference: humanoid/envs/base/legged_robot.py:437-456):

    root_states    [N, 13]        pos(3) quat xyzw(4) linvel(3) angvel(3)
    dof_state      [N*ndof, 2]    (pos, vel) interleaved
    contact_forces [N, nbody, 3]
    rigid_state    [N, nbody, 13]

Everything is generated on the CPU with a seeded `torch.Generator`, so the
same tape can be replayed through the CUDA path, the oracle and (in the build
container only) the reference itself.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List

import torch

# hector dimensions (hector_config.py:8-20, robot.urdf)
NDOF = 10
NBODY = 11
FEET = (5, 10)
KNEES = (4, 9)
TERM_BODIES = (0, 3, 8)
DEFAULT_DOF_POS = (0.0, 0.0, 0.785, -1.578, 0.785, 0.0, 0.0, 0.785, -1.578, 0.785)


@dataclass(frozen=True)
class TaskDims:
    """What the generators need to know about a task's robot (default: hector)."""
    ndof: int = NDOF
    nbody: int = NBODY
    feet: tuple = FEET
    knees: tuple = KNEES
    term_bodies: tuple = TERM_BODIES
    default_dof_pos: tuple = DEFAULT_DOF_POS
    kp: tuple = (40.0, 40.0, 60.0, 120.0, 20.0) * 2
    kd: tuple = (3.0, 3.0, 5.0, 4.0, 1.0) * 2
    base_height: float = 0.55
    nobs: int = 41


def task_dims(cfg) -> TaskDims:
    """Dimensions, body indices and nominal gains of a task config (the way LeggedRobot._create_envs / _init_buffers
    resolve them by name: legged_robot.py:485-500,627-634,668-681)."""
    a = cfg.asset
    find = lambda pats: tuple(i for p in pats for i, nm in enumerate(a.body_names) if p in nm)
    kp, kd = [0.0] * len(a.dof_names), [0.0] * len(a.dof_names)
    for i, name in enumerate(a.dof_names):
        for key in cfg.control.stiffness:
            if key in name:
                kp[i], kd[i] = float(cfg.control.stiffness[key]), float(cfg.control.damping[key])
    return TaskDims(ndof=len(a.dof_names), nbody=len(a.body_names), feet=find([a.foot_name]), knees=find([a.knee_name]),
                    term_bodies=find(a.terminate_after_contacts_on),
                    default_dof_pos=tuple(cfg.init_state.default_joint_angles[k] for k in a.dof_names), kp=tuple(kp), kd=tuple(kd),
                    base_height=float(cfg.init_state.pos[2]), nobs=cfg.env.num_single_obs)


@dataclass
class PhysicsFrame:
    """One refresh of the gym tensors (what PhysX would hand back)."""
    root_states: torch.Tensor      # [N,13]
    dof_state: torch.Tensor        # [N*ndof,2]
    contact_forces: torch.Tensor   # [N,nbody,3]
    rigid_state: torch.Tensor      # [N,nbody,13]

    def to(self, device) -> "PhysicsFrame":
        return PhysicsFrame(*(t.to(device) for t in
                              (self.root_states, self.dof_state, self.contact_forces, self.rigid_state)))


@dataclass
class NoiseFrame:
    """Every random draw one `step()` can consume, indexed by env (SURVEY.md §8c:
    'all stochastic draws must be injected').  Uniforms are in [0,1)."""
    actions: torch.Tensor          # [N,10]  policy output fed to step()
    u_delay: torch.Tensor          # [N,1]   torch.rand for the action delay   (hector_env.py:166)
    z_action: torch.Tensor         # [N,10]  torch.randn_like(actions)         (hector_env.py:168)
    u_cmd: torch.Tensor            # [N,3]   command resample every 800 steps  (legged_robot.py:327-330)
    u_push: torch.Tensor           # [N,5]   push lin xy(2) + ang(3)           (hector_env.py:58-63)
    u_reset: torch.Tensor          # [N,15]  dof(10) + root xy(2) + cmd(3)     (legged_robot.py:366,384,327-330)
    z_obs: torch.Tensor            # [N,41]  torch.randn_like(obs_buf)         (hector_env.py:241)

    def to(self, device) -> "NoiseFrame":
        return NoiseFrame(*(t.to(device) for t in
                            (self.actions, self.u_delay, self.z_action, self.u_cmd, self.u_push,
                             self.u_reset, self.z_obs)))


def _quat_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack((aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw,
                        aw * bw - ax * bx - ay * by - az * bz), dim=-1)


def make_physics_frame(n: int, gen: torch.Generator, fall_prob: float = 0.005, dims: TaskDims = TaskDims()) -> PhysicsFrame:
    ndof, nbody, FEET, KNEES, TERM_BODIES = dims.ndof, dims.nbody, tuple(dims.feet), tuple(dims.knees), tuple(dims.term_bodies)
    f32 = dict(dtype=torch.float32, generator=gen)
    root = torch.zeros(n, 13)
    root[:, 0:2] = torch.rand(n, 2, **f32) * 2 - 1
    root[:, 2] = dims.base_height + 0.02 * torch.randn(n, **f32)
    tilt = torch.stack((0.1 * torch.randn(n, **f32), 0.1 * torch.randn(n, **f32),
                        0.3 * torch.randn(n, **f32), torch.ones(n)), dim=-1)
    tilt = tilt / tilt.norm(dim=-1, keepdim=True)
    # a quarter of the envs also get an arbitrary yaw so heading wrap-around is exercised
    yaw = (torch.rand(n, **f32) * 2 - 1) * math.pi
    yaw = torch.where(torch.rand(n, **f32) < 0.25, yaw, torch.zeros(n))
    qyaw = torch.stack((torch.zeros(n), torch.zeros(n), torch.sin(yaw / 2), torch.cos(yaw / 2)), dim=-1)
    q = _quat_mul(qyaw, tilt)
    root[:, 3:7] = q / q.norm(dim=-1, keepdim=True)
    root[:, 7:10] = 0.3 * torch.randn(n, 3, **f32)
    root[:, 10:13] = 0.5 * torch.randn(n, 3, **f32)

    q0 = torch.tensor(dims.default_dof_pos)
    dof = torch.zeros(n, ndof, 2)
    dof[..., 0] = q0 + 0.2 * torch.randn(n, ndof, **f32)
    dof[..., 1] = 1.5 * torch.randn(n, ndof, **f32)
    # some envs sit close to the default pose so the exp(-100*x) branch of
    # default_joint_pos is exercised away from zero
    near = torch.rand(n, **f32) < 0.3
    dof[near, :, 0] = q0 + 0.02 * torch.randn(int(near.sum()), ndof, **f32)

    contact = torch.zeros(n, nbody, 3)
    for f in FEET:
        on = (torch.rand(n, **f32) < 0.6).float()
        contact[:, f, 2] = torch.rand(n, **f32) * 250 * on
        contact[:, f, 0:2] = 10 * torch.randn(n, 2, **f32) * on[:, None]
    for b in TERM_BODIES:
        fall = torch.rand(n, **f32) < fall_prob
        d = torch.randn(n, 3, **f32)
        contact[:, b, :] = torch.where(fall[:, None], 50 * d / d.norm(dim=-1, keepdim=True), contact[:, b, :])
        graze = (torch.rand(n, **f32) < 0.02) & ~fall
        contact[:, b, :] = torch.where(graze[:, None], 0.3 * d / d.norm(dim=-1, keepdim=True), contact[:, b, :])

    rigid = 0.3 * torch.randn(n, nbody, 13, **f32)
    for k, body in enumerate(FEET + KNEES):
        side = 0.1 if k % 2 == 0 else -0.1
        spread = 1.0 + 2.5 * torch.rand(n, **f32)          # feet/knee distance 0.2 .. 0.7 m
        rigid[:, body, 0] = root[:, 0] + 0.05 * torch.randn(n, **f32)
        rigid[:, body, 1] = root[:, 1] + side * spread
        rigid[:, body, 2] = torch.rand(n, **f32) * 0.12 + (0.0 if body in FEET else 0.25)
        rigid[:, body, 7:10] = 0.5 * torch.randn(n, 3, **f32)
    return PhysicsFrame(root, dof.reshape(n * ndof, 2).contiguous(), contact, rigid)


def make_noise_frame(n: int, gen: torch.Generator, dims: TaskDims = TaskDims()) -> NoiseFrame:
    ndof, nobs = dims.ndof, dims.nobs
    f32 = dict(dtype=torch.float32, generator=gen)
    return NoiseFrame(actions=torch.randn(n, ndof, **f32),
                      u_delay=torch.rand(n, 1, **f32),
                      z_action=torch.randn(n, ndof, **f32),
                      u_cmd=torch.rand(n, 3, **f32),
                      u_push=torch.rand(n, 5, **f32),
                      u_reset=torch.rand(n, ndof + 5, **f32),
                      z_obs=torch.randn(n, nobs, **f32))


@dataclass
class EnvStatics:
    """Per-env constants fixed at env creation (legged_robot.py:256-301,465-500,683-709)."""
    p_gains: torch.Tensor       # [N,10]
    d_gains: torch.Tensor       # [N,10]
    env_frictions: torch.Tensor # [N,1]
    body_mass: torch.Tensor     # [N,1]
    env_origins: torch.Tensor   # [N,3]
    episode_length0: torch.Tensor  # [N] int64, init_at_random_ep_len (on_policy_runner.py:103-106)


KP_NOMINAL = (40.0, 40.0, 60.0, 120.0, 20.0) * 2      # hector_config.py:93-94 via legged_robot.py:485-500
KD_NOMINAL = (3.0, 3.0, 5.0, 4.0, 1.0) * 2            # hector_config.py:95-96


def make_env_statics(n: int, gen: torch.Generator, randomize_gains: bool = False,
                     max_episode_length: int = 2400, dims: TaskDims = TaskDims()) -> EnvStatics:
    f32 = dict(dtype=torch.float32, generator=gen)
    kp = torch.tensor(dims.kp).repeat(n, 1)
    kd = torch.tensor(dims.kd).repeat(n, 1)
    if randomize_gains:   # BASELINE.json config 3: kp/kd domain randomisation
        kp = kp * (0.8 + 0.4 * torch.rand(n, dims.ndof, **f32))
        kd = kd * (0.8 + 0.4 * torch.rand(n, dims.ndof, **f32))
    buckets = 0.1 + 0.9 * torch.rand(256, 1, **f32)               # legged_robot.py:256-268
    fric = buckets[torch.randint(0, 256, (n,), generator=gen)]
    mass = 13.0 + (torch.rand(n, 1, **f32) * 6 - 2)               # legged_robot.py:295-301
    origins = torch.zeros(n, 3)
    origins[:, 0:2] = torch.randint(0, 20, (n, 2), generator=gen).float() * 8.0
    origins[:, 2] = 0.1 * torch.rand(n, **f32)
    ep0 = torch.randint(0, max_episode_length, (n,), generator=gen, dtype=torch.int64)
    return EnvStatics(kp, kd, fric, mass, origins, ep0)


@dataclass
class Tape:
    """T steps of physics frames and noise for N envs."""
    statics: EnvStatics
    physics: List[PhysicsFrame] = field(default_factory=list)
    noise: List[NoiseFrame] = field(default_factory=list)


def make_tape(n: int, steps: int, seed: int = 1234, randomize_gains: bool = False,
              fall_prob: float = 0.005, cfg=None) -> Tape:
    """`cfg`: a task config (isaac_b200.envs: HectorCfg / HectorFullCfg / XBotLCfg); default = hector."""
    dims = task_dims(cfg) if cfg is not None else TaskDims()
    gen = torch.Generator().manual_seed(seed)
    tape = Tape(make_env_statics(n, gen, randomize_gains, dims=dims))
    for _ in range(steps):
        tape.physics.append(make_physics_frame(n, gen, fall_prob, dims))
        tape.noise.append(make_noise_frame(n, gen, dims))
    return tape
