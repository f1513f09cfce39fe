"""Hector task configuration for the B200 hot path.

Field names and nesting follow the reference's config objects so that either this
module's `HectorCfg()` or the reference's own instance can be handed to
`HectorFreeEnvB200` (it only reads attributes).  Values restate
humanoid/envs/custom/hector_config.py:4-234 and the base defaults it inherits from
humanoid/envs/base/legged_robot_config.py:34-237.  Only what the hot path reads is kept
(no asset/terrain-generation/viewer/PhysX solver settings: SURVEY.md §8 out of scope).
"""
from __future__ import annotations

import inspect


class _Cfg:
    """Instantiate nested classes recursively (same behaviour as
    humanoid/envs/base/base_config.py:34-55) so `cfg.env.num_envs` is an instance attribute."""

    def __init__(self):
        _Cfg._instantiate(self)

    @staticmethod
    def _instantiate(obj):
        for key in dir(obj):
            if key.startswith("__"):
                continue
            val = getattr(obj, key)
            if inspect.isclass(val):
                inst = val()
                setattr(obj, key, inst)
                _Cfg._instantiate(inst)


class HectorCfg(_Cfg):
    class env:
        frame_stack = 15
        c_frame_stack = 15
        num_single_obs = 41
        num_observations = 15 * 41
        single_num_privileged_obs = 70
        num_privileged_obs = 15 * 70
        num_actions = 10
        num_envs = 4096
        episode_length_s = 24
        use_ref_actions = False
        send_timeouts = True            # legged_robot_config.py:41
        env_spacing = 3.0

    class safety:
        pos_limit = 0.8
        vel_limit = 0.5
        torque_limit = 0.85

    class asset:
        name = "hector"
        foot_name = "toe"
        knee_name = "calf"
        terminate_after_contacts_on = ["base", "thigh"]
        penalize_contacts_on = ["base", "thigh"]
        fix_base_link = False
        # What Isaac Gym reports for resources/robots/hector_v2/xacro/robot.urdf after
        # collapse_fixed_joints: order of bodies/DOFs and URDF effort limits.
        body_names = ["base", "L_hip", "L_hip2", "L_thigh", "L_calf", "L_toe",
                      "R_hip", "R_hip2", "R_thigh", "R_calf", "R_toe"]
        dof_names = ["L_hip_joint", "L_hip_roll_joint", "L_thigh_joint", "L_calf_joint", "L_toe_joint",
                     "R_hip_joint", "R_hip_roll_joint", "R_thigh_joint", "R_calf_joint", "R_toe_joint"]
        dof_effort = [33.5, 33.5, 33.5, 67.0, 33.5, 33.5, 33.5, 33.5, 67.0, 33.5]

    class terrain:
        mesh_type = "trimesh"
        curriculum = False
        measure_heights = False
        # height sampling (legged_robot_config.py:46-48,55-56; used when measure_heights is switched on)
        horizontal_scale = 0.1
        vertical_scale = 0.005
        border_size = 25
        measured_points_x = [-0.8, -0.7, -0.6, -0.5, -0.4, -0.3, -0.2, -0.1, 0., 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8]
        measured_points_y = [-0.5, -0.4, -0.3, -0.2, -0.1, 0., 0.1, 0.2, 0.3, 0.4, 0.5]

    class noise:
        add_noise = True
        noise_level = 0.6

        class noise_scales:
            dof_pos = 0.05
            dof_vel = 0.5
            ang_vel = 0.1
            lin_vel = 0.05
            quat = 0.03
            height_measurements = 0.1

    class init_state:
        pos = [0.0, 0.0, 0.55]
        rot = [0.0, 0.0, 0.0, 1.0]
        lin_vel = [0.0, 0.0, 0.0]
        ang_vel = [0.0, 0.0, 0.0]
        default_joint_angles = {
            "L_hip_joint": 0.0, "L_hip_roll_joint": 0.0, "L_thigh_joint": 0.785,
            "L_calf_joint": -1.578, "L_toe_joint": 0.785,
            "R_hip_joint": 0.0, "R_hip_roll_joint": 0.0, "R_thigh_joint": 0.785,
            "R_calf_joint": -1.578, "R_toe_joint": 0.785,
        }

    class control:
        stiffness = {"hip_joint": 40.0, "hip_roll": 40.0, "thigh": 60.0, "calf": 120.0, "toe": 20.0}
        damping = {"hip_joint": 3.0, "hip_roll": 3.0, "thigh": 5.0, "calf": 4.0, "toe": 1.0}
        action_scale = 0.25
        decimation = 10

    class sim:
        dt = 0.001

    class domain_rand:
        randomize_friction = True
        friction_range = [0.1, 1]
        randomize_base_mass = True
        added_mass_range = [-2.0, 4.0]
        push_robots = True
        push_interval_s = 4
        max_push_vel_xy = 0.3
        max_push_ang_vel = 0.4
        action_delay = 0.0
        action_noise = 0.02

    class commands:
        curriculum = False
        num_commands = 4
        resampling_time = 8.0
        heading_command = True

        class ranges:
            lin_vel_x = [-0.6, 0.6]
            lin_vel_y = [-0.3, 0.3]
            ang_vel_yaw = [-0.3, 0.3]
            heading = [-3.14, 3.14]

    class rewards:
        base_height_target = 0.55
        min_dist = 0.1
        max_dist = 0.5
        target_joint_pos_scale = 0.17
        target_feet_height = 0.06
        cycle_time = 0.64
        only_positive_rewards = True
        tracking_sigma = 5
        max_contact_force = 180

        class scales:
            joint_pos = 0.0
            feet_clearance = 1.5
            feet_contact_number = 2.5
            feet_air_time = 2.0
            foot_slip = -0.05
            feet_distance = 0.2
            knee_distance = 0.2
            feet_contact_forces = -0.05
            tracking_lin_vel = 2.5
            tracking_ang_vel = 1.5
            vel_mismatch_exp = 0.0
            low_speed = 0.0
            track_vel_hard = 0.0
            default_joint_pos = 1.7
            orientation = 2
            base_height = 1.0
            base_acc = 0.3
            action_smoothness = -0.008
            torques = -1e-5
            dof_vel = -1e-4
            dof_acc = -1e-6
            collision = -0.5

    class normalization:
        class obs_scales:
            lin_vel = 2.0
            ang_vel = 1.0
            dof_pos = 1.0
            dof_vel = 0.05
            quat = 1.0
            height_measurements = 5.0
        clip_observations = 100
        clip_actions = 100


class HectorCfgPPO(_Cfg):
    seed = 5
    runner_class_name = "OnPolicyRunner"

    class policy:
        init_noise_std = 1.0
        actor_hidden_dims = [512, 256, 128]
        critic_hidden_dims = [768, 256, 128]

    class algorithm:
        value_loss_coef = 1.0
        use_clipped_value_loss = True
        clip_param = 0.2
        entropy_coef = 0.001
        num_learning_epochs = 2
        num_mini_batches = 4
        learning_rate = 1e-5
        schedule = "adaptive"
        gamma = 0.994
        lam = 0.9
        desired_kl = 0.01
        max_grad_norm = 1.0

    class runner:
        policy_class_name = "ActorCritic"
        algorithm_class_name = "PPO"
        num_steps_per_env = 60
        max_iterations = 10001
        save_interval = 100
        experiment_name = "hector"
        run_name = ""
        resume = False
        load_run = -1
        checkpoint = -1
        resume_path = None


def class_to_dict(obj) -> dict:
    """Attribute tree -> dict, keys in `dir()` (alphabetical) order.  The alphabetical
    order is what fixes the reward accumulation order (humanoid/utils/helpers.py:43-58)."""
    if not hasattr(obj, "__dict__"):
        return obj
    out = {}
    for key in dir(obj):
        if key.startswith("_"):
            continue
        val = getattr(obj, key)
        out[key] = [class_to_dict(v) for v in val] if isinstance(val, list) else class_to_dict(val)
    return out
