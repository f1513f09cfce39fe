"""rsl_rl-compatible RL stage (reference: humanoid/algo/__init__.py exports PPO, ActorCritic, RolloutStorage and the
runner)."""
from .actor_critic import ActorCritic
from .on_policy_runner import OnPolicyRunner
from .ppo import PPO
from .rollout_storage import RolloutStorage

__all__ = ["ActorCritic", "OnPolicyRunner", "PPO", "RolloutStorage"]
