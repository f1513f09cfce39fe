"""rsl_rl-compatible RL stage (reference: humanoid/algo/__init__.py exports PPO, ActorCritic, RolloutStorage)."""
from .actor_critic import ActorCritic
from .ppo import PPO
from .rollout_storage import RolloutStorage

__all__ = ["ActorCritic", "PPO", "RolloutStorage"]
