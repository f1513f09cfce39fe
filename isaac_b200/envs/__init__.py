"""Env stage of the hector hot path: the drop-in env class and its config (reference: humanoid/envs)."""
from .hector_config import HectorCfg
from .hector_env import HectorFreeEnvB200, build_env_params

__all__ = ["HectorCfg", "HectorFreeEnvB200", "build_env_params"]
