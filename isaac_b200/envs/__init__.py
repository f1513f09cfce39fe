"""Env stage of the hot path: the drop-in env classes and their configs (reference: humanoid/envs/__init__.py registers
`hector`, `hector_full` and `humanoid_ppo`)."""
from .hector_config import HectorCfg, HectorCfgPPO
from .hector_env import HectorFreeEnvB200, HectorFullFreeEnvB200, XBotLFreeEnvB200, build_env_params
from .tasks import TASKS, HectorFullCfg, HectorFullCfgPPO, XBotLCfg, XBotLCfgPPO, layout_for

__all__ = ["HectorCfg", "HectorCfgPPO", "HectorFreeEnvB200", "HectorFullFreeEnvB200", "XBotLFreeEnvB200", "HectorFullCfg",
           "HectorFullCfgPPO", "XBotLCfg", "XBotLCfgPPO", "TASKS", "layout_for", "build_env_params"]
