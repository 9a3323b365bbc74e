"""envs.hanabi.rl_env -> HanabiEnv on the game kernels (envs/hanabi/rl_env.py:26-442)."""
from hanabizero_b200.hanabi_env import Discrete, HanabiEnv, HanabiVecEnv  # noqa: F401
