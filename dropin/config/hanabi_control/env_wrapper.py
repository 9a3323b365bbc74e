"""config.hanabi_control.env_wrapper -> HanabiControlWrapper (env_wrapper.py:6-34)."""
from hanabizero_b200.env_wrapper import Game, HanabiControlWrapper  # noqa: F401
