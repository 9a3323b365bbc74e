"""config.hanabi_control.model -> same architectures / parameter names (model.py:127-335)."""
from hanabizero_b200.model import HanabiMuZeroNet, MuZeroNet, MuZeroNetFull  # noqa: F401
