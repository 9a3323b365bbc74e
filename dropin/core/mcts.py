"""core.mcts -> device-resident MCTS.run_multi (core/mcts.py:7-57)."""
from hanabizero_b200.mcts import MCTS, SearchConfig  # noqa: F401
