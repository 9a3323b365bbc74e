"""core.ctree.cytree -> the sm_100a tree engine (same names as core/ctree/cytree.pyx:17-101)."""
from hanabizero_b200.cytree import *  # noqa: F401,F403
from hanabizero_b200.cytree import (MinMaxStatsList, Node, ResultsWrapper, Roots, batch_back_propagate,  # noqa: F401
                                    batch_traverse, multi_back_propagate, multi_traverse)
