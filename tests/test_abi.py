"""CPU: the C-ABI library loads and exports every function include/hzb200.h declares, and the ctypes
binding lists exactly those (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

from helpers import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "hzb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hz_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_surface():
    names = _declared()
    for must in ("hz_trees_create", "hz_trees_prepare", "hz_trees_traverse", "hz_trees_backprop",
                 "hz_trees_backprop_traverse", "hz_trees_root_stats", "hz_gather_hidden", "hz_envs_create",
                 "hz_envs_reset", "hz_envs_step", "hz_envs_observe", "hz_envs_step_observe", "hz_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from hanabizero_b200.build import build_library
    lib = ctypes.CDLL(build_library())
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, f"declared in include/hzb200.h but not exported: {missing}"


def test_binding_matches_header():
    from hanabizero_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load()
    assert lib.hz_version() >= 100
    assert lib.hz_launch_count() == 0           # nothing launched: no GPU work in CPU tests


def test_bad_arguments_fail_loudly_without_a_gpu():
    from hanabizero_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.hz_trees_create(ctypes.byref(h), 0, 0, 20, 50) == _lib.HZ_ERR_ARG
    assert b"num_trees" in lib.hz_last_error()
    assert lib.hz_trees_create(ctypes.byref(h), 0, 8, 33, 50) == _lib.HZ_ERR_ARG     # > 32 actions
    assert lib.hz_envs_create(ctypes.byref(h), 0, 4, 7, None) == _lib.HZ_ERR_ARG
    assert lib.hz_trees_prepare(None, None, 0.25, None, None, None, None) == _lib.HZ_ERR_ARG
    with pytest.raises(_lib.HzError):
        _lib.check(lib.hz_trees_root_stats(None, None, None, None))


def test_product_code_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under hanabizero_b200/ may reference it."""
    pkg = os.path.join(ROOT, "hanabizero_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "oracle/_ref" not in text and "liboracle" not in text, f


def test_dropin_import_paths_resolve_to_the_b200_modules():
    """dropin/ first on PYTHONPATH: the reference's own import lines reach hanabizero_b200."""
    import subprocess
    import sys
    code = ("import core.ctree.cytree as tree; from core.mcts import MCTS; from envs import HanabiEnv; "
            "from envs.hanabi.rl_env import HanabiEnv as H2; "
            "from config.hanabi_control.env_wrapper import HanabiControlWrapper; "
            "from config.hanabi_control.model import MuZeroNetFull; "
            "assert tree.Roots.__module__ == 'hanabizero_b200.cytree' and MCTS.__module__ == 'hanabizero_b200.mcts'; "
            "assert HanabiEnv is H2 and hasattr(tree, 'multi_traverse') and hasattr(tree, 'batch_traverse')")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "dropin"), ROOT]))
    subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd="/tmp")
