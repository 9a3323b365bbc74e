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


def test_ctypes_structs_mirror_the_header_layouts(tmp_path):
    """The structs passed by pointer through the C ABI: size and every field offset of the ctypes mirrors in
    hanabizero_b200/_lib.py equal what a C compiler makes of include/hzb200.h."""
    import subprocess
    from hanabizero_b200 import _lib
    mirrors = {"hz_search_io": _lib.SearchIO, "hz_traj_view": _lib.TrajView, "hz_gemm_step": _lib.GemmStep,
               "hz_rowchain_weights": _lib.RowChainWeights}
    lines = []
    for cname, cls in mirrors.items():
        lines.append(f'printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('printf("\\n");')
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "hzb200.h"\nint main(void) {\n' + "\n".join(lines) + "\nreturn 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    for line in filter(None, out):
        name, size, *offsets = line.split()
        cls = mirrors[name]
        assert ctypes.sizeof(cls) == int(size), name
        assert [getattr(cls, f).offset for f, _ in cls._fields_] == [int(o) for o in offsets], name


def test_rowchain_rejects_plans_it_was_not_built_for():
    """hz_rowchain covers the fp16 Hanabi-Full shapes only; anything else is an argument error, not a wrong answer."""
    from hanabizero_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    w = _lib.RowChainWeights()
    assert lib.hz_rowchain_create(ctypes.byref(h), 0, None, 128, None, 544, None, None) == _lib.HZ_ERR_ARG
    for f, _ in _lib.RowChainWeights._fields_:
        if f not in ("ld_w1", "state_cols", "head_cols", "onehot_cols", "logit_cols"):
            setattr(w, f, 0x1000)
    w.ld_w1, w.state_cols, w.head_cols, w.onehot_cols, w.logit_cols = 544, 512, 128, 32, 208      # Hanabi-Small heads
    buf = ctypes.c_void_p(0x1000)
    assert lib.hz_rowchain_create(ctypes.byref(h), 0, ctypes.byref(w), 128, buf, 544, buf, buf) == _lib.HZ_ERR_ARG
    assert b"512" in lib.hz_last_error()
    w.head_cols, w.logit_cols = 256, 201                                                            # unpadded logit rows
    assert lib.hz_rowchain_create(ctypes.byref(h), 0, ctypes.byref(w), 128, buf, 544, buf, buf) == _lib.HZ_ERR_ARG
    w.logit_cols = 208
    w.b1 = 0x1008                                                                                   # misaligned pointer
    assert lib.hz_rowchain_create(ctypes.byref(h), 0, ctypes.byref(w), 128, buf, 544, buf, buf) == _lib.HZ_ERR_ARG
    assert lib.hz_rowchain_run(None, None) == _lib.HZ_ERR_ARG and lib.hz_rowchain_grid(None) == 0
    assert lib.hz_rowchain_destroy(None) == _lib.HZ_OK
