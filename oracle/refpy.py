"""TEST INFRASTRUCTURE — the reference's own PYTHON on the hot path, in installed (byte-compiled) form.

`install()` (run by `make -C oracle refpy`, only where /root/reference exists) byte-compiles, unmodified,

    core/mcts.py                 (MCTS.run_multi, the search driver the drop-in must serve)
    envs/hanabi/rl_env.py        (HanabiEnv.reset/step, the Python API the callers pay for)
    envs/hanabi/pyhanabi.py      (its cffi binding)

into oracle/_ref/refpy/ as sourceless byte-code files (*.rpyc), next to the cffi header pyhanabi.py parses at import and the
libpyhanabi.so built by oracle/Makefile.  Like every other file under oracle/_ref/ these are build outputs of the
reference (git-ignored, shipped to the GPU box); no reference source enters the repository.

Consumers (tests/, bench.py's cpu_baseline / --impl reference legs only):
    load_env_class()             -> the reference's HanabiEnv class
    load_mcts_class(cytree)      -> the reference's MCTS class bound to the given `core.ctree.cytree` module
                                    (the drop-in façade, or the reference's own Cython build)
"""
import importlib.machinery
import importlib.util
import os
import py_compile
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFPY = os.path.join(REF_DIR, "refpy")
FILES = ["core/mcts.py", "envs/hanabi/rl_env.py", "envs/hanabi/pyhanabi.py"]
EXT = ".rpyc"   # byte-code, but not named *.pyc: snapshot / ignore rules commonly drop that pattern


def install(reference="/root/reference"):
    for rel in FILES:
        dst = os.path.join(REFPY, rel[:-3] + EXT)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(os.path.join(reference, rel), cfile=dst, dfile=rel, doraise=True)
    hdir = os.path.join(REFPY, "envs", "hanabi")
    shutil.copyfile(os.path.join(reference, "envs", "hanabi", "pyhanabi.h"), os.path.join(hdir, "pyhanabi.h"))
    shutil.copyfile(os.path.join(REF_DIR, "libpyhanabi.so"), os.path.join(hdir, "libpyhanabi.so"))


def available():
    return all(os.path.exists(os.path.join(REFPY, rel[:-3] + EXT)) for rel in FILES)


def _load_pyc(name, rel, package=None):
    path = os.path.join(REFPY, rel[:-3] + EXT)
    loader = importlib.machinery.SourcelessFileLoader(name, path)
    spec = importlib.util.spec_from_loader(name, loader, origin=path)
    mod = importlib.util.module_from_spec(spec)
    mod.__file__ = path
    if package is not None:
        mod.__package__ = package
    sys.modules[name] = mod
    loader.exec_module(mod)
    return mod


def _package(name, path):
    pkg = types.ModuleType(name)
    pkg.__path__ = [path]
    pkg.__package__ = name
    sys.modules[name] = pkg
    return pkg


def load_env_class():
    """The reference's HanabiEnv (envs/hanabi/rl_env.py:87).  Shims for this image: numpy.int (rl_env.py:256,429) and
    a `gym.spaces.Discrete` stub (rl_env.py:21,141); envs/__init__.py's absl side effect is skipped."""
    import numpy as np
    if "refpy_envs.hanabi.rl_env" in sys.modules:
        return sys.modules["refpy_envs.hanabi.rl_env"].HanabiEnv
    if not hasattr(np, "int"):
        np.int = int
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        spaces = types.ModuleType("gym.spaces")

        class Discrete:
            def __init__(self, n):
                self.n = n

        spaces.Discrete = Discrete
        gym.spaces = spaces
        sys.modules["gym"], sys.modules["gym.spaces"] = gym, spaces
    _package("refpy_envs", os.path.join(REFPY, "envs"))
    _package("refpy_envs.hanabi", os.path.join(REFPY, "envs", "hanabi"))
    _load_pyc("refpy_envs.hanabi.pyhanabi", "envs/hanabi/pyhanabi.py", "refpy_envs.hanabi")
    mod = _load_pyc("refpy_envs.hanabi.rl_env", "envs/hanabi/rl_env.py", "refpy_envs.hanabi")
    assert sys.modules["refpy_envs.hanabi.pyhanabi"].lib_loaded_flag, "libpyhanabi.so not found next to pyhanabi.pyc"
    return mod.HanabiEnv


_mcts_serial = [0]


def load_mcts_class(cytree_module):
    """The reference's MCTS class (core/mcts.py:7-57), its `import core.ctree.cytree as tree` resolved to
    `cytree_module`.  Every call executes the .pyc afresh, so two bindings can coexist in one process."""
    saved = {k: sys.modules.get(k) for k in ("core", "core.ctree", "core.ctree.cytree")}
    core = types.ModuleType("core")
    core.__path__ = []
    ctree = types.ModuleType("core.ctree")
    ctree.__path__ = []
    core.ctree, ctree.cytree = ctree, cytree_module
    sys.modules.update({"core": core, "core.ctree": ctree, "core.ctree.cytree": cytree_module})
    try:
        _mcts_serial[0] += 1
        mod = _load_pyc(f"refpy_core_mcts_{_mcts_serial[0]}", "core/mcts.py")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod.MCTS


def load_ref_cytree(deterministic=True):
    """The reference's own Cython module built by oracle/Makefile: `det/` = compiled with the rand() == 0 shim (the
    parity contract), top level = stock rand().  Import AFTER numpy/torch."""
    import numpy  # noqa: F401
    d = os.path.join(REF_DIR, "det") if deterministic else REF_DIR
    so = [f for f in os.listdir(d) if f.startswith("cytree.") and f.endswith(".so")]
    spec = importlib.util.spec_from_file_location("cytree", os.path.join(d, so[0]))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    install(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("installed", REFPY)
