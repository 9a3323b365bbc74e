"""TEST INFRASTRUCTURE — ctypes loaders for the CPU oracle (oracle/_build/liboracle.so, the plain-C
restatement) and, when built, the unmodified reference compiled into oracle/_ref/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  Nothing under hanabizero_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(ref=True):
    """Compile the oracle (and oracle/_ref when /root/reference exists). Building is not using."""
    target = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-C", HERE] + target, check=True, stdout=subprocess.DEVNULL)


def _load(path):
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} not built: run `make -C oracle`")
    return C.CDLL(path)


def _nullable(arr, dtype):
    if arr is None:
        return None
    a = np.ascontiguousarray(arr, dtype=dtype)
    return a


class TreeEngine:
    """Same driver API over either the C restatement (prefix 'otree') or the compiled reference
    (prefix 'ref_trees', libref_ctree_{det,stock}.so)."""

    def __init__(self, lib, prefix, num, actions, sims, delta=0.006):
        self.lib, self.p = lib, prefix
        self.num, self.actions, self.sims = num, actions, sims
        f = lambda name: getattr(lib, f"{prefix}_{name}")
        f("new").restype = C.c_void_p
        f("new").argtypes = [C.c_int, C.c_int, C.c_int, C.c_float]
        f("free").argtypes = [C.c_void_p]
        f("prepare").argtypes = [C.c_void_p, C.c_float, C.c_void_p, _f32p, _f32p, _i32p]
        f("traverse").argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, _i32p, _i32p, _i32p]
        f("backprop").argtypes = [C.c_void_p, C.c_int, C.c_float, _f32p, _f32p, _f32p]
        f("stats").argtypes = [C.c_void_p, _i32p, _f32p, _f32p]
        f("root_priors").argtypes = [C.c_void_p, _f32p]
        f("trajectories").argtypes = [C.c_void_p, _i32p, C.c_int]
        f("path_len").argtypes = [C.c_void_p, C.c_int]
        f("path_len").restype = C.c_int
        f("expanded_stats").argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, _i32p, C.c_int]
        f("expanded_stats").restype = C.c_int
        self._f = f
        self.h = f("new")(num, actions, sims, delta)

    def __del__(self):
        if getattr(self, "h", None):
            self._f("free")(self.h)
            self.h = None

    def set_tie(self, mode, seed=0, tree_offset=0):
        """Tie rule (C restatement only): 0 = first of the tie list, 1 = counter-based uniform draw (hz::tie_hash)."""
        fn = self._f("set_tie")
        fn.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int]
        fn.restype = None
        fn(self.h, int(mode), int(seed) & (2 ** 64 - 1), int(tree_offset))

    def prepare(self, frac, noises, rewards, logits, masks):
        nz = None if noises is None else np.ascontiguousarray(noises, np.float32)
        self._f("prepare")(self.h, frac, None if nz is None else nz.ctypes.data,
                           np.ascontiguousarray(rewards, np.float32),
                           np.ascontiguousarray(logits, np.float32),
                           np.ascontiguousarray(masks, np.int32))

    def traverse(self, pb_c_base, pb_c_init, discount):
        ix = np.empty(self.num, np.int32)
        iy = np.empty(self.num, np.int32)
        la = np.empty(self.num, np.int32)
        self._f("traverse")(self.h, pb_c_base, pb_c_init, discount, ix, iy, la)
        return ix, iy, la

    def path_lens(self):
        return np.array([self._f("path_len")(self.h, i) for i in range(self.num)], np.int32)

    def backprop(self, x, discount, rewards, values, logits):
        self._f("backprop")(self.h, x, discount, np.ascontiguousarray(rewards, np.float32),
                            np.ascontiguousarray(values, np.float32),
                            np.ascontiguousarray(logits, np.float32))

    def stats(self):
        visits = np.empty((self.num, self.actions), np.int32)
        values = np.empty(self.num, np.float32)
        minmax = np.empty((self.num, 2), np.float32)
        self._f("stats")(self.h, visits, values, minmax)
        return visits, values, minmax

    def root_priors(self):
        out = np.empty((self.num, self.actions), np.float32)
        self._f("root_priors")(self.h, out)
        return out

    def trajectories(self, max_len):
        out = np.empty((self.num, max_len), np.int32)
        self._f("trajectories")(self.h, out, max_len)
        return out

    def expanded_stats(self, i, cap):
        r = np.zeros(cap, np.float32)
        vs = np.zeros(cap, np.float32)
        vc = np.zeros(cap, np.int32)
        n = self._f("expanded_stats")(self.h, i, r, vs, vc, cap)
        return r[:n], vs[:n], vc[:n]


def oracle_tree(num, actions, sims, delta=0.006):
    return TreeEngine(_load(ORACLE_SO), "otree", num, actions, sims, delta)


def ref_tree(num, actions, sims, delta=0.006, deterministic=True):
    so = "libref_ctree_det.so" if deterministic else "libref_ctree_stock.so"
    return TreeEngine(_load(os.path.join(REF_DIR, so)), "ref_trees", num, actions, sims, delta)


def have_ref():
    return all(os.path.exists(os.path.join(REF_DIR, s))
               for s in ("libref_ctree_det.so", "libref_hanabi.so"))


class HanabiGameCPU:
    """One scalar game over either the C restatement ('ohanabi') or the compiled reference
    ('ref_env', libref_hanabi.so). preset 0 = Hanabi-Full, 1 = Hanabi-Small."""

    def __init__(self, lib, prefix, preset, seed):
        f = lambda name: getattr(lib, f"{prefix}_{name}")
        f("new").restype = C.c_void_p
        f("new").argtypes = [C.c_int, C.c_int]
        f("free").argtypes = [C.c_void_p]
        f("dims").argtypes = [C.c_void_p, _i32p]
        f("reset").argtypes = [C.c_void_p, _i32p, _i32p, _i32p]
        f("step").argtypes = [C.c_void_p, C.c_int, _i32p, _i32p, _i32p, _i32p]
        f("step").restype = C.c_int if prefix == "ohanabi" else None
        f("dump").argtypes = [C.c_void_p, _i32p]
        f("dump").restype = C.c_int
        f("play").argtypes = [C.c_void_p, C.c_long, C.c_uint, C.POINTER(C.c_long)]
        f("play").restype = C.c_long
        self._f = f
        self.h = f("new")(preset, seed)
        d = np.zeros(9, np.int32)
        f("dims")(self.h, d)
        (self.enc_len, self.own_len, self.players, self.actions, self.colors, self.ranks,
         self.hand_size, self.max_info, self.max_life) = [int(x) for x in d]
        self.local_dim = self.enc_len + self.players
        self.global_dim = self.own_len + self.local_dim

    def __del__(self):
        if getattr(self, "h", None):
            self._f("free")(self.h)
            self.h = None

    def _bufs(self):
        return (np.zeros(self.global_dim, np.int32), np.zeros(self.local_dim, np.int32),
                np.zeros(self.actions, np.int32))

    def reset(self):
        g, l, m = self._bufs()
        self._f("reset")(self.h, g, l, m)
        return g, l, m

    def step(self, action):
        g, l, m = self._bufs()
        rds = np.zeros(3, np.int32)
        rc = self._f("step")(self.h, int(action), g, l, m, rds)
        if rc is not None and rc < 0:
            raise ValueError(f"illegal action {action}")
        return g, l, m, int(rds[0]), bool(rds[1]), int(rds[2])

    def dump(self):
        out = np.zeros(256, np.int32)
        n = self._f("dump")(self.h, out)
        return out[:n].copy()

    def play(self, steps, lcg_seed=1):
        chk = C.c_long(0)
        self._f("play")(self.h, steps, lcg_seed, C.byref(chk))
        return chk.value


def oracle_hanabi(preset, seed):
    return HanabiGameCPU(_load(ORACLE_SO), "ohanabi", preset, seed)


def ref_hanabi(preset, seed):
    return HanabiGameCPU(_load(os.path.join(REF_DIR, "libref_hanabi.so")), "ref_env", preset, seed)
