"""Debug experiment: per-game cycle stamps of the env kernel's phases (needs the -DHZ_TRACE build:
   HZ_LIB=hanabizero_b200/csrc/libhzb200_trace.so python scripts/exp_env_trace.py)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from hanabizero_b200 import _lib

_lib.LIB_PATH = os.environ.get("HZ_LIB", _lib.LIB_PATH)
from hanabizero_b200.hanabi_env import HanabiVecEnv

N = 4096
env = HanabiVecEnv(N, "Hanabi-Full", np.arange(N))
lib = _lib.load()
g, l, legal = env.reset_all()
gen = torch.Generator(device="cuda").manual_seed(0)
trace = torch.zeros(N, 16, dtype=torch.int64, device="cuda")
lib.hz_debug_set_env_trace.argtypes = [ctypes.c_void_p]
for t in range(120):
    acts = torch.multinomial(legal, 1, generator=gen).view(-1).int()
    if t == 100:
        _lib.check(lib.hz_debug_set_env_trace(trace.data_ptr()))
        trace.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g, l, legal, r, d, s = env.step_all(acts, auto_reset=True, want_local=False)
    e1.record()
    torch.cuda.synchronize()
    if t >= 100 and t % 5 == 0:
        tr = trace.cpu().numpy().astype(np.float64)
        reset = (tr[:, 9] % 16) > 0
        names = [(0, 1, "legality check"), (1, 2, "apply_move (lane 0)"), (2, 3, "deal after the move"), (3, 4, "score/terminal (+ re-deal)"),
                 (4, 5, "state write-back"), (5, 6, "observation bits"), (6, 7, "expand + store floats"), (7, 8, "legal mask")]
        tot = tr[:, 8] - tr[:, 0]
        print(f"step {t}: kernel {e0.elapsed_time(e1) * 1e3:.1f} us; {int(reset.sum())} games re-dealt; cycles per game mean {tot.mean():.0f}, "
              f"max {tot.max():.0f}; re-dealt games mean {tot[reset].mean() if reset.any() else 0:.0f}")
        for a, b, nm in names:
            dd = tr[:, b] - tr[:, a]
            print(f"    {nm:28s} mean {dd.mean():7.0f}  p50 {np.median(dd):7.0f}  max {dd.max():7.0f}"
                  + (f"   re-dealt mean {dd[reset].mean():7.0f}" if reset.any() else ""))
        deals = tr[:, 9] // 16
        has = deals > 0
        if has.any():
            names2 = ["w = cnt/deck + sync", "sum chain (25 dadd)", "qn = w/sum + sync", "prefix chain (25 dadd)", "draws + pick", "apply deal (lane 0)"]
            print(f"    per deal ({int(deals.sum())} deals over the traced steps):")
            for k, nm in enumerate(names2):
                print(f"        {nm:26s} {tr[has, 10 + k].sum() / deals.sum():7.0f} cycles")
        trace.zero_()
env.check()
