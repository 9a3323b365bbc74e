"""Profiling driver: the fused env launch (step + auto-reset + observe, float32 observations) on N games of random legal
play, for `ncu --kernel-name regex:k_env`.   N=4096 python scripts/prof_env.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hanabizero_b200.hanabi_env import HanabiVecEnv

N = int(os.environ.get("N", "4096"))
env = HanabiVecEnv(N, "Hanabi-Full", np.arange(N))
g, l, legal = env.reset_all()
gen = torch.Generator(device="cuda").manual_seed(0)
for t in range(60):
    acts = torch.multinomial(legal, 1, generator=gen).view(-1).int()
    g, l, legal, r, d, s = env.step_all(acts, auto_reset=True, want_local=False)
env.check()
print("done", N)
