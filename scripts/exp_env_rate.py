"""Device-resident random-play rate of the env kernel (fused policy, ten steps per CUDA graph), as bench.py times it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hanabizero_b200.hanabi_env import HanabiVecEnv

for N in (4096, 65536):
    env = HanabiVecEnv(N, "Hanabi-Full", np.arange(N))
    env.reset_all(observe=False)
    buf = torch.zeros(N, dtype=torch.int32, device="cuda")
    env.set_random_policy(buf, seed=3)
    env.observe(want_local=False)
    for _ in range(30):
        env.step_all(buf, auto_reset=True, want_local=False)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            env.step_all(buf, auto_reset=True, want_local=False)
    g.replay(); torch.cuda.synchronize()
    best = 0.0
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        best = max(best, N * 300 / (e0.elapsed_time(e1) * 1e-3))
    env.check()
    print(f"N={N}: {best / 1e6:.1f} M steps/s ({N / best * 1e6:.2f} us per step)", flush=True)
