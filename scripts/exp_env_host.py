"""Where does a host-driven env step go?  Per-phase host times of the EnvPipeline loop (wait / pick / submit)."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hanabizero_b200.hanabi_env import EnvPipeline, HanabiVecEnv

N, T = 4096, 300
for G in (1, 2, 4):
    for zero_copy in (True, False):
        n = N // G
        envs = [HanabiVecEnv(n, "Hanabi-Full", np.arange(n) + 1000 * g) for g in range(G)]
        for e in envs:
            e.reset_all(observe=False)
        pipe = EnvPipeline(envs, fmt="bits", zero_copy=zero_copy)
        acts = [pipe.actions(g) for g in range(G)]
        for g in range(G):
            pipe.observe_now(g)
        tw = tp = ts = 0.0

        def one(g, step, acc):
            t0 = time.perf_counter()
            obs, meta = pipe.wait(g)
            t1 = time.perf_counter()
            envs[g].random_legal_host(meta, acts[g], seed=1, step=step)
            t2 = time.perf_counter()
            pipe.step(g)
            t3 = time.perf_counter()
            acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2

        for step in range(5):
            for g in range(G):
                one(g, step, [0, 0, 0])
        accs = [[0.0, 0.0, 0.0] for _ in range(G)]
        # (a) one thread round-robin
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for step in range(5, 5 + T):
            for g in range(G):
                one(g, step, accs[g])
        pipe.drain(); dt = time.perf_counter() - t0
        a = np.array(accs).sum(0) / (T * G) * 1e6
        print(f"G={G} zero_copy={zero_copy} one thread : {N * T / dt / 1e6:6.1f} M steps/s; per group-step: wait {a[0]:.1f} us, pick {a[1]:.1f} us, submit {a[2]:.1f} us", flush=True)
        # (b) one thread per group
        accs = [[0.0, 0.0, 0.0] for _ in range(G)]
        def drive(g):
            for step in range(5 + T, 5 + 2 * T):
                one(g, step, accs[g])
        ths = [threading.Thread(target=drive, args=(g,)) for g in range(G)]
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for th in ths: th.start()
        for th in ths: th.join()
        pipe.drain(); dt = time.perf_counter() - t0
        a = np.array(accs).sum(0) / (T * G) * 1e6
        print(f"G={G} zero_copy={zero_copy} {G} thread(s): {N * T / dt / 1e6:6.1f} M steps/s; per group-step: wait {a[0]:.1f} us, pick {a[1]:.1f} us, submit {a[2]:.1f} us", flush=True)
        for e in envs:
            e.check()
