"""Experiment: throughput of one GPU as a function of the trees in flight — one search of n trees, or G concurrent
searches of 4096 trees on G streams (independent root batches, e.g. two actors sharing a GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from hanabizero_b200 import _lib, cytree
from hanabizero_b200.mcts import MCTS, SearchConfig
from hanabizero_b200.model import MuZeroNetFull
from hanabizero_b200.plan import RecurrentPlan

dev = torch.device("cuda")
A, S, F = 20, 50, 512
torch.manual_seed(0)
model = MuZeroNetFull(785 * 4, A).randomize_heads().to(dev).eval()
cfg = SearchConfig(num_simulations=S, amp_type="torch_amp")


def measure(n, G):
    rng = np.random.default_rng(0)
    groups = []
    for g in range(G):
        m = MCTS(cfg)
        m._plan = (lambda p: (lambda model: p))(RecurrentPlan(model, torch.float16))   # private chain buffers per group
        roots = cytree.Roots(n, A, S)
        d = dict(noise=torch.from_numpy(rng.dirichlet([0.3] * A, n).astype(np.float32)).to(dev),
                 logits=torch.randn(n, A, device=dev), legal=torch.ones(n, A, dtype=torch.int32, device=dev),
                 hidden=torch.rand(n, F, device=dev).half(), rew=torch.zeros(n, device=dev))
        groups.append((m, roots, d, torch.cuda.Stream()))

    def run_all():
        main = torch.cuda.current_stream()
        for m, roots, d, s in groups:
            s.wait_stream(main)
            with torch.cuda.stream(s):
                roots.prepare(0.25, d["noise"], d["rew"], d["logits"], d["legal"])
                m.run_multi(roots, model, d["hidden"], use_graph=False)
        for m, roots, d, s in groups:
            main.wait_stream(s)

    run_all()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run_all()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{G} concurrent search(es) of {n} trees: {ms:.3f} ms -> {G * n * (S - 1) / ms / 1e3:.1f} M simulations/s", flush=True)
    for m, roots, d, s in groups:
        _lib.check(roots._lib.hz_trees_set_progress(roots.handle, S - 1))


for n, G in ((4096, 1), (8192, 1), (16384, 1), (4096, 2), (4096, 3), (8192, 2)):
    measure(n, G)
