"""Experiment: do the kernels of independent searches overlap on the GPU?  G graphs (49 GEMM chains each, or 49 tree
steps each) on G streams: time for G concurrent graphs vs one graph alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hanabizero_b200 import _lib
from hanabizero_b200.model import MuZeroNetFull
from hanabizero_b200.plan import BoundChain

dev = torch.device("cuda"); N = int(os.environ.get("N", "512")); GMAX = int(os.environ.get("G", "8"))
lib = _lib.load()
torch.manual_seed(0)
model = MuZeroNetFull(785 * 4, 20).randomize_heads().to(dev).eval()
plan = model.recurrent_plan(torch.float16)
chains = [BoundChain(plan, N) for _ in range(GMAX)]
streams = [torch.cuda.Stream() for _ in range(GMAX)]
TARGET = int(os.environ.get("SM_TARGET", "0"))
for ch in chains:
    ch.x0.copy_(torch.rand_like(ch.x0.float()).half())
    if TARGET:
        ch.set_sm_target(TARGET)
print(f"SM target {TARGET or 'whole device'}")

def make_graph(ch, first, count, reps=49):
    st0 = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.hz_gemm_plan_run(ch._h, st0, first, count)); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(reps):
            _lib.check(lib.hz_gemm_plan_run(ch._h, st, first, count))
    return g

def run(graphs, G, rounds=3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record()
    for s in streams[:G]:
        s.wait_stream(main)
    for _ in range(rounds):
        for g, s in zip(graphs[:G], streams[:G]):
            with torch.cuda.stream(s):
                g.replay()
    for s in streams[:G]:
        main.wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / rounds

for label, first, count in (("whole chain (7 GEMMs)", 0, 7), ("L2 512->512 only", 1, 1)):
    graphs = [make_graph(ch, first, count) for ch in chains]
    run(graphs, GMAX)
    base = run(graphs, 1)
    print(f"N={N} {label}: 1 graph {base / 49:.2f} us per chain", flush=True)
    for G in (2, 4, 8):
        if G > GMAX: break
        t = run(graphs, G)
        print(f"    {G} graphs on {G} streams: {t / 49:.2f} us per round of {G} = {t / 49 / G:.2f} us per chain-equivalent ({base * G / t:.2f}x overlap)", flush=True)
