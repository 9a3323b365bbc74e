"""Experiment: in-graph time of the GEMM chain alone (per simulation), per step, and autotune effect."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hanabizero_b200 import _lib
from hanabizero_b200.model import MuZeroNetFull

dev = torch.device("cuda"); N = int(os.environ.get("N", "4096"))
lib = _lib.load()
torch.manual_seed(0)
model = MuZeroNetFull(785 * 4, 20).randomize_heads().to(dev).eval()
plan = model.recurrent_plan(torch.float16); ch = plan.chain(N)
ch.x0.copy_(torch.rand_like(ch.x0.float()).half())
def timed(first, count, reps=49):
    g = torch.cuda.CUDAGraph()
    st0 = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.hz_gemm_plan_run(ch._h, st0, first, count)); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(reps):
            _lib.check(lib.hz_gemm_plan_run(ch._h, st, first, count))
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (2 * reps)
print(f"N={N} autotune={os.environ.get('HZ_GEMM_AUTOTUNE','1')}: whole chain {timed(0, ch.n_steps):.1f} us per simulation")
names = ["L1 544->512", "L2 512->512", "L3 512->512 (+res)", "H1 512->768", "B2 3x(256->256)", "A2 256->256 (+res)", "B3 3x(256->208)"]
for i in range(ch.n_steps):
    print(f"   step {i} {names[i]:22s} {timed(i, 1):6.2f} us back-to-back")
