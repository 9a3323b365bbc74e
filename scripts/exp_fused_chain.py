"""Experiment: the persistent fused GEMM chain (hz_chain.cu) against the cuBLASLt chain on the same plan.
    python scripts/exp_fused_chain.py [N] [small]
Prints max abs/rel differences per output buffer and the per-chain time of both executors inside a CUDA graph."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hanabizero_b200 import _lib
from hanabizero_b200.model import MuZeroNet, MuZeroNetFull

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
small = len(sys.argv) > 2 and sys.argv[2] == "small"
dev = torch.device("cuda")
torch.manual_seed(0)
model = (MuZeroNet(193 * 4, 11) if small else MuZeroNetFull(785 * 4, 20)).randomize_heads().to(dev).eval()
plan = model.recurrent_plan(torch.float16)
ch = plan.chain(N)
lib = _lib.load()
print("fused grid:", lib.hz_gemm_plan_fused(ch._h), "steps:", ch.n_steps, flush=True)
ch.x0.copy_(torch.randn(N, plan.KP, device=dev) * 0.5)
st = torch.cuda.current_stream().cuda_stream
bufs = {"y1": ch.y1, "y2": ch.y2, "state": ch.state, "h1": ch.h1, "out": ch.out}
if ch.xb is not None:
    bufs["xb"] = ch.xb


def run_lt():
    _lib.check(lib.hz_gemm_plan_run(ch._h, st, 0, 1))
    _lib.check(lib.hz_gemm_plan_run(ch._h, st, 1, ch.n_steps - 1))


run_lt()
torch.cuda.synchronize()
ref = {k: v.clone() for k, v in bufs.items()}
for v in bufs.values():
    v.zero_()
_lib.check(lib.hz_gemm_plan_run(ch._h, st, 0, ch.n_steps))
torch.cuda.synchronize()
print("fused run returned", flush=True)
for k, v in bufs.items():
    d = (v.float() - ref[k].float()).abs()
    print(f"{k:6s} max|diff| {d.max().item():.4e}  max|ref| {ref[k].float().abs().max().item():.3e}  "
          f"mismatch>1e-2: {(d > 1e-2 + 1e-2 * ref[k].float().abs()).sum().item()}", flush=True)
# second launch: the barrier counters must have been restored
for v in bufs.values():
    v.zero_()
_lib.check(lib.hz_gemm_plan_run(ch._h, st, 0, ch.n_steps))
torch.cuda.synchronize()
print("second launch ok, state diff", (ch.state.float() - ref["state"].float()).abs().max().item(), flush=True)


def timed(fn, reps=49):
    g = torch.cuda.CUDAGraph()
    fn(); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    return best


st = None


def lt_graph():
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.hz_gemm_plan_run(ch._h, s, 0, 1))
    _lib.check(lib.hz_gemm_plan_run(ch._h, s, 1, ch.n_steps - 1))


def fused_graph():
    _lib.check(lib.hz_gemm_plan_run(ch._h, torch.cuda.current_stream().cuda_stream, 0, ch.n_steps))


print(f"cuBLASLt chain: {timed(lt_graph):.2f} us   fused chain: {timed(fused_graph):.2f} us", flush=True)

if os.environ.get("HZ_CHAIN_TRACE") == "1":
    import ctypes as C
    import numpy as np
    grid = lib.hz_gemm_plan_fused(ch._h)
    raw = C.CDLL(_lib.LIB_PATH)
    buf = np.zeros(grid * 8 * 16, np.uint64)
    _lib.check(lib.hz_gemm_plan_run(ch._h, torch.cuda.current_stream().cuda_stream, 0, ch.n_steps))
    torch.cuda.synchronize()
    raw.hz_debug_chain_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    raw.hz_debug_chain_trace(ch._h, buf.ctypes.data, buf.size)
    t = buf.reshape(grid, 8, 16).astype(np.int64)
    t0 = t[t > 0].min()
    names = ["pre-wait", "barrier passed", "mma issued", "acc ready", "stores done", "arrived", "tmem drained", "A issued", "-", "-", "-", "-", "start"]
    for cta in (0, grid // 2, grid - 1):
        print(f"CTA {cta}: ns since kernel's first stamp")
        for s_ in range(ch.n_steps):
            print("  step", s_, " ".join(f"{names[j]}={t[cta, s_, j] - t0 if t[cta, s_, j] else -1}" for j in (12, 0, 1, 7, 2, 3, 6, 4, 5)))
    arr = t[:, :ch.n_steps, 5]
    print("per-step last arrival - first arrival (ns):", (arr.max(0) - arr.min(0)).tolist())
    print("per-step: last 'arrived' -> median 'barrier passed' of next step (ns):",
          [int(np.median(t[:, s_ + 1, 1]) - arr[:, s_].max()) for s_ in range(ch.n_steps - 1)])
