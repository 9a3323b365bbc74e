"""Profiling driver: the row-block resident network executor alone (hz_rowchain_run), back to back, for
`ncu --kernel-name regex:k_row_chain`.   N=4096 python scripts/prof_rowchain.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hanabizero_b200.model import MuZeroNetFull
from hanabizero_b200.plan import BoundChain

dev = torch.device("cuda"); N = int(os.environ.get("N", "4096"))
torch.manual_seed(0)
model = MuZeroNetFull(785 * 4, 20).randomize_heads().to(dev).eval()
plan = model.recurrent_plan(torch.float16)
ch = BoundChain(plan, N)
ch.x0.zero_()
ch.x0[:, :plan.F].copy_((torch.randn(N, plan.F, device=dev).clamp_min(0) * 0.7).half())
ch.x0[:, plan.F:].scatter_(1, torch.randint(0, plan.A, (N, 1), device=dev), 1.0)
ch.set_executor("rows")
pool = torch.zeros(4, N, plan.F, device=dev, dtype=torch.float16)
st = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(4):
    ch.bind_state(pool[i % 4]); ch.run(st)
e0.record()
for i in range(8):
    ch.bind_state(pool[i % 4]); ch.run(st)
e1.record(); torch.cuda.synchronize()
print(f"N={N}: {e0.elapsed_time(e1) * 1e3 / 8:.1f} us per launch (eager, one stream), out checksum {ch.out.float().abs().sum().item():.3f}")
