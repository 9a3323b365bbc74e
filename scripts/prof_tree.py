"""Profiling driver: the fused tree step alone (hz_trees_search_step on synthetic network outputs), one launch per
simulation, for `ncu --kernel-name regex:k_search_step`.   N=512 S=50 STAGE_LIMIT=0 python scripts/prof_tree.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hanabizero_b200 import _lib, cytree
from hanabizero_b200.model import MuZeroNetFull

dev = torch.device("cuda"); N = int(os.environ.get("N", "4096")); A, S, F = 20, int(os.environ.get("S", "50")), 512
lib = _lib.load()
torch.manual_seed(0)
model = MuZeroNetFull(785 * 4, A).randomize_heads().to(dev).eval()
plan = model.recurrent_plan(torch.float16); ch = plan.chain(N)
rng = np.random.default_rng(0)
noise = torch.from_numpy(rng.dirichlet([0.3] * A, N).astype(np.float32)).to(dev)
roots = cytree.Roots(N, A, S); roots.prepare(0.25, noise, torch.zeros(N, device=dev), torch.randn(N, A, device=dev), torch.ones(N, A, dtype=torch.int32, device=dev))
mm = cytree.MinMaxStatsList(N); mm.set_delta(0.006)
pool = torch.rand(S, N, F, device=dev).half()
io = _lib.SearchIO()
io.value_logits, io.ld_value = ch.value_logits.data_ptr(), ch.value_logits.stride(0)
io.reward_logits, io.ld_reward = ch.reward_logits.data_ptr(), ch.reward_logits.stride(0)
io.policy_logits, io.ld_policy = ch.policy_logits.data_ptr(), ch.policy_logits.stride(0)
io.next_state, io.ld_state = None, 0
io.support, io.support_width, io.support_delta = plan.support.data_ptr(), plan.n_support, 1.0
io.elem_bytes, io.sanitize_nan = 2, 1
io.pool, io.state_cols = pool.data_ptr(), F
io.out_batch, io.ld_batch, io.onehot_cols = ch.x0.data_ptr(), ch.x0.stride(0), plan.OH
io.minmax, io.value_delta_max = mm.tensor(dev).data_ptr(), 0.006
io.discount, io.pb_c_base, io.pb_c_init = 0.999, 19652, 1.25
io.stage_limit = int(os.environ.get("STAGE_LIMIT", "0"))     # 4 = the setting of searches in flight
st = torch.cuda.current_stream().cuda_stream; ref = ctypes.byref(io)
gen = torch.Generator(device=dev).manual_seed(1)
outs = [torch.randn(ch.out.shape, device=dev, generator=gen).half() for _ in range(8)]
_lib.check(lib.hz_trees_search_step(roots.handle, st, 0, 1, ref))
for x in range(1, S - 1):
    ch.out.copy_(outs[x % 8])
    _lib.check(lib.hz_trees_search_step(roots.handle, st, x, 1, ref))
torch.cuda.synchronize()
print("done", N, S)
