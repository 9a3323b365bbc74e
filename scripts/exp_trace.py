"""Debug experiment: per-warp cycle stamps of the fused search-step kernel (needs a -DHZ_TRACE build:
   nvcc ... -DHZ_TRACE -o /tmp/libhz_trace.so; run with HZ_LIB=/tmp/libhz_trace.so)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hanabizero_b200 import _lib
_lib.LIB_PATH = os.environ.get("HZ_LIB", _lib.LIB_PATH)
from hanabizero_b200 import cytree
from hanabizero_b200.model import MuZeroNetFull

dev = torch.device("cuda"); N = int(os.environ.get("N", "4096")); A, S, F = 20, 50, 512
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
st = torch.cuda.current_stream().cuda_stream; ref = ctypes.byref(io)
trace = torch.zeros(N, 16, dtype=torch.int64, device=dev)
lib.hz_debug_set_trace.argtypes = [ctypes.c_void_p]
_lib.check(lib.hz_debug_set_trace(trace.data_ptr()))
_lib.check(lib.hz_trees_search_step(roots.handle, st, 0, 1, ref))
gen = torch.Generator(device=dev).manual_seed(1)
names = ["prologue+decode", "softmax+expand", "stage wait+backprop", "minmax", "traverse", "gather"]
for x in range(1, S - 1):
    ch.out.copy_(torch.randn(ch.out.shape, device=dev, generator=gen).half())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _lib.check(lib.hz_trees_search_step(roots.handle, st, x, 1, ref)); e1.record()
    torch.cuda.synchronize()
    if x in (10, 25, 45):
        tr = trace.cpu().numpy().astype(np.float64)
        d = np.diff(tr[:, :7], axis=1)
        depth = tr[:, 7] - 1
        tot = tr[:, 6] - tr[:, 0]
        print(f"sim {x}: kernel {e0.elapsed_time(e1)*1e3:.1f} us (eager, warm L2), depth mean {depth.mean():.2f} max {depth.max():.0f}; "
              f"per-warp total cycles mean {tot.mean():.0f} max {tot.max():.0f}")
        for k, nm in enumerate(names):
            print(f"    {nm:24s} mean {d[:, k].mean():8.0f}  p50 {np.median(d[:, k]):8.0f}  max {d[:, k].max():8.0f}")
        for a, b, nm in ((0, 8, "issue staging + loads"), (8, 1, "decode value+reward")):
            dd = tr[:, b] - tr[:, a]
            print(f"      {nm:24s} mean {dd.mean():8.0f}  p50 {np.median(dd):8.0f}  max {dd.max():8.0f}")
        lv = d[:, 4] / np.maximum(depth, 1)
        print(f"    traverse cycles per level: mean {lv.mean():.0f} p90 {np.percentile(lv, 90):.0f}")
        start, end = tr[:, 0], tr[:, 6]
        t0 = start.min()
        print(f"    warps start within {start.max() - t0:.0f} cycles; end-time percentiles (cycles after first start): "
              + ", ".join(f"p{q}={np.percentile(end - t0, q):.0f}" for q in (10, 50, 90, 99, 100)))
        deep = depth >= np.percentile(depth, 99)
        print(f"    deepest 1% of trees: depth {depth[deep].mean():.1f}, traverse {d[deep, 4].mean():.0f} cycles = {d[deep, 4].mean() / depth[deep].mean():.0f} per level")
