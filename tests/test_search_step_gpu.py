"""GPU: hz_trees_search_step (the production launch: raw support logits in, one launch per simulation)
against the generic calls it fuses — hz_support_decode + hz_trees_backprop_traverse + a torch gather —
on synthetic network outputs.  Bit-exact trees and hand-off batches, including the register-q variants
(capacity <= 64 and > 64) and the deep-path fallback (path longer than a warp)."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import CONST, bits_equal, tree_inputs

pytestmark = pytest.mark.gpu


def _run(N, A, S, width, dtype, deep=False, seed=0):
    from hanabizero_b200 import _lib, cytree
    lib = _lib.load()
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(seed)
    F, OH = 512, 32 if A > 16 else 16
    P3 = (width + 7) // 8 * 8
    d = tree_inputs(N, A, S, seed)
    if deep:   # one forced line of play: path length grows by one every simulation
        d["logits"][:] = -30.0
        d["logits"][:, 0] = 30.0
        d["mask"][:] = 1
    support = (torch.arange(width, device=dev, dtype=torch.float32) - (width - 1) // 2)
    outs = []
    for x in range(S):
        o = torch.randn(3, N, P3, device=dev, generator=gen)
        if deep:
            o[2, :, :] = -30.0
            o[2, :, 0] = 30.0
            o[0, :, :] = -20.0
            o[0, :, (width - 1) // 2 + 3 + (x % 3)] = 20.0      # value ~ +3..5: the line stays attractive
            o[1, :, :] = -20.0
            o[1, :, (width - 1) // 2] = 20.0                    # reward 0
        outs.append(o.to(dtype))
    states = [torch.rand(N, F, device=dev, generator=gen).to(dtype) for _ in range(S)]
    root_hidden = torch.rand(N, F, device=dev, generator=gen).to(dtype)
    eb = outs[0].element_size()
    st = torch.cuda.current_stream().cuda_stream
    results = []
    for fused in (True, False):
        roots = cytree.Roots(N, A, S)
        roots.prepare(CONST["frac"], d["noise"], d["reward"], d["logits"], d["mask"])
        mm = cytree.MinMaxStatsList(N)
        mm.set_delta(CONST["delta"])
        mmt = mm.tensor(dev)
        pool = torch.zeros(S + 1, N, F, device=dev, dtype=dtype)
        pool[0] = root_hidden
        batch = torch.zeros(N, F + OH, device=dev, dtype=dtype)
        ix = torch.zeros(N, dtype=torch.int32, device=dev)
        la = torch.zeros(N, dtype=torch.int32, device=dev)
        trace = []
        if fused:
            io = _lib.SearchIO()
            cur = torch.zeros(3, N, P3, device=dev, dtype=dtype)
            state = torch.zeros(N, F, device=dev, dtype=dtype)
            io.value_logits, io.ld_value = cur[0].data_ptr(), P3
            io.reward_logits, io.ld_reward = cur[1].data_ptr(), P3
            io.policy_logits, io.ld_policy = cur[2].data_ptr(), P3
            io.next_state, io.ld_state = state.data_ptr(), F
            io.support, io.support_width, io.support_delta = support.data_ptr(), width, 1.0
            io.elem_bytes, io.sanitize_nan = eb, 1
            io.pool, io.state_cols = pool.data_ptr(), F
            io.out_batch, io.ld_batch, io.onehot_cols = batch.data_ptr(), F + OH, OH
            io.out_ix, io.out_action = ix.data_ptr(), la.data_ptr()
            io.minmax, io.value_delta_max = mmt.data_ptr(), CONST["delta"]
            io.discount, io.pb_c_base, io.pb_c_init = CONST["discount"], CONST["pb_c_base"], CONST["pb_c_init"]
            ref = ctypes.byref(io)
            _lib.check(lib.hz_trees_search_step(roots.handle, st, 0, 1, ref))
            trace.append((ix.clone(), la.clone(), batch.clone()))
            for x in range(1, S):
                cur.copy_(outs[x])
                state.copy_(states[x])
                _lib.check(lib.hz_trees_search_step(roots.handle, st, x, 1 if x < S - 1 else 0, ref))
                if x < S - 1:
                    trace.append((ix.clone(), la.clone(), batch.clone()))
        else:
            hidden = torch.zeros(N, F, device=dev, dtype=dtype)
            dec = torch.zeros(2 * N, device=dev)

            def record():
                b = torch.zeros(N, F + OH, device=dev, dtype=dtype)
                b[:, :F] = hidden
                b[torch.arange(N, device=dev), F + la.long()] = 1.0
                trace.append((ix.clone(), la.clone(), b))

            _lib.check(lib.hz_trees_traverse(roots.handle, st, CONST["pb_c_base"], CONST["pb_c_init"], CONST["discount"],
                                             mmt.data_ptr(), CONST["delta"], ix.data_ptr(), None, la.data_ptr(), None,
                                             pool.data_ptr(), hidden.data_ptr(), F * eb))
            record()
            for x in range(1, S):
                vr = outs[x][:2].reshape(2 * N, P3)
                _lib.check(lib.hz_support_decode(st, vr.data_ptr(), eb, support.data_ptr(), dec.data_ptr(), 2 * N, width, P3, 1.0))
                logits = outs[x][2][:, :A].float().contiguous()
                pool[x] = states[x]
                if x < S - 1:
                    _lib.check(lib.hz_trees_backprop_traverse(
                        roots.handle, st, x, CONST["discount"], dec[N:].data_ptr(), dec[:N].data_ptr(), logits.data_ptr(), 1,
                        mmt.data_ptr(), CONST["delta"], CONST["pb_c_base"], CONST["pb_c_init"], ix.data_ptr(), None,
                        la.data_ptr(), None, pool.data_ptr(), hidden.data_ptr(), F * eb))
                    record()
                else:
                    _lib.check(lib.hz_trees_backprop(roots.handle, st, x, CONST["discount"], dec[N:].data_ptr(),
                                                     dec[:N].data_ptr(), logits.data_ptr(), 1, mmt.data_ptr()))
        v, val = roots.get_stats_tensors()
        e = roots.export(S)
        results.append(dict(trace=trace, visits=v.cpu(), values=val.cpu(), mm=mmt.cpu().clone(), ev=e["visits"].cpu(),
                            evs=e["value_sum"].cpu(), er=e["reward"].cpu(), plen=e["path_len"].cpu(), pool=pool.cpu(),
                            traj=roots.get_trajectories()))
    a, b = results
    assert len(a["trace"]) == len(b["trace"]) == S - 1
    for s, ((ix0, la0, b0), (ix1, la1, b1)) in enumerate(zip(a["trace"], b["trace"])):
        assert torch.equal(ix0, ix1) and torch.equal(la0, la1), f"traverse differs at simulation {s}"
        assert torch.equal(b0, b1), f"hand-off batch differs at simulation {s}"
    assert torch.equal(a["visits"], b["visits"]) and torch.equal(a["ev"], b["ev"]) and torch.equal(a["plen"], b["plen"])
    for k in ("values", "mm", "evs", "er"):
        assert bits_equal(a[k].numpy(), b[k].numpy()), k
    assert torch.equal(a["pool"], b["pool"]) and a["traj"] == b["traj"]
    assert (a["visits"].sum(1) == S - 1).all()
    return a


@pytest.mark.parametrize("N,A,S,width,dtype", [
    (300, 20, 30, 201, torch.float16),     # Hanabi-Full shapes, q in 2 registers per lane
    (129, 11, 40, 51, torch.float32),      # Hanabi-Small shapes, float pool
    (64, 20, 130, 201, torch.float16),     # capacity > 64: 8 q registers per lane
    (33, 20, 300, 201, torch.float32),     # capacity > 256: general back-propagation routine
    (2048, 20, 200, 201, torch.float16),   # BASELINE configs[4] per-GPU shard: 16384 trees x 200 sims over 8 GPUs
])
def test_search_step_equals_generic_calls(N, A, S, width, dtype):
    _run(N, A, S, width, dtype)


def test_search_step_deep_path_fallback():
    r = _run(24, 20, 60, 201, torch.float16, deep=True)
    assert int(r["plen"].max()) > 40           # the path outgrew one warp: general routine taken
