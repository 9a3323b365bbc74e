"""GPU parity tests of the tree kernels, called through the cytree façade -> C ABI (libhzb200.so),
against (a) golden traces made by the unmodified reference and (b) the CPU oracle on the same
seeded inputs.  Bar: (ix, iy, last_action) per simulation, visit counts and selected actions
bit-exact; root values / min-max / per-node Q statistics within 1e-5 relative (north-star) and in
fact asserted bit-exact."""
import numpy as np
import pytest
import torch

from helpers import (CONST, bits_equal, golden_files, golden_tree_inputs, load_golden,
                     run_tree_lockstep, tree_inputs)

pytestmark = pytest.mark.gpu


def _engines(N, A, S):
    from gpu_adapters import GpuTreeEngine
    from oracle import loader as L
    return GpuTreeEngine(N, A, S), L.oracle_tree(N, A, S)


@pytest.mark.parametrize("name", golden_files("tree_"))
def test_cuda_matches_reference_golden(name):
    from gpu_adapters import GpuTreeEngine
    g = load_golden(name)
    d = golden_tree_inputs(g)
    eng = GpuTreeEngine(int(g["N"]), int(g["A"]), int(g["S"]), CONST["delta"])
    run_tree_lockstep(eng, g, d)
    assert (eng.trajectories(int(g["S"])) == g["traj"]).all()
    assert (eng.path_lens() == g["plen"][-1]).all()


def _lockstep_vs_oracle(N, A, S, seed, noise=True, mask_mode="random", mutate=None, every=1):
    gpu, cpu = _engines(N, A, S)
    d = tree_inputs(N, A, S, seed, mask_mode)
    if mutate:
        mutate(d)
    for e in (gpu, cpu):
        e.prepare(CONST["frac"], d["noise"] if noise else None, d["reward"], d["logits"], d["mask"])
    assert bits_equal(gpu.root_priors(), cpu.root_priors()), "root priors"
    for s in range(S - 1):
        a = gpu.traverse(CONST["pb_c_base"], CONST["pb_c_init"], CONST["discount"])
        b = cpu.traverse(CONST["pb_c_base"], CONST["pb_c_init"], CONST["discount"])
        for x, y, nm in zip(a, b, ("ix", "iy", "action")):
            assert (x == y).all(), f"{nm} differs at simulation {s}: trees {np.flatnonzero(x != y)[:8]}"
        for e in (gpu, cpu):
            e.backprop(s + 1, CONST["discount"], d["sim_reward"][s], d["sim_value"][s], d["sim_logits"][s])
        if s % every == 0 or s == S - 2:
            for x, y, nm in zip(gpu.stats(), cpu.stats(), ("visits", "values", "minmax")):
                if x.dtype == np.float32:
                    np.testing.assert_allclose(x, y, rtol=1e-5, atol=0, err_msg=f"{nm} at sim {s}")
                assert bits_equal(x, y), f"{nm} not bit-exact at simulation {s}"
    assert (gpu.path_lens() == cpu.path_lens()).all()
    r, vs, vc = gpu.expanded_stats_all(S)
    for i in range(0, N, max(1, N // 16)):
        ro, vso, vco = cpu.expanded_stats(i, S)
        n = len(ro)
        assert n == S - 1
        assert bits_equal(r[i, :n], ro) and bits_equal(vs[i, :n], vso) and (vc[i, :n] == vco).all()
    assert (gpu.trajectories(S) == cpu.trajectories(S)).all()
    return gpu, cpu


@pytest.mark.parametrize("N,A,S,seed", [(256, 20, 50, 1), (16, 11, 50, 2), (1, 20, 50, 3),
                                        (1000, 20, 50, 4), (37, 20, 200, 5), (130, 32, 20, 6),
                                        (64, 1, 10, 7), (5, 2, 70, 8)])
def test_lockstep_vs_oracle(N, A, S, seed):
    _lockstep_vs_oracle(N, A, S, seed, every=7)


def test_no_noise_and_zero_mask_rows():
    _lockstep_vs_oracle(48, 20, 40, 21, noise=False)
    _lockstep_vs_oracle(48, 20, 40, 22, mask_mode="zero_rows")


def test_nan_inf_and_extreme_inputs():
    def mutate(d):
        d["logits"][0, 3] = np.nan
        d["logits"][1, :] = -200.0
        d["logits"][1, 5] = 50.0
        d["logits"][2, :] = 1e30
        d["sim_value"][2, 3] = np.nan
        d["sim_reward"][4, 4] = np.inf
        d["sim_logits"][5, 6, :] = -1e4
        d["sim_logits"][5, 6, 2] = 88.0
        d["sim_value"][1:, 7] = 0.0      # constant values -> delta == 0 / below value_delta_max
        d["sim_reward"][1:, 7] = 0.0
        d["sim_value"][:, 8] = 1e-4 * d["sim_value"][:, 8]
        d["sim_reward"][:, 8] = 0.0
        d["mask"][5, :] = 1
    _lockstep_vs_oracle(12, 20, 24, 33, mutate=mutate)


def test_deep_chain_beyond_one_warp_of_path():
    """Logits/values that force a single line of play: depth > 32 exercises the chunked path walk."""
    def mutate(d):
        d["logits"][:] = -30.0
        d["logits"][:, 0] = 30.0
        d["sim_logits"][:] = -30.0
        d["sim_logits"][:, :, 0] = 30.0
        d["sim_value"][:] = 1.0 + 0.01 * np.arange(d["sim_value"].shape[0])[:, None]
        d["sim_reward"][:] = 0.0
        d["mask"][:] = 1
    gpu, cpu = _lockstep_vs_oracle(6, 20, 80, 44, mutate=mutate)
    assert cpu.path_lens().max() > 40


def test_full_size_config4_properties_and_oracle():
    """BASELINE configs[3] size: 4096 Hanabi-Full trees x 50 simulations, lock-step vs the oracle,
    plus size-independent invariants."""
    N, A, S = 4096, 20, 50
    gpu, cpu = _lockstep_vs_oracle(N, A, S, 77, every=16)
    visits, values, minmax = gpu.stats()
    assert (visits.sum(1) == S - 1).all()          # every simulation visits exactly one root child
    assert (minmax[:, 0] <= minmax[:, 1]).all()
    assert (gpu.path_lens() >= 2).all()


def test_full_size_config5_shard_properties_and_oracle():
    """BASELINE configs[4] per-GPU shard at 8 GPUs: 2048 Hanabi-Full trees x 200 simulations (4020 node slots per
    tree, deep paths), lock-step vs the oracle plus the same invariants."""
    N, A, S = 2048, 20, 200
    gpu, cpu = _lockstep_vs_oracle(N, A, S, 78, every=40)
    visits, values, minmax = gpu.stats()
    assert (visits.sum(1) == S - 1).all()
    assert (minmax[:, 0] <= minmax[:, 1]).all() and np.isfinite(values).all()
    assert gpu.path_lens().max() >= 8              # the deep-tree stress really is deep


def test_fused_backprop_traverse_equals_separate_calls():
    from hanabizero_b200 import _lib, cytree
    N, A, S = 300, 20, 30
    d = tree_inputs(N, A, S, 5)
    dev = torch.device("cuda")
    t = {k: torch.from_numpy(v).to(dev) for k, v in d.items()}
    outs = []
    for fused in (False, True):
        roots = cytree.Roots(N, A, S)
        mm = cytree.MinMaxStatsList(N)
        mm.set_delta(CONST["delta"])
        roots.prepare(CONST["frac"], t["noise"], t["reward"], t["logits"], t["mask"])
        lib, st, mmt = roots._lib, roots._stream(), mm.tensor(dev)
        ix = torch.empty(N, dtype=torch.int32, device=dev)
        la = torch.empty(N, dtype=torch.int32, device=dev)
        trace = []
        _lib.check(lib.hz_trees_traverse(roots.handle, st, CONST["pb_c_base"], CONST["pb_c_init"], CONST["discount"],
                                         mmt.data_ptr(), CONST["delta"], ix.data_ptr(), None, la.data_ptr(), None, None, None, 0))
        trace.append((ix.clone(), la.clone()))
        for s in range(S - 1):
            args = (t["sim_reward"][s].data_ptr(), t["sim_value"][s].data_ptr(), t["sim_logits"][s].data_ptr())
            last = s == S - 2
            if fused and not last:
                _lib.check(lib.hz_trees_backprop_traverse(
                    roots.handle, st, s + 1, CONST["discount"], *args, 0, mmt.data_ptr(), CONST["delta"],
                    CONST["pb_c_base"], CONST["pb_c_init"], ix.data_ptr(), None, la.data_ptr(), None, None, None, 0))
            else:
                _lib.check(lib.hz_trees_backprop(roots.handle, st, s + 1, CONST["discount"], *args, 0, mmt.data_ptr()))
                if not last:
                    _lib.check(lib.hz_trees_traverse(roots.handle, st, CONST["pb_c_base"], CONST["pb_c_init"], CONST["discount"],
                                                     mmt.data_ptr(), CONST["delta"], ix.data_ptr(), None, la.data_ptr(), None, None, None, 0))
            if not last:
                trace.append((ix.clone(), la.clone()))
        visits, values = roots.get_stats_tensors()
        outs.append((trace, visits.cpu(), values.cpu(), mmt.cpu().clone()))
    (tr0, v0, val0, mm0), (tr1, v1, val1, mm1) = outs
    for (a, b), (c, e) in zip(tr0, tr1):
        assert torch.equal(a, c) and torch.equal(b, e)
    assert torch.equal(v0, v1)
    assert bits_equal(val0.numpy(), val1.numpy()) and bits_equal(mm0.numpy(), mm1.numpy())


def test_fused_gather_matches_index_select():
    from hanabizero_b200 import cytree
    N, A, S, F = 200, 20, 12, 512
    d = tree_inputs(N, A, S, 9)
    dev = torch.device("cuda")
    roots = cytree.Roots(N, A, S)
    mm = cytree.MinMaxStatsList(N)
    mm.set_delta(CONST["delta"])
    roots.prepare(CONST["frac"], d["noise"], d["reward"], d["logits"], d["mask"])
    pool = torch.randn(S, N, F, device=dev)
    hidden = torch.empty(N, F, device=dev)
    act64 = torch.empty(N, 1, dtype=torch.int64, device=dev)
    for s in range(S - 1):
        res = cytree.ResultsWrapper(N)
        ix, iy, la = cytree.multi_traverse(roots, CONST["pb_c_base"], CONST["pb_c_init"], CONST["discount"],
                                           mm, res, as_tensor=True, pool=pool, out_hidden=hidden, out_action64=act64)
        assert torch.equal(hidden, pool[ix.long(), iy.long()])
        assert torch.equal(act64.view(-1), la.long())
        assert (ix <= s).all() and torch.equal(iy.cpu(), torch.arange(N, dtype=torch.int32))
        cytree.multi_back_propagate(s + 1, CONST["discount"], d["sim_reward"][s], d["sim_value"][s], d["sim_logits"][s], mm, res)


def test_list_api_like_reference_caller():
    """Lists in, lists out, exactly as core/mcts.py and core/selfplay_worker.py drive cytree."""
    from hanabizero_b200 import cytree
    from oracle import loader as L
    N, A, S = 16, 20, 20
    d = tree_inputs(N, A, S, 12)
    roots = cytree.Roots(N, A, S)
    legal = [row.astype(np.float64) for row in d["mask"]]           # selfplay passes float arrays
    roots.prepare(0.25, d["noise"].tolist(), d["reward"].tolist(), d["logits"].tolist(), legal)
    mm = cytree.MinMaxStatsList(N)
    mm.set_delta(0.006)
    cpu = L.oracle_tree(N, A, S)
    cpu.prepare(0.25, d["noise"], d["reward"], d["logits"], d["mask"])
    for s in range(S - 1):
        res = cytree.ResultsWrapper(N)
        ix, iy, la = cytree.multi_traverse(roots, 19652, 1.25, 0.999, mm, res)
        assert isinstance(ix, list) and isinstance(la[0], int)
        b = cpu.traverse(19652, 1.25, 0.999)
        assert ix == b[0].tolist() and iy == b[1].tolist() and la == b[2].tolist()
        cytree.multi_back_propagate(s + 1, 0.999, d["sim_reward"][s].tolist(), d["sim_value"][s].tolist(),
                                    d["sim_logits"][s].tolist(), mm, res)
        cpu.backprop(s + 1, 0.999, d["sim_reward"][s], d["sim_value"][s], d["sim_logits"][s])
    assert roots.get_distributions() == cpu.stats()[0].tolist()
    np.testing.assert_allclose(roots.get_values(), cpu.stats()[1], rtol=1e-5)
    roots.clear()


def test_errors_are_loud():
    from hanabizero_b200 import cytree
    from hanabizero_b200._lib import HzError
    roots = cytree.Roots(4, 20, 5)
    mm, res = cytree.MinMaxStatsList(4), cytree.ResultsWrapper(4)
    with pytest.raises(HzError):
        cytree.multi_traverse(roots, 19652, 1.25, 0.999, mm, res)      # not prepared
    roots.prepare_no_noise([0.0] * 4, np.zeros((4, 20)), np.ones((4, 20)))
    with pytest.raises(ValueError):
        roots.prepare_no_noise([0.0] * 3, np.zeros((4, 20)), np.ones((4, 20)))  # wrong size
    cytree.multi_traverse(roots, 19652, 1.25, 0.999, mm, res)
    with pytest.raises(HzError):
        cytree.multi_back_propagate(3, 0.999, [0.0] * 4, [0.0] * 4, np.zeros((4, 20)), mm, res)  # x out of order


def test_random_tie_break_generic_calls_match_oracle_on_degenerate_ties():
    """The reference's stock initialisation zeroes the output heads: uniform priors, zero values -> every select is
    an A-way tie.  Tie mode "random" must walk the same trees as the oracle running the same generator,
    and must not always descend action 0."""
    from gpu_adapters import GpuTreeEngine
    from oracle import loader as L
    N, A, S = 64, 20, 40
    d = tree_inputs(N, A, S, 3)
    for k in ("logits", "sim_logits", "sim_value", "sim_reward", "reward"):
        d[k][:] = 0.0
    d["mask"][:] = 1
    gpu, cpu = GpuTreeEngine(N, A, S), L.oracle_tree(N, A, S)
    gpu.roots.set_tie_break("random", 99, 7)
    cpu.set_tie(1, 99, 7)
    for e in (gpu, cpu):
        e.prepare(CONST["frac"], None, d["reward"], d["logits"], d["mask"])
    first_actions = []
    for s in range(S - 1):
        a = gpu.traverse(CONST["pb_c_base"], CONST["pb_c_init"], CONST["discount"])
        b = cpu.traverse(CONST["pb_c_base"], CONST["pb_c_init"], CONST["discount"])
        for x, y in zip(a, b):
            assert (x == y).all(), f"simulation {s}"
        first_actions.append(a[2].copy())
        for e in (gpu, cpu):
            e.backprop(s + 1, CONST["discount"], d["sim_reward"][s], d["sim_value"][s], d["sim_logits"][s])
    assert (gpu.stats()[0] == cpu.stats()[0]).all()
    assert len(np.unique(first_actions[0])) > A // 2        # the first pick is spread over the actions
    # a shard of the batch with tree_offset draws exactly what the full batch draws for those trees
    lo, hi = 24, 56
    shard = GpuTreeEngine(hi - lo, A, S)
    shard.roots.set_tie_break("random", 99, 7 + lo)
    shard.prepare(CONST["frac"], None, d["reward"][lo:hi], d["logits"][lo:hi], d["mask"][lo:hi])
    for s in range(S - 1):
        a = shard.traverse(CONST["pb_c_base"], CONST["pb_c_init"], CONST["discount"])
        assert (a[2] == first_actions[s][lo:hi]).all(), f"shard diverges at simulation {s}"
        shard.backprop(s + 1, CONST["discount"], d["sim_reward"][s][lo:hi], d["sim_value"][s][lo:hi], d["sim_logits"][s][lo:hi])
    assert (shard.stats()[0] == gpu.stats()[0][lo:hi]).all()
    # default mode: element 0 of the tie list, i.e. action 0 at the first select of every tree
    gpu0 = GpuTreeEngine(N, A, S)
    gpu0.prepare(CONST["frac"], None, d["reward"], d["logits"], d["mask"])
    assert (gpu0.traverse(CONST["pb_c_base"], CONST["pb_c_init"], CONST["discount"])[2] == 0).all()
