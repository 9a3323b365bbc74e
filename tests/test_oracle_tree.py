"""CPU: the plain-C tree oracle (oracle/tree_oracle.c) against golden traces produced by the
unmodified reference ctree (oracle/_ref, rand()==0 shim) and, when present, live against it."""
import numpy as np
import pytest

from helpers import (CONST, bits_equal, golden_files, golden_tree_inputs, load_golden,
                     run_tree_lockstep, tree_inputs)
from oracle import loader as L


@pytest.mark.parametrize("name", golden_files("tree_"))
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    d = golden_tree_inputs(g)
    eng = L.oracle_tree(int(g["N"]), int(g["A"]), int(g["S"]), CONST["delta"])
    run_tree_lockstep(eng, g, d)
    assert (eng.trajectories(int(g["S"])) == g["traj"]).all()
    assert (eng.path_lens() == g["plen"][-1]).all()


def test_input_generator_in_sync_with_golden():
    g = load_golden("tree_full_n8_s50.npz")
    d = tree_inputs(8, 20, 50, 1234)
    for k in d:
        assert bits_equal(d[k], g["in_" + k]), k


@pytest.mark.skipif(not L.have_ref(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("N,A,S,seed,noise,mask_mode", [
    (32, 20, 50, 11, True, "random"), (16, 11, 50, 12, True, "random"),
    (8, 20, 120, 13, False, "random"), (9, 20, 40, 14, True, "zero_rows"), (1, 20, 50, 15, True, "random")])
def test_oracle_matches_live_reference(N, A, S, seed, noise, mask_mode):
    d = tree_inputs(N, A, S, seed, mask_mode)
    o, r = L.oracle_tree(N, A, S), L.ref_tree(N, A, S)
    for e in (o, r):
        e.prepare(CONST["frac"], d["noise"] if noise else None, d["reward"], d["logits"], d["mask"])
    assert bits_equal(o.root_priors(), r.root_priors())
    for s in range(S - 1):
        a = o.traverse(CONST["pb_c_base"], CONST["pb_c_init"], CONST["discount"])
        b = r.traverse(CONST["pb_c_base"], CONST["pb_c_init"], CONST["discount"])
        for x, y in zip(a, b):
            assert (x == y).all(), s
        for e in (o, r):
            e.backprop(s + 1, CONST["discount"], d["sim_reward"][s], d["sim_value"][s], d["sim_logits"][s])
        for x, y in zip(o.stats(), r.stats()):
            assert bits_equal(x, y), s
    for i in range(N):
        for x, y in zip(o.expanded_stats(i, S + 1), r.expanded_stats(i, S + 1)):
            assert bits_equal(x, y)


@pytest.mark.skipif(not L.have_ref(), reason="oracle/_ref not built")
def test_nan_and_extreme_inputs_match_reference():
    """NaN logits at the root (not sanitised by the reference), huge/tiny logits, NaN values."""
    N, A, S = 6, 20, 12
    d = tree_inputs(N, A, S, 3)
    d["logits"][0, 3] = np.nan
    d["logits"][1, :] = -200.0
    d["logits"][1, 5] = 50.0
    d["logits"][2, :] = 1e30
    d["sim_value"][2, 3] = np.nan
    d["sim_reward"][4, 4] = np.inf
    d["mask"][5, :] = 1
    o, r = L.oracle_tree(N, A, S), L.ref_tree(N, A, S)
    for e in (o, r):
        e.prepare(CONST["frac"], d["noise"], d["reward"], d["logits"], d["mask"])
    assert bits_equal(o.root_priors(), r.root_priors())
    for s in range(S - 1):
        a = o.traverse(CONST["pb_c_base"], CONST["pb_c_init"], CONST["discount"])
        b = r.traverse(CONST["pb_c_base"], CONST["pb_c_init"], CONST["discount"])
        for x, y in zip(a, b):
            assert (x == y).all(), s
        for e in (o, r):
            e.backprop(s + 1, CONST["discount"], d["sim_reward"][s], d["sim_value"][s], d["sim_logits"][s])
    for x, y in zip(o.stats(), r.stats()):
        assert bits_equal(x, y)


def test_random_tie_mode_is_uniform_like_the_stock_reference():
    """Tie mode 1 of the oracle (the counter-based draw the CUDA path uses) against the reference's stock build
    (rand() % ties, clock-seeded): with the reference's zero-initialised heads every select is an A-way tie, so the
    first action of a search is uniform over the actions for both — and never 'always action 0'."""
    from helpers import tree_inputs
    N, A, S = 4000, 20, 3
    d = tree_inputs(N, A, S, 5)
    for k in ("logits", "sim_logits", "sim_value", "sim_reward", "reward"):
        d[k][:] = 0.0
    d["mask"][:] = 1
    counts = {}
    engines = {"oracle": L.oracle_tree(N, A, S)}
    engines["oracle"].set_tie(1, 2024, 0)
    if L.have_ref():
        engines["reference"] = L.ref_tree(N, A, S, deterministic=False)
    for name, e in engines.items():
        e.prepare(0.25, None, d["reward"], d["logits"], d["mask"])
        la = e.traverse(19652, 1.25, 0.999)[2]
        counts[name] = np.bincount(la, minlength=A)
    for name, c in counts.items():
        chi2 = ((c - N / A) ** 2 / (N / A)).sum()
        assert chi2 < 60.0, f"{name}: first actions not uniform over {A}-way ties (chi2 {chi2:.1f}, 19 dof): {c}"
