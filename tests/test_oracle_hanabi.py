"""CPU: the plain-C Hanabi oracle (oracle/hanabi_oracle.c) against golden episodes produced by the
reference's own Python HanabiEnv, and live against the compiled reference library."""
import numpy as np
import pytest

from helpers import env_policy, golden_files, load_golden, playable_from_dump, unpack_env_golden
from oracle import loader as L


@pytest.mark.parametrize("name", golden_files("env_"))
def test_oracle_replays_reference_python_env(name):
    g = load_golden(name)
    glob_, loc = unpack_env_golden(g)
    env = L.oracle_hanabi(int(g["preset"]), int(g["seed"]))
    assert env.global_dim == int(g["glob_dim"]) and env.local_dim == int(g["loc_dim"])
    starts = set(int(s) for s in g["ep_start"])
    for t in range(len(g["action"])):
        if t in starts:
            go, lo, legal = env.reset()
            r, d, sc = 0, False, 0
        else:
            go, lo, legal, r, d, sc = env.step(int(g["action"][t]))
        assert (go == glob_[t]).all(), f"global obs differs at step {t}"
        assert (lo == loc[t]).all(), f"local obs differs at step {t}"
        assert (legal == g["legal"][t]).all(), f"legal mask differs at step {t}"
        assert (r, int(d), sc) == (int(g["reward"][t]), int(g["done"][t]), int(g["score"][t])), t


def test_shapes():
    full, small = L.oracle_hanabi(0, 0), L.oracle_hanabi(1, 0)
    assert (full.enc_len, full.own_len, full.actions, full.local_dim, full.global_dim) == (658, 125, 20, 660, 785)
    assert (small.enc_len, small.own_len, small.actions, small.local_dim, small.global_dim) == (171, 20, 11, 173, 193)


def test_illegal_action_rejected():
    env = L.oracle_hanabi(0, 0)
    _, _, legal = env.reset()
    bad = int(np.flatnonzero(legal == 0)[0])
    with pytest.raises(ValueError):
        env.step(bad)


@pytest.mark.skipif(not L.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("preset", [0, 1])
@pytest.mark.parametrize("mode", ["random", "noplay", "smart"])
def test_oracle_matches_live_reference(preset, mode):
    steps = 0
    for seed in range(6):
        rng = np.random.default_rng(17 * seed + preset)
        o, r = L.oracle_hanabi(preset, seed), L.ref_hanabi(preset, seed)
        for ep in range(4):  # same game object: the mt19937 stream persists across resets
            a, b = o.reset(), r.reset()
            for x, y in zip(a, b):
                assert (x == y).all()
            done = False
            while not done:
                dump = o.dump()
                assert (dump == r.dump()).all()
                act = env_policy(mode, rng, o.hand_size, a[2],
                                 playable_from_dump(dump, o.colors, o.ranks, o.hand_size))
                a, b = o.step(act), r.step(act)
                for x, y in zip(a[:3], b[:3]):
                    assert (x == y).all(), (seed, ep, steps)
                assert a[3:] == b[3:]
                done = a[4]
                steps += 1
    assert steps > 100
