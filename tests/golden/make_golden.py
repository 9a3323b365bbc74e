"""Generates the golden fixtures in this directory FROM THE REFERENCE ITSELF.  Run in the build
container only (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

* env_*.npz   — episodes played through the reference's own Python environment
  (/root/reference/envs/hanabi/rl_env.py HanabiEnv over pyhanabi.py + cffi + libpyhanabi.so, the
  latter compiled unmodified into oracle/_ref).  Two import shims only: numpy.int (removed in
  numpy >= 1.24, used at rl_env.py:256,429) and a stub gym.spaces.Discrete (rl_env.py:21).
* tree_*.npz  — lock-step MCTS traces from the reference tree engine compiled with the rand()==0
  shim (oracle/_ref/libref_ctree_det.so), inputs included.

Nothing here runs on the GPU box; the fixtures are what travels.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONST = dict(pb_c_base=19652, pb_c_init=1.25, discount=0.999, delta=0.006, frac=0.25)


def import_reference_env():
    np.int = int  # rl_env.py:256,429
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")

    class Discrete:  # rl_env.py:21,141
        def __init__(self, n):
            self.n = n

    spaces.Discrete = Discrete
    gym.spaces = spaces
    sys.modules["gym"] = gym
    sys.modules["gym.spaces"] = spaces
    os.chdir(os.path.join(ROOT, "oracle", "_ref"))  # pyhanabi.try_load looks in "." for libpyhanabi.so
    sys.path.insert(0, "/root/reference")
    pkg = types.ModuleType("envs")  # skip envs/__init__.py's absl FLAGS side effect
    pkg.__path__ = ["/root/reference/envs"]
    sys.modules["envs"] = pkg
    from envs.hanabi.rl_env import HanabiEnv
    from envs.hanabi import pyhanabi
    assert pyhanabi.lib_loaded_flag
    return HanabiEnv


def policy(mode, rng, hand_size, legal, playable_slots):
    """Shared with tests/test_oracle_hanabi.py: 'noplay' avoids play moves (long games, deck runs
    out), 'smart' plays playable cards (fireworks complete, perfect scores), 'random' = uniform."""
    ids = np.flatnonzero(legal)
    h = hand_size
    nonplay = [a for a in ids if not (h <= a < 2 * h)]
    if mode == "random":
        return int(ids[rng.integers(len(ids))])
    if mode == "smart" and playable_slots and rng.random() < 0.9:
        return h + int(playable_slots[rng.integers(len(playable_slots))])
    if nonplay:
        return int(nonplay[rng.integers(len(nonplay))])
    return int(ids[rng.integers(len(ids))])


def playable_from_env(env):
    st = env.state
    cur = st.cur_player()
    fw = st.fireworks()
    hand = st.player_hands()[cur]
    return [k for k, c in enumerate(hand) if c.rank() == fw[c.color()]]


def make_env(HanabiEnv, name, seed, episodes, mode, tag):
    env = HanabiEnv({"hanabi_name": name, "seed": seed})
    rng = np.random.default_rng(1000 * seed + episodes)
    h = env.game.hand_size()
    rec = dict(glob=[], loc=[], legal=[], reward=[], done=[], score=[], action=[], ep_start=[])
    for ep in range(episodes):
        share_obs, obs, legal = env.reset()
        rec["ep_start"].append(len(rec["glob"]))
        rec["glob"].append(share_obs); rec["loc"].append(obs); rec["legal"].append(legal)
        rec["reward"].append(0); rec["done"].append(0); rec["score"].append(0); rec["action"].append(-1)
        done = False
        while not done:
            a = policy(mode, rng, h, np.asarray(legal), playable_from_env(env))
            share_obs, obs, reward, done, info, legal = env.step(a)
            rec["glob"].append(share_obs); rec["loc"].append(obs); rec["legal"].append(legal)
            rec["reward"].append(int(reward)); rec["done"].append(int(done))
            rec["score"].append(int(info["score"])); rec["action"].append(a)
    g = np.asarray(rec["glob"], np.uint8)
    l = np.asarray(rec["loc"], np.uint8)
    path = os.path.join(OUT, f"env_{tag}.npz")
    np.savez_compressed(
        path, preset=0 if name == "Hanabi-Full" else 1, seed=seed, mode=mode,
        glob_bits=np.packbits(g, axis=1), glob_dim=g.shape[1],
        loc_bits=np.packbits(l, axis=1), loc_dim=l.shape[1],
        legal=np.asarray(rec["legal"], np.uint8), reward=np.asarray(rec["reward"], np.int32),
        done=np.asarray(rec["done"], np.uint8), score=np.asarray(rec["score"], np.int32),
        action=np.asarray(rec["action"], np.int32), ep_start=np.asarray(rec["ep_start"], np.int32))
    print(path, "steps", len(rec["action"]), "max score", max(rec["score"]))


def tree_inputs(N, A, S, seed, mask_mode):
    rng = np.random.default_rng(seed)
    d = dict(
        logits=rng.standard_normal((N, A)).astype(np.float32),
        noise=rng.dirichlet([0.3] * A, N).astype(np.float32),
        reward=rng.standard_normal(N).astype(np.float32),
        sim_reward=rng.standard_normal((S - 1, N)).astype(np.float32),
        sim_value=rng.standard_normal((S - 1, N)).astype(np.float32),
        sim_logits=rng.standard_normal((S - 1, N, A)).astype(np.float32))
    mask = (rng.random((N, A)) < 0.6).astype(np.int32)
    mask[np.arange(N), rng.integers(0, A, N)] = 1
    if mask_mode == "zero_rows":  # reanalyze: out-of-trajectory roots carry all-zero masks
        mask[::3] = 0
        d["noise"] = d["noise"] * mask  # core/reanalyze_worker.py:344
    d["mask"] = mask
    return d


def make_tree(N, A, S, seed, tag, noise=True, mask_mode="random", store_inputs=True):
    from oracle import loader as L
    d = tree_inputs(N, A, S, seed, mask_mode)
    r = L.ref_tree(N, A, S, CONST["delta"])
    r.prepare(CONST["frac"], d["noise"] if noise else None, d["reward"], d["logits"], d["mask"])
    priors = r.root_priors()
    ix, iy, la, plen = [], [], [], []
    for s in range(S - 1):
        a = r.traverse(CONST["pb_c_base"], CONST["pb_c_init"], CONST["discount"])
        ix.append(a[0]); iy.append(a[1]); la.append(a[2]); plen.append(r.path_lens())
        r.backprop(s + 1, CONST["discount"], d["sim_reward"][s], d["sim_value"][s], d["sim_logits"][s])
    visits, values, minmax = r.stats()
    out = dict(N=N, A=A, S=S, seed=seed, noise=int(noise), mask_mode=mask_mode, priors=priors,
               ix=np.asarray(ix, np.int16), iy=np.asarray(iy, np.int32), la=np.asarray(la, np.int8),
               plen=np.asarray(plen, np.int16), visits=visits, values=values, minmax=minmax,
               traj=r.trajectories(S))
    if store_inputs:
        out.update({"in_" + k: v for k, v in d.items()})
    path = os.path.join(OUT, f"tree_{tag}.npz")
    np.savez_compressed(path, **out)
    print(path, "root0", visits[0].tolist(), repr(float(values[0])))


if __name__ == "__main__":
    make_tree(8, 20, 50, 1234, "full_n8_s50")
    make_tree(16, 11, 50, 7, "small_n16_s50")
    make_tree(8, 20, 50, 99, "full_n8_s50_nonoise", noise=False)
    make_tree(12, 20, 30, 31, "full_n12_s30_zero_rows", mask_mode="zero_rows")
    make_tree(64, 20, 200, 5, "full_n64_s200", store_inputs=False)
    HanabiEnv = import_reference_env()
    for seed in (0, 1):
        make_env(HanabiEnv, "Hanabi-Full", seed, 2, "noplay", f"full_seed{seed}_noplay")
        make_env(HanabiEnv, "Hanabi-Small", seed, 3, "noplay", f"small_seed{seed}_noplay")
    make_env(HanabiEnv, "Hanabi-Full", 2, 3, "smart", "full_seed2_smart")
    make_env(HanabiEnv, "Hanabi-Full", 3, 4, "random", "full_seed3_random")
    make_env(HanabiEnv, "Hanabi-Small", 2, 4, "smart", "small_seed2_smart")
    make_env(HanabiEnv, "Hanabi-Small", 3, 6, "random", "small_seed3_random")
