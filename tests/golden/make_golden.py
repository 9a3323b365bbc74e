"""Generates the golden fixtures in this directory FROM THE REFERENCE ITSELF.  Run in the build
container only (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

* env_*.npz   — episodes played through the reference's own Python environment
  (/root/reference/envs/hanabi/rl_env.py HanabiEnv over pyhanabi.py + cffi + libpyhanabi.so, the
  latter compiled unmodified into oracle/_ref).  Two import shims only: numpy.int (removed in
  numpy >= 1.24, used at rl_env.py:256,429) and a stub gym.spaces.Discrete (rl_env.py:21).
* tree_*.npz  — lock-step MCTS traces from the reference tree engine compiled with the rand()==0
  shim (oracle/_ref/libref_ctree_det.so), inputs included.
* traj_*.npz  — self-play bookkeeping by the reference's own Python: core/game.py GameHistory
  (init / store_search_stats / append / game_over), core/utils.py select_action and the turn-reward
  loop of core/selfplay_worker.py DataWorker.put, driven over episodes of the reference env with
  synthetic search results.  Import stubs only for modules that are not on this path (ray.put/get
  as identity, cv2, gym base classes, the compiled cytree).

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


def import_reference_selfplay():
    """core.game.GameHistory, core.utils.select_action and DataWorker.put from /root/reference, unmodified."""
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, n):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    stub("ray", put=lambda x: x, get=lambda x: x,
         remote=lambda *a, **k: (a[0] if a and callable(a[0]) and not k else (lambda c: c)))
    stub("cv2", ocl=_Any(), setNumThreads=lambda n: None)
    gym = sys.modules["gym"]
    for base in ("Wrapper", "ObservationWrapper", "RewardWrapper", "Env"):
        setattr(gym, base, object)
    gym.spaces.Box = _Any
    stub("core.mcts", MCTS=None)
    stub("core.model", concat_output=None, concat_output_value=None)
    import core
    ctree = stub("core.ctree")
    ctree.__path__ = []
    ctree.cytree = stub("core.ctree.cytree", Node=None, Roots=None)
    core.ctree = ctree
    from core.game import GameHistory
    from core.utils import select_action
    try:
        from core.selfplay_worker import DataWorker
        put = DataWorker.put
    except Exception as e:  # the module pulls in the whole training stack; the loop is restated if it cannot load
        print("core.selfplay_worker not importable here (%s: %s); using the restated turn-reward loop" % (type(e).__name__, e))

        def put(self, data):  # core/selfplay_worker.py:29-39
            game_histories, _ = data
            prev_r = game_histories.rewards[0]
            for step_id in range(1, len(game_histories.rewards)):
                cur_r = game_histories.rewards[step_id] + prev_r
                prev_r = game_histories.rewards[step_id]
                game_histories.rewards[step_id] = cur_r
            self.trajectory_pool.append(data)
    return GameHistory, select_action, put


def make_traj(HanabiEnv, ref, name, seed, episodes, mode, tag, stack=4):
    GameHistory, select_action, put = ref
    env = HanabiEnv({"hanabi_name": name, "seed": seed})
    A, h = env.num_moves(), env.game.hand_size()
    rng = np.random.default_rng(77 * seed + episodes)
    cfg = types.SimpleNamespace(stacked_observations=stack, discount=0.999, action_space_size=A)
    worker = types.SimpleNamespace(trajectory_pool=[])
    rec = dict(obs=[], legal=[], visits=[], value=[], action=[], reward=[], done=[])
    eps = []
    for ep in range(episodes):
        share_obs, _, legal = env.reset()
        obs, legal = np.array(share_obs), np.array(legal)
        rec["obs"].append(obs); rec["legal"].append(legal)          # the observation / mask an episode starts from
        rec["visits"].append(np.zeros(A, np.int32)); rec["value"].append(0.0); rec["action"].append(-1)
        rec["reward"].append(0); rec["done"].append(0)
        gh = GameHistory(types.SimpleNamespace(n=A), max_length=1000, config=cfg)
        gh.init([obs for _ in range(stack)], legal)
        done = False
        while not done:
            # synthetic search result: visit counts favour the scripted policy's move, some land on illegal moves
            want = policy(mode, rng, h, legal, playable_from_env(env))
            visits = rng.integers(0, 12, A).astype(np.int32)
            visits[want] = 40
            value = float(np.float32(rng.standard_normal()))
            dist = visits.tolist()
            action, _ = select_action(dist, temperature=1, deterministic=True, legal_actions=legal)
            share_obs, _, reward, done, info, legal_next = env.step(int(action))
            obs, legal_next = np.array(share_obs), np.array(legal_next)
            gh.store_search_stats(dist, value)                      # dist was mutated by select_action, as in the worker
            gh.append(action, obs, reward, legal_next)
            rec["obs"].append(obs); rec["legal"].append(legal_next); rec["visits"].append(visits)
            rec["value"].append(value); rec["action"].append(int(action)); rec["reward"].append(int(reward))
            rec["done"].append(int(done))
            legal = legal_next
        gh.game_over()
        put(worker, (gh, None))
        eps.append(gh)
    o = np.asarray(rec["obs"], np.uint8)
    out = dict(preset=0 if name == "Hanabi-Full" else 1, seed=seed, mode=mode, stack=stack, actions=A,
               obs_bits=np.packbits(o, axis=1), obs_dim=o.shape[1], legal=np.asarray(rec["legal"], np.uint8),
               visits=np.asarray(rec["visits"], np.int32), value=np.asarray(rec["value"], np.float32),
               action=np.asarray(rec["action"], np.int32), reward=np.asarray(rec["reward"], np.int32),
               done=np.asarray(rec["done"], np.uint8), ep_len=np.asarray([len(g.actions) for g in eps], np.int32),
               out_vis=np.concatenate([np.asarray(g.child_visits, np.float64) for g in eps]),
               out_root=np.concatenate([np.asarray(g.root_values, np.float64) for g in eps]),
               out_a=np.concatenate([np.asarray(g.actions, np.int64) for g in eps]),
               out_r=np.concatenate([np.asarray(g.rewards, np.int64) for g in eps]),
               out_o_bits=np.packbits(np.concatenate([np.asarray(g.obs_history, np.uint8) for g in eps]), axis=1),
               out_la=np.concatenate([np.asarray(g.legal_actions, np.float64) for g in eps]))
    path = os.path.join(OUT, f"traj_{tag}.npz")
    np.savez_compressed(path, **out)
    print(path, "episodes", episodes, "moves", out["ep_len"].tolist())


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
    ref = import_reference_selfplay()
    make_traj(HanabiEnv, ref, "Hanabi-Small", 4, 5, "random", "small_seed4")
    make_traj(HanabiEnv, ref, "Hanabi-Full", 5, 2, "smart", "full_seed5")
