"""GPU: the "next" rows N3 / N4 — device trajectory recording against a line-by-line Python
port of the reference's GameHistory bookkeeping, and the reanalyze caller of the search."""
import numpy as np
import pytest
import torch

from helpers import RefGameHistory, golden_files, load_golden, ref_select_action, unpack_traj_golden

pytestmark = pytest.mark.gpu


def _same(ep, ref):
    assert (ep["o"] == ref.obs_history).all() and ep["o"].shape == ref.obs_history.shape
    assert (ep["a"] == ref.actions).all() and (ep["r"] == ref.rewards).all()
    assert ep["vis"].dtype == np.float64 and (ep["vis"] == ref.child_visits).all()      # bit-exact double quotients
    assert (ep["root"] == ref.root_values).all()
    assert (ep["la"] == ref.legal_actions).all() and ep["la"].shape == ref.legal_actions.shape


def test_recorder_kernels_match_game_history_port():
    """Synthetic moves with ragged episode lengths, strided observation rows, an inactive-game mask and up to two
    episodes of a game finishing between flushes (three banks in use)."""
    from hanabizero_b200.trajectory import TrajectoryRecorder
    rng = np.random.default_rng(0)
    N, D, A, stack, L = 37, 173, 11, 4, 24
    rec = TrajectoryRecorder(N, D, A, stack, max_len=L, banks=3)
    refs = [RefGameHistory(stack) for _ in range(N)]
    done_eps = [[] for _ in range(N)]
    cuda = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()

    def fresh():
        obs = (rng.random((N, D + 5)) < 0.4).astype(np.float32)        # row stride D + 5
        legal = (rng.random((N, A)) < 0.6).astype(np.float32)
        return obs, legal

    obs, legal = fresh()
    rec.begin(cuda(obs)[:, :D], cuda(legal))
    for i in range(N):
        refs[i].init([obs[i, :D]] * stack, legal[i])
    ep_len = rng.integers(1, 9, N)
    got = [[] for _ in range(N)]
    for t in range(60):
        active = (rng.random(N) < 0.8).astype(np.uint8)
        acts = rng.integers(0, A, N).astype(np.int32)
        rew = rng.integers(-1, 3, N).astype(np.int32)
        vis = rng.integers(0, 30, (N, A)).astype(np.int32)
        vis[np.arange(N), rng.integers(0, A, N)] += 1
        root = rng.standard_normal(N).astype(np.float32)
        obs, legal = fresh()
        done = np.zeros(N, np.uint8)
        for i in range(N):
            if not active[i]:
                continue
            refs[i].store_search_stats(vis[i].tolist(), float(root[i]))
            refs[i].append(int(acts[i]), obs[i, :D], int(rew[i]), legal[i])
            if len(refs[i].actions) >= ep_len[i]:
                done[i] = 1
        rec.append(cuda(acts), cuda(obs)[:, :D], cuda(legal), cuda(rew), cuda(vis), cuda(root), cuda(done), cuda(active))
        if done.any():
            obs2, legal2 = fresh()
            rec.begin(cuda(obs2)[:, :D], cuda(legal2), mask=cuda(done))
            for i in np.flatnonzero(done):
                refs[i].game_over(); refs[i].put()
                done_eps[i].append(refs[i])
                refs[i] = RefGameHistory(stack)
                refs[i].init([obs2[i, :D]] * stack, legal2[i])
                ep_len[i] = rng.integers(1, 9)
        if t % 2 == 1:     # episodes last >= 1 move: at most two of a game finish between flushes, the third bank records
            for ep in rec.flush():
                got[ep["game"]].append(ep)
    for ep in rec.flush():
        got[ep["game"]].append(ep)
    assert sum(len(g) for g in got) == sum(len(d) for d in done_eps) > N
    for i in range(N):
        assert len(got[i]) == len(done_eps[i])
        for ep, ref in zip(got[i], done_eps[i]):
            _same(ep, ref)


def test_recorder_reports_overflow():
    from hanabizero_b200.trajectory import TrajectoryOverflow, TrajectoryRecorder
    N, D, A = 5, 16, 4
    rec = TrajectoryRecorder(N, D, A, stack=2, max_len=3)
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device="cuda")
    rec.begin(z(N, D), z(N, A))
    for _ in range(3):
        rec.append(z(N, dt=torch.int32), z(N, D), z(N, A), z(N, dt=torch.int32), torch.ones(N, A, dtype=torch.int32, device="cuda"),
                   z(N), z(N, dt=torch.uint8))
    rec.check()
    rec.append(z(N, dt=torch.int32), z(N, D), z(N, A), z(N, dt=torch.int32), torch.ones(N, A, dtype=torch.int32, device="cuda"),
               z(N), z(N, dt=torch.uint8))
    with pytest.raises(TrajectoryOverflow):
        rec.check()
    with pytest.raises(TypeError):
        rec.begin(z(N, D, dt=torch.uint8), z(N, A))


@pytest.mark.parametrize("mdp", ["global", "local"])
def test_selfplay_engine_records_reference_trajectories(mdp):
    """SelfPlayEngine(record=True) on Hanabi-Small: the per-move outputs are replayed on the host through the
    GameHistory port; every flushed episode must equal the port's record, terminal observation included."""
    from hanabizero_b200.mcts import SearchConfig
    from hanabizero_b200.model import MuZeroNet
    from hanabizero_b200.selfplay import SelfPlayEngine
    N, S, stack = 48, 6, 4
    torch.manual_seed(0)
    dim = 193 if mdp == "global" else 173
    model = MuZeroNet(dim * stack, 11).randomize_heads().cuda().eval()
    eng = SelfPlayEngine(N, "Hanabi-Small", model, SearchConfig(num_simulations=S), seeds=np.arange(N) + 5, mdp=mdp,
                         stack=stack, record=True, max_episode_len=64, record_banks=4)
    eng.reset()
    first, legal0 = eng.frames[:, -1].cpu().numpy(), eng.legal.cpu().numpy()
    refs = [RefGameHistory(stack) for _ in range(N)]
    for i in range(N):
        refs[i].init([first[i]] * stack, legal0[i].astype(np.float64))
    want, got = [[] for _ in range(N)], [[] for _ in range(N)]
    for t in range(70):
        out = eng.step(temperature=1.0, deterministic=False)
        h = {k: v.cpu().numpy() for k, v in out.items()}
        nxt, nxt_legal = eng.frames[:, -1].cpu().numpy(), eng.legal.cpu().numpy()
        for i in range(N):
            refs[i].store_search_stats(h["visits"][i].tolist(), float(h["root_value"][i]))
            refs[i].append(int(h["action"][i]), h["obs"][i], int(h["reward"][i]), h["legal"][i].astype(np.float64))
            if h["done"][i]:
                refs[i].game_over(); refs[i].put()
                want[i].append(refs[i])
                refs[i] = RefGameHistory(stack)
                refs[i].init([nxt[i]] * stack, nxt_legal[i].astype(np.float64))
        if t % 3 == 2:     # a Hanabi-Small game can end on its first move: up to 3 episodes pending + 1 recording
            for ep in eng.recorder.flush():
                got[ep["game"]].append(ep)
    for ep in eng.recorder.flush():
        got[ep["game"]].append(ep)
    eng.env.check()
    assert sum(len(w) for w in want) >= N        # Hanabi-Small episodes are short: every game finished at least once
    for i in range(N):
        assert len(got[i]) == len(want[i]), i
        for ep, ref in zip(got[i], want[i]):
            _same(ep, ref)
            assert len(ep["a"]) + stack == len(ep["o"]) and len(ep["la"]) == len(ep["a"]) + 1


def test_visit_policy_is_pythons_int_division():
    from hanabizero_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(1)
    N, A = 300, 20
    vis = rng.integers(0, 200, (N, A)).astype(np.int32)
    vis[np.arange(N), rng.integers(0, A, N)] += 1
    mask = (rng.random(N) < 0.7).astype(np.uint8)
    d_vis, d_mask = torch.from_numpy(vis).cuda(), torch.from_numpy(mask).cuda()
    o64 = torch.empty(N, A, dtype=torch.float64, device="cuda")
    o32 = torch.empty(N, A, dtype=torch.float32, device="cuda")
    _lib.check(lib.hz_visit_policy(torch.cuda.current_stream().cuda_stream, d_vis.data_ptr(), d_mask.data_ptr(), N, A,
                                   o64.data_ptr(), o32.data_ptr()))
    want = np.array([[v / sum(row) if m else 0.0 for v in row] for row, m in zip(vis.tolist(), mask)])
    assert (o64.cpu().numpy() == want).all()
    assert (o32.cpu().numpy() == want.astype(np.float32)).all()


def test_reanalyze_policies_equal_a_direct_search():
    """core/reanalyze_worker.py:339-367: masked noise, all-zero legal rows for padded positions, visit counts ->
    targets, zero rows where policy_mask == 0.  Checked against the same search driven by hand through cytree."""
    from hanabizero_b200 import cytree
    from hanabizero_b200.hanabi_env import HanabiVecEnv
    from hanabizero_b200.mcts import MCTS, SearchConfig
    from hanabizero_b200.model import MuZeroNetFull
    from hanabizero_b200.reanalyze import reanalyze_policies
    batch, unroll, A, S, stack = 12, 5, 20, 10, 4
    B = batch * (unroll + 1)
    torch.manual_seed(0)
    env = HanabiVecEnv(B, "Hanabi-Full", np.arange(B))
    g, _, legal = env.reset_all()
    obs = g.repeat(1, stack)
    rng = np.random.default_rng(2)
    mask = (rng.random(B) < 0.75).astype(np.uint8)
    legal = legal.clone()
    legal[torch.from_numpy(mask == 0).cuda()] = 0          # padded positions carry all-zero legal masks
    noise = rng.dirichlet([0.3] * A, B).astype(np.float32)
    model = MuZeroNetFull(785 * stack, A).randomize_heads().cuda().eval()
    cfg = SearchConfig(num_simulations=S)
    cfg.action_space_size = A
    got = reanalyze_policies(cfg, model, obs.cpu().numpy(), legal.cpu().numpy(), mask, unroll, noises=noise)
    assert got.shape == (batch, unroll + 1, A) and got.dtype == np.float64
    _, logits, hidden = model.initial_inference_device(obs)
    roots = cytree.Roots(B, A, S)
    roots.prepare(cfg.root_exploration_fraction, (noise * legal.cpu().numpy()).tolist(), [0.0] * B, logits.tolist(),
                  [row.astype(int) for row in legal.cpu().numpy()])
    MCTS(cfg).run_multi(roots, model, hidden)
    dist = roots.get_distributions()
    want = np.array([[v / sum(d) for v in d] if m else [0.0] * A for d, m in zip(dist, mask)]).reshape(batch, unroll + 1, A)
    assert (got == want).all()
    assert all(sum(d) == S - 1 for d in dist)
    draws = reanalyze_policies(cfg, model, obs, legal, mask, unroll, as_tensor=True)      # device noise draw
    assert draws.is_cuda and torch.allclose(draws.sum(-1).view(-1), torch.from_numpy(mask.astype(np.float64)).cuda())


@pytest.mark.parametrize("name", golden_files("traj_"))
def test_recorder_and_select_action_replay_reference_trajectories(name):
    """Fixtures written by the reference's own GameHistory / select_action / DataWorker.put (tests/golden/make_golden.py):
    the same per-move search results go through hz_select_action and the recorder kernels; actions, mutated visit
    distributions, observations, turn rewards, root values and legal masks must come out as the reference stored them."""
    from hanabizero_b200.selfplay import select_action_batch
    from hanabizero_b200.trajectory import TrajectoryRecorder
    g = load_golden(name)
    obs, eps = unpack_traj_golden(g)
    A, stack, D = int(g["actions"]), int(g["stack"]), int(g["obs_dim"])
    rec = TrajectoryRecorder(1, D, A, stack, max_len=int(g["ep_len"].max()) + 1)
    cuda = lambda x, dt: torch.as_tensor(np.ascontiguousarray(x), dtype=dt).cuda().view(1, -1)
    got = []
    for t in range(len(g["action"])):
        o_t, la_t = cuda(obs[t], torch.float32), cuda(g["legal"][t], torch.float32)
        if g["action"][t] < 0:                      # an episode starts here
            rec.begin(o_t, la_t)
            prev_legal = la_t
            continue
        visits = cuda(g["visits"][t], torch.int32)
        action, _ = select_action_batch(visits, prev_legal, 1.0, deterministic=True)
        assert int(action.item()) == int(g["action"][t]), t
        rec.append(action, o_t, la_t, cuda([g["reward"][t]], torch.int32).view(-1), visits,
                   cuda([g["value"][t]], torch.float32).view(-1), cuda([g["done"][t]], torch.uint8).view(-1))
        prev_legal = la_t
        if g["done"][t]:
            got += rec.flush()
    assert len(got) == len(eps)
    for ep, want in zip(got, eps):
        for k in ("vis", "root", "a", "r", "o", "la"):
            assert ep[k].shape == want[k].shape and (ep[k] == want[k]).all(), k
        assert ep["vis"].dtype == np.float64
