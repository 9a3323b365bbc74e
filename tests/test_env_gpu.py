"""GPU parity tests of the Hanabi kernels (through HanabiVecEnv / HanabiEnv -> C ABI) against golden
episodes from the reference's own Python HanabiEnv and against the CPU oracle in lock step.
Everything here is integer/bit work: the bar is bit-exact."""
import numpy as np
import pytest
import torch

from helpers import env_policy, golden_files, load_golden, playable_from_dump, unpack_env_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_files("env_"))
def test_cuda_replays_reference_python_env(name):
    from gpu_adapters import GpuHanabiGame
    g = load_golden(name)
    glob_, loc = unpack_env_golden(g)
    env = GpuHanabiGame(int(g["preset"]), int(g["seed"]))
    starts = set(int(s) for s in g["ep_start"])
    for t in range(len(g["action"])):
        if t in starts:
            go, lo, legal = env.reset()
            r, d, sc = 0, False, 0
        else:
            go, lo, legal, r, d, sc = env.step(int(g["action"][t]))
        assert (go == glob_[t]).all(), f"global obs differs at step {t}: bits {np.flatnonzero(go != glob_[t])[:10]}"
        assert (lo == loc[t]).all(), f"local obs differs at step {t}"
        assert (legal == g["legal"][t]).all(), f"legal mask differs at step {t}"
        assert (r, int(d), sc) == (int(g["reward"][t]), int(g["done"][t]), int(g["score"][t])), t


@pytest.mark.parametrize("preset", [0, 1])
@pytest.mark.parametrize("mode", ["random", "noplay", "smart"])
def test_batched_lockstep_vs_oracle(preset, mode):
    """256 games with distinct seeds advance in one launch per step; every game is compared with its
    own oracle instance: observations, legal masks, reward/done/score and the full hidden state.
    Finished games are reset (same game object => the mt19937 stream continues)."""
    from hanabizero_b200.hanabi_env import HanabiVecEnv
    from oracle import loader as L
    N, T = 256, 260 if preset == 0 else 160
    seeds = np.arange(N) * 3 + 100 * preset
    vec = HanabiVecEnv(N, "Hanabi-Full" if preset == 0 else "Hanabi-Small", seeds)
    cpus = [L.oracle_hanabi(preset, int(s)) for s in seeds]
    rng = np.random.default_rng(5 + preset)
    g, l, a = vec.reset_all()
    g, l, a = g.cpu().numpy(), l.cpu().numpy(), a.cpu().numpy()
    legal = []
    for i, c in enumerate(cpus):
        go, lo, lg = c.reset()
        assert (go == g[i]).all() and (lo == l[i]).all() and (lg == a[i]).all(), i
        legal.append(lg)
    episodes = 0
    for t in range(T):
        dumps = vec.dump().cpu().numpy()
        acts = np.zeros(N, np.int32)
        for i, c in enumerate(cpus):
            d = c.dump()
            assert (d == dumps[i]).all(), (t, i)
            acts[i] = env_policy(mode, rng, c.hand_size, legal[i], playable_from_dump(d, c.colors, c.ranks, c.hand_size))
        g, l, a, r, dn, sc = vec.step_all(torch.from_numpy(acts).cuda())
        vec.check()
        g, l, a, r, dn, sc = (x.cpu().numpy() for x in (g, l, a, r, dn, sc))
        reset_mask = np.zeros(N, np.uint8)
        for i, c in enumerate(cpus):
            go, lo, lg, rr, dd, ss = c.step(int(acts[i]))
            assert (go == g[i]).all(), (t, i, np.flatnonzero(go != g[i])[:10])
            assert (lo == l[i]).all() and (lg == a[i]).all(), (t, i)
            assert (rr, int(dd), ss) == (int(r[i]), int(dn[i]), int(sc[i])), (t, i)
            legal[i] = lg
            if dd:
                reset_mask[i] = 1
        if reset_mask.any():
            episodes += int(reset_mask.sum())
            g, l, a = vec.reset_all(mask=torch.from_numpy(reset_mask).cuda())
            g, l, a = g.cpu().numpy(), l.cpu().numpy(), a.cpu().numpy()
            for i in np.flatnonzero(reset_mask):
                go, lo, lg = cpus[i].reset()
                assert (go == g[i]).all() and (lo == l[i]).all() and (lg == a[i]).all(), (t, i)
                legal[i] = lg
    assert episodes > N // 4


def test_auto_reset_equals_explicit_reset():
    from hanabizero_b200.hanabi_env import HanabiVecEnv
    N, T = 128, 120
    seeds = np.arange(N) + 7
    a_env, b_env = HanabiVecEnv(N, "Hanabi-Small", seeds), HanabiVecEnv(N, "Hanabi-Small", seeds)
    ga, _, la = a_env.reset_all()
    gb, _, lb = b_env.reset_all()
    gen = torch.Generator(device="cuda").manual_seed(0)
    for t in range(T):
        assert torch.equal(ga, gb) and torch.equal(la, lb)
        acts = torch.multinomial(la + 1e-9, 1, generator=gen).view(-1).int()
        ga, _, la, ra, da, sa = a_env.step_all(acts, auto_reset=True)
        gb, _, lb, rb, db, sb = b_env.step_all(acts)
        assert torch.equal(ra, rb) and torch.equal(da, db) and torch.equal(sa, sb)
        if db.any():
            gb, _, lb = b_env.reset_all(mask=db.clone())
    a_env.check(); b_env.check()


def test_full_size_invariants_4096_games():
    """BASELINE-size batch: 4096 Hanabi-Full games, random legal play with auto-reset; checks
    domain invariants that do not need the oracle: card conservation, legal mask consistency,
    observation bits in {0,1}, score/reward bookkeeping."""
    from hanabizero_b200.hanabi_env import HanabiVecEnv
    N = 4096
    vec = HanabiVecEnv(N, "Hanabi-Full", np.arange(N))
    g, l, a = vec.reset_all()
    gen = torch.Generator(device="cuda").manual_seed(1)
    prev_score = torch.zeros(N, dtype=torch.int32, device="cuda")
    for t in range(150):
        assert ((g == 0) | (g == 1)).all() and (a.sum(1) > 0).all()
        assert torch.equal(g[:, vec.own_len:], l)
        acts = torch.multinomial(a, 1, generator=gen).view(-1).int()
        g, l, a, r, d, s = vec.step_all(acts, auto_reset=True)
        assert torch.equal(r, s - prev_score)
        prev_score = torch.where(d.bool(), torch.zeros_like(s), s)
        dump = vec.dump()
        C, R, H = 5, 5, 5
        deck = dump[:, 5 + C:5 + C + C * R].sum(1)
        disc = dump[:, 5 + C + C * R:5 + C + 2 * C * R].sum(1)
        fw = dump[:, 5:5 + C].sum(1)
        base = 5 + C + 2 * C * R
        hands = dump[:, base] + dump[:, base + 1 + 5 * H]
        assert (deck == dump[:, 3]).all()
        assert (deck + disc + fw + hands == 50).all(), "cards are conserved"
    vec.check()


def test_scalar_dropin_api_and_errors():
    from hanabizero_b200.env_wrapper import HanabiControlWrapper
    from hanabizero_b200.hanabi_env import HanabiEnv
    from oracle import loader as L
    env = HanabiEnv({"hanabi_name": "Hanabi-Full", "seed": None})
    cpu = L.oracle_hanabi(0, 0)
    assert env.players == 2 and env.action_space[0].n == 20 and env.num_moves() == 20
    assert env.vectorized_observation_shape() == [658] and env.vectorized_share_observation_shape() == [783]
    share, obs, legal = env.reset()
    go, lo, lg = cpu.reset()
    assert share == go.tolist() and obs == lo.tolist() and legal == lg.astype(float).tolist()
    assert isinstance(share, list) and len(share) == 785 and len(obs) == 660 and len(legal) == 20
    with pytest.raises(ValueError):
        env.step(np.int64(5))                       # rl_env.py:415: only dict or int
    with pytest.raises(ValueError):
        env.step(int(np.flatnonzero(lg == 0)[0]))   # illegal uid: the reference aborts the process
    a = int(np.flatnonzero(lg)[-1])
    out = env.step(a)
    ref = cpu.step(a)
    assert out[0] == ref[0].tolist() and out[1] == ref[1].tolist() and out[5] == ref[2].astype(float).tolist()
    assert (out[2], out[3], out[4]) == (ref[3], ref[4], {"score": ref[5]})
    out = env.step({"action_type": "PLAY", "card_index": 0})
    ref = cpu.step(5)
    assert out[0] == ref[0].tolist() and out[2] == ref[3]
    assert env.state.score() == ref[5] and env.state.cur_player() == int(cpu.dump()[0])
    w = HanabiControlWrapper(HanabiEnv({"hanabi_name": "Hanabi-Small", "seed": 3}), discount=0.999, mdp="local")
    o, lg2 = w.reset()
    assert isinstance(o, np.ndarray) and o.shape == (173,) and lg2.shape == (11,)
    o, r, d, info, lg2 = w.step(int(np.flatnonzero(lg2)[0]))
    assert o.shape == (173,) and info.item()["score"] >= 0 and w.action_space_size == 11


@pytest.mark.parametrize("name,preset", [("Hanabi-Full", 0), ("Hanabi-Small", 1)])
def test_byte_observations_equal_float_observations(name, preset):
    """hz_envs_*_u8: the 0/1 byte rows (aligned, padded rows take the 4-bytes-per-lane store; odd strides and
    odd base addresses take the byte store) hold exactly the float rows' values, for reset, step and
    auto-reset, and match the oracle on the first observation."""
    from hanabizero_b200.hanabi_env import HanabiVecEnv
    from oracle import loader as L
    N, T = 130, 90
    seeds = np.arange(N) + 11
    f_env = HanabiVecEnv(N, name, seeds)
    b_env = HanabiVecEnv(N, name, seeds, obs_dtype=torch.uint8)
    u_env = HanabiVecEnv(N, name, seeds, obs_dtype=torch.uint8)
    gd, ld, A = f_env.global_dim, f_env.local_dim, f_env.num_actions
    assert b_env.global_obs.stride(0) % 16 == 0 and b_env.global_obs.shape == (N, gd)
    # unaligned targets for the third env: odd row stride and an odd base address
    ug = torch.zeros(N * (gd + 3) + 1, dtype=torch.uint8, device="cuda")[1:].view(N, gd + 3)[:, :gd]
    ul = torch.zeros(N * (ld + 1) + 3, dtype=torch.uint8, device="cuda")[3:].view(N, ld + 1)[:, :ld]
    fg, fl, fa = f_env.reset_all()
    bg, bl, ba = b_env.reset_all()
    u_env.reset_all(observe=False)
    _, _, ua = u_env.observe(out_global=ug, out_local=ul)
    go, lo, lg = L.oracle_hanabi(preset, int(seeds[5])).reset()
    assert (bg[5].cpu().numpy() == go).all() and (bl[5].cpu().numpy() == lo).all() and (ba[5].cpu().numpy() == lg).all()
    gen = torch.Generator(device="cuda").manual_seed(3)
    for t in range(T):
        for g8, l8, a8 in ((bg, bl, ba), (ug, ul, ua)):
            assert g8.dtype == torch.uint8 and torch.equal(g8.float(), fg), t
            assert torch.equal(l8.float(), fl) and torch.equal(a8.float(), fa), t
        acts = torch.multinomial(fa, 1, generator=gen).view(-1).int()
        fg, fl, fa, fr, fd, fs = f_env.step_all(acts, auto_reset=True)
        bg, bl, ba, br, bd, bs = b_env.step_all(acts, auto_reset=True)
        _, _, ua, ur, ud, us = u_env.step_all(acts, auto_reset=True, out_global=ug, out_local=ul)
        assert torch.equal(fr, br) and torch.equal(fd, bd) and torch.equal(fs, bs)
        assert torch.equal(fr, ur) and torch.equal(fd, ud) and torch.equal(fs, us)
    with pytest.raises(TypeError):
        f_env.observe(out_global=ug)     # float legal/local with a byte global row
    for e in (f_env, b_env, u_env):
        e.check()


@pytest.mark.parametrize("name", ["Hanabi-Full", "Hanabi-Small"])
def test_bit_packed_rows_equal_float_observations(name):
    """hz_envs_step_observe_bits: the packed row (observation bits, legal mask, reward, done, score) unpacks to
    exactly what the float launch returns, for reset, step and auto-reset (finished games re-dealt in the launch)."""
    from hanabizero_b200.hanabi_env import HanabiVecEnv
    N, T = 130, 120
    seeds = np.arange(N) + 23
    f_env, p_env = HanabiVecEnv(N, name, seeds), HanabiVecEnv(N, name, seeds)
    fg, fl, fa = f_env.reset_all()
    p_env.reset_all(observe=False)
    rows = p_env.step_bits(None)
    assert rows.shape == (N, p_env.bits_words) and rows.dtype == torch.int32
    gen = torch.Generator(device="cuda").manual_seed(4)
    fr = fd = fs = None
    resets = 0
    for t in range(T):
        u = p_env.unpack_bits(rows.cpu())
        assert (u["global_obs"] == fg.cpu().numpy()).all(), t
        assert (u["local_obs"] == fl.cpu().numpy()).all() and (u["legal"] == fa.cpu().numpy()).all(), t
        if fr is not None:
            assert (u["reward"] == fr.cpu().numpy()).all() and (u["done"] == fd.cpu().numpy().astype(bool)).all(), t
            assert (u["score"] == fs.cpu().numpy()).all(), t
            resets += int(u["done"].sum())
        else:
            assert (u["reward"] == 0).all() and not u["done"].any()
        acts = torch.multinomial(fa, 1, generator=gen).view(-1).int()
        fg, fl, fa, fr, fd, fs = f_env.step_all(acts, auto_reset=True)
        rows = p_env.step_bits(acts, auto_reset=True)
    assert resets > 0
    assert torch.equal(f_env.dump(), p_env.dump())
    f_env.check()
    p_env.check()


@pytest.mark.parametrize("fmt,zero_copy", [("bits", True), ("bits", False), ("u8", False), ("f32", False)])
def test_env_pipeline_matches_direct_stepping(fmt, zero_copy):
    """EnvPipeline (two groups, each on its own stream: pinned host actions in, kernel, result out) plays exactly the
    games a plain step loop plays — with staging copies, and with the kernel reading / writing the pinned host buffers
    itself (hz_envs_host_step)."""
    from hanabizero_b200.hanabi_env import EnvPipeline, HanabiVecEnv
    n, T, G = 48, 40, 2
    seeds = [np.arange(n) + 7 + 1000 * g for g in range(G)]
    refs = [HanabiVecEnv(n, "Hanabi-Full", sd) for sd in seeds]
    envs = [HanabiVecEnv(n, "Hanabi-Full", sd) for sd in seeds]
    A, gd = envs[0].num_actions, envs[0].global_dim
    ref_g, ref_l = [], []
    for r, e in zip(refs, envs):
        g0, _, l0 = r.reset_all()
        ref_g.append(g0.clone())
        ref_l.append(l0.clone())
        e.reset_all(observe=False)
    pipe = EnvPipeline(envs, fmt=fmt, zero_copy=zero_copy)
    h_act = [torch.zeros(n, dtype=torch.int32).pin_memory() for _ in range(G)]
    for g in range(G):
        pipe.observe_now(g)
    rng = np.random.default_rng(0)
    for t in range(T):
        for g in range(G):
            obs, leg = pipe.wait(g)
            if fmt == "bits":
                u = envs[g].unpack_bits(obs, leg)
                got_g, got_l = u["global_obs"], u["legal"]
                assert obs.shape == (n, envs[g].bits_words - 4) and leg.shape == (n, 4)
                picked = torch.zeros(n, dtype=torch.int32)
                envs[g].random_legal_host(leg, picked, seed=3, step=t)
                assert (got_l[np.arange(n), picked.numpy()] == 1).all()      # the host helper only picks legal moves
            else:
                got_g, got_l = obs.numpy()[:, :gd], leg.numpy()
            assert (got_g == ref_g[g].cpu().numpy()).all(), (t, g)
            assert (got_l == ref_l[g].cpu().numpy()).all(), (t, g)
            acts = np.array([rng.choice(np.flatnonzero(row)) for row in got_l], np.int32)
            h_act[g].copy_(torch.from_numpy(acts))
            pipe.step(g, h_act[g])
            gg, _, ll, _, _, _ = refs[g].step_all(torch.from_numpy(acts).cuda(), auto_reset=True)
            ref_g[g], ref_l[g] = gg.clone(), ll.clone()
    pipe.drain()
    for r, e in zip(refs, envs):
        assert torch.equal(r.dump(), e.dump())
        e.check()


def test_fused_random_policy_equals_its_host_twin():
    """hz_envs_set_random_policy: every observing launch also writes a random legal move per game; feeding the buffer
    back makes a random-play step one launch.  Draw d of game i must equal hz_host_random_legal(seed, step=d) on the
    legal mask of the position it was drawn for (CUDA-graph replays included: the draw counters live on the device)."""
    from hanabizero_b200.hanabi_env import HanabiVecEnv
    n, seed = 130, 99
    env = HanabiVecEnv(n, "Hanabi-Full", np.arange(n) + 5)
    env.reset_all(observe=False)
    buf = torch.zeros(n, dtype=torch.int32, device="cuda")
    env.set_random_policy(buf, seed=seed)
    rows = torch.zeros(n, env.bits_words - 4, dtype=torch.int32, device="cuda")
    meta = torch.zeros(n, 4, dtype=torch.int32, device="cuda")
    env.step_bits(None, out=rows, out_meta=meta)                 # observe: draw 0
    host = torch.zeros(n, dtype=torch.int32)
    graph = None
    for d in range(40):
        env.random_legal_host(meta.cpu(), host, seed=seed, step=d)
        assert torch.equal(buf.cpu(), host), d
        if d == 20:                                              # the second half replays a captured step
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                env.step_bits(buf, auto_reset=True, out=rows, out_meta=meta)
        if graph is not None:
            graph.replay()
        else:
            env.step_bits(buf, auto_reset=True, out=rows, out_meta=meta)   # plays the picks, writes the next ones
    env.check()
    env.set_random_policy(None)
    before = buf.clone()
    env.step_bits(buf, auto_reset=True, out=rows, out_meta=meta)
    assert torch.equal(buf, before)                               # switched off: the buffer is no longer written
    env.check()


def test_host_step_rejects_pageable_memory_and_observes_without_actions():
    """hz_envs_host_step reads and writes the caller's HOST buffers from the kernel: pinned memory works (also with
    h_actions = NULL: observe only), ordinary pageable memory must be refused with an error instead of faulting."""
    from hanabizero_b200 import _lib
    from hanabizero_b200.hanabi_env import HanabiVecEnv
    lib = _lib.load()
    n = 33
    env = HanabiVecEnv(n, "Hanabi-Full", np.arange(n) + 3)
    g0, _, l0 = env.reset_all()
    w = env.bits_words - 4
    bits = torch.zeros(n, w, dtype=torch.int32).pin_memory()
    meta = torch.zeros(n, 4, dtype=torch.int32).pin_memory()
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.hz_envs_host_step(env._h, st, None, 1, bits.data_ptr(), w, meta.data_ptr()))
    _lib.check(lib.hz_envs_host_wait(env._h))
    u = env.unpack_bits(bits, meta)
    assert (u["global_obs"] == g0.cpu().numpy()).all() and (u["legal"] == l0.cpu().numpy()).all()
    pageable = torch.zeros(n, w, dtype=torch.int32)
    rc = lib.hz_envs_host_step(env._h, st, None, 1, pageable.data_ptr(), w, meta.data_ptr())
    assert rc != 0 and b"page-locked" in lib.hz_last_error()
    env.check()
