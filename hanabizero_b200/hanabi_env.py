"""Hanabi environment on the sm_100a game kernels.

`HanabiVecEnv` is the B200-native form: N independent games resident in HBM, stepped by one
kernel launch, observations / legal masks / rewards produced as CUDA tensors.

`HanabiEnv` is the drop-in for the reference's gym-like class
(/root/reference/envs/hanabi/rl_env.py:26-442): same constructor argument
(`{'hanabi_name': 'Hanabi-Full' | 'Hanabi-Small', 'seed': int | None}`), same `reset()` /
`step(action)` return tuples, same attributes the callers touch (`players`, `action_space[i].n`,
`num_moves()`, `vectorized_observation_shape()`, `vectorized_share_observation_shape()`, `game`,
`state`).  It is a one-game view of the same kernels — there is no Python or CPU game logic here.
"""
import numpy as np
import torch

from . import _lib
from ._lib import check, ptr

PRESETS = {"Hanabi-Full": 0, "Hanabi-Small": 1}  # rl_env.py:110-131
COLOR_CHAR = ["R", "Y", "G", "W", "B"]            # pyhanabi.py:26
CHANCE_PLAYER_ID = -1


class HanabiVecEnv:
    """N games of one preset on one GPU.  Game i owns its own std::mt19937 stream seeded with
    seeds[i], which persists across resets exactly like the reference's per-env HanabiGame object
    (rl_env.py:137,249)."""

    def __init__(self, num_games, hanabi_name="Hanabi-Full", seeds=None, device=None, obs_dtype=torch.float32):
        """obs_dtype: torch.float32 (what the network eats) or torch.uint8 (0/1 bytes: the encoder's own value
        type at a quarter of the traffic; rows are padded to a 16-byte multiple so the kernel stores words)."""
        if obs_dtype not in (torch.float32, torch.uint8):
            raise ValueError("obs_dtype must be torch.float32 or torch.uint8")
        self.obs_dtype = obs_dtype
        if hanabi_name not in PRESETS:
            raise ValueError("Unknown environment {}".format(hanabi_name))  # rl_env.py:133
        self.num_games = int(num_games)
        self.hanabi_name = hanabi_name
        from .cytree import _device_index   # device="cuda" (no index) means the CURRENT device, not cuda:0
        self.device_index = _device_index(device)
        self.device = torch.device("cuda", self.device_index)
        if seeds is None:
            seeds = np.zeros(self.num_games, np.int32)  # seed=None -> 0 (rl_env.py:106-109)
        seeds = np.ascontiguousarray(np.broadcast_to(np.asarray(seeds, np.int64), (self.num_games,)).astype(np.int32))
        self._lib = _lib.load()
        h = _lib.C.c_void_p()
        check(self._lib.hz_envs_create(_lib.C.byref(h), self.device_index, self.num_games,
                                       PRESETS[hanabi_name], seeds.ctypes.data))
        self._h = h
        dims = np.zeros(12, np.int32)
        check(self._lib.hz_envs_dims(self._h, dims.ctypes.data))
        (self.enc_len, self.own_len, self.players, self.num_actions, self.colors, self.ranks,
         self.hand_size, self.max_info, self.max_life, self.local_dim, self.global_dim,
         self.dump_len) = [int(x) for x in dims]
        n, dev = self.num_games, self.device
        self.reward = torch.zeros(n, dtype=torch.int32, device=dev)
        self.done = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.score = torch.zeros(n, dtype=torch.int32, device=dev)
        pad = (lambda d: (d + 15) // 16 * 16) if obs_dtype == torch.uint8 else (lambda d: d)
        self.global_obs = torch.zeros(n, pad(self.global_dim), dtype=obs_dtype, device=dev)[:, :self.global_dim]
        self.local_obs = torch.zeros(n, pad(self.local_dim), dtype=obs_dtype, device=dev)[:, :self.local_dim]
        self.legal = torch.zeros(n, self.num_actions, dtype=obs_dtype, device=dev)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None:
            try:
                self._lib.hz_envs_destroy(h)
            except Exception:
                pass
            self._h = None

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    @staticmethod
    def _ld(t):
        return 0 if t is None else t.stride(0)

    def reset_all(self, mask=None, observe=True):
        """HanabiEnv.reset for every game (or those with mask[i] != 0). Returns
        (global_obs, local_obs, legal) CUDA tensors when observe."""
        m = None if mask is None else mask.to(self.device, torch.uint8).contiguous()
        check(self._lib.hz_envs_reset(self._h, self._stream(), ptr(m)))
        if observe:
            return self.observe()

    def observe(self, out_global=None, out_local=None, out_legal=None):
        """Current player's observation tuple of every game, written into the given (possibly
        strided: row stride >= dim) float32 CUDA tensors or the env's own buffers."""
        g = self.global_obs if out_global is None else out_global
        l = self.local_obs if out_local is None else out_local
        a = self.legal if out_legal is None else out_legal
        fn = self._lib.hz_envs_observe_u8 if self._all_u8(g, l, a) else self._lib.hz_envs_observe
        check(fn(self._h, self._stream(), ptr(g), self._ld(g), ptr(l), self._ld(l), ptr(a)))
        return g, l, a

    @staticmethod
    def _all_u8(*tensors):
        """True if the output tensors are uint8, False if float32; mixed sets are rejected."""
        kinds = {t.dtype for t in tensors if t is not None}
        if kinds <= {torch.uint8}:
            return bool(kinds)
        if kinds <= {torch.float32}:
            return False
        raise TypeError(f"observation outputs must be all float32 or all uint8, got {sorted(map(str, kinds))}")

    def step_all(self, actions, active=None, auto_reset=False, observe=True, out_global=None,
                 out_local=None, out_legal=None, want_local=True, want_global=True):
        """HanabiEnv.step for every (active) game in one launch.  actions: int32 CUDA tensor [N]
        (anything else is converted).  Returns (global_obs, local_obs, legal, reward, done, score).
        With auto_reset a finished game is re-dealt inside the same launch and the returned
        observation is the first one of its next episode (reward/done/score describe the finished
        step).  Illegal actions leave the game untouched; call check() to surface them."""
        if not (isinstance(actions, torch.Tensor) and actions.dtype == torch.int32 and actions.is_cuda):
            actions = torch.as_tensor(np.asarray(actions, np.int32) if not isinstance(actions, torch.Tensor)
                                      else actions, dtype=torch.int32).to(self.device)
        actions = actions.contiguous().view(-1)
        if actions.numel() != self.num_games:
            raise ValueError(f"expected {self.num_games} actions, got {actions.numel()}")
        act = None if active is None else active.to(self.device, torch.uint8).contiguous()
        if not observe:
            check(self._lib.hz_envs_step(self._h, self._stream(), ptr(actions), ptr(act),
                                         ptr(self.reward), ptr(self.done), ptr(self.score)))
            return None, None, None, self.reward, self.done, self.score
        g = (self.global_obs if out_global is None else out_global) if want_global else None
        l = (self.local_obs if out_local is None else out_local) if want_local else None
        a = self.legal if out_legal is None else out_legal
        fn = self._lib.hz_envs_step_observe_u8 if self._all_u8(g, l, a) else self._lib.hz_envs_step_observe
        check(fn(self._h, self._stream(), ptr(actions), ptr(act), 1 if auto_reset else 0, ptr(self.reward),
                 ptr(self.done), ptr(self.score), ptr(g), self._ld(g), ptr(l), self._ld(l), ptr(a)))
        return g, l, a, self.reward, self.done, self.score

    def check(self):
        """Synchronises and raises IllegalMoveError if any game was handed an illegal move."""
        bad = _lib.C.c_int32(-1)
        check(self._lib.hz_envs_check(self._h, self._stream(), _lib.C.byref(bad)))

    def dump(self):
        """int32 [N, dump_len] full hidden state (layout of oracle/hanabi_oracle.c:ohanabi_dump)."""
        out = torch.empty(self.num_games, self.dump_len, dtype=torch.int32, device=self.device)
        check(self._lib.hz_envs_dump(self._h, self._stream(), ptr(out)))
        return out


class Discrete:
    """gym.spaces.Discrete stand-in (rl_env.py:21,141): the callers only read `.n`."""

    def __init__(self, n):
        self.n = n


class _GameView:
    """The part of pyhanabi.HanabiGame the env's callers use (pyhanabi.py:679-785)."""

    def __init__(self, vec):
        self._v = vec

    def num_players(self): return self._v.players
    def num_colors(self): return self._v.colors
    def num_ranks(self): return self._v.ranks
    def hand_size(self): return self._v.hand_size
    def max_information_tokens(self): return self._v.max_info
    def max_life_tokens(self): return self._v.max_life
    def max_moves(self): return self._v.num_actions

    def num_cards(self, color, rank):
        return 3 if rank == 0 else (1 if rank == self._v.ranks - 1 else 2)

    def get_move(self, move_uid):
        return move_dict(self._v, int(move_uid))

    def get_move_uid(self, move):
        return move_uid(self._v, move)


class _StateView:
    """Read-only view of pyhanabi.HanabiState (pyhanabi.py:495-660) decoded from the device state."""

    def __init__(self, vec):
        self._v = vec

    def _d(self):
        return self._v.dump()[0].cpu().numpy()

    def cur_player(self): return int(self._d()[0])
    def information_tokens(self): return int(self._d()[1])
    def life_tokens(self): return int(self._d()[2])
    def deck_size(self): return int(self._d()[3])
    def is_terminal(self): return bool(self._d()[4])
    def fireworks(self): return [int(x) for x in self._d()[5:5 + self._v.colors]]

    def score(self):
        d = self._d()
        return 0 if d[2] <= 0 else int(d[5:5 + self._v.colors].sum())

    def player_hands(self):
        v, d = self._v, self._d()
        base = 5 + v.colors + 2 * v.colors * v.ranks
        hands = []
        for p in range(v.players):
            off = base + p * (1 + 5 * v.hand_size)
            hands.append([{"color": COLOR_CHAR[int(d[off + 1 + 5 * k]) // v.ranks],
                           "rank": int(d[off + 1 + 5 * k]) % v.ranks} for k in range(int(d[off]))])
        return hands

    def discard_counts(self):
        v, d = self._v, self._d()
        o = 5 + v.colors + v.colors * v.ranks
        return d[o:o + v.colors * v.ranks].reshape(v.colors, v.ranks).tolist()


def move_dict(v, uid):
    """HanabiGame::ConstructMove (hanabi_game.cc:159-183) as the dict form of pyhanabi's to_dict."""
    h, c = v.hand_size, v.colors
    if uid < 0 or uid >= v.num_actions:
        raise ValueError(f"move uid {uid} out of range")
    if uid < h:
        return {"action_type": "DISCARD", "card_index": uid}
    if uid < 2 * h:
        return {"action_type": "PLAY", "card_index": uid - h}
    uid -= 2 * h
    if uid < (v.players - 1) * c:
        return {"action_type": "REVEAL_COLOR", "target_offset": 1 + uid // c, "color": COLOR_CHAR[uid % c]}
    uid -= (v.players - 1) * c
    return {"action_type": "REVEAL_RANK", "target_offset": 1 + uid // v.ranks, "rank": uid % v.ranks}


def move_uid(v, action):
    """HanabiGame::GetMoveUid (hanabi_game.cc:79-95) from the dict form (rl_env.py:516-575)."""
    assert isinstance(action, dict), "Expected dict, got: {}".format(action)
    assert "action_type" in action, "Action should contain `action_type`. action: {}".format(action)
    t, h = action["action_type"], v.hand_size
    if t == "DISCARD":
        return int(action["card_index"])
    if t == "PLAY":
        return h + int(action["card_index"])
    if t == "REVEAL_COLOR":
        assert isinstance(action["color"], str)
        return 2 * h + (int(action["target_offset"]) - 1) * v.colors + COLOR_CHAR.index(action["color"])
    if t == "REVEAL_RANK":
        return (2 * h + (v.players - 1) * v.colors + (int(action["target_offset"]) - 1) * v.ranks
                + int(action["rank"]))
    raise ValueError("Unknown action_type: {}".format(t))


class HanabiEnv:
    """Drop-in for envs.hanabi.rl_env.HanabiEnv (rl_env.py:26-442) backed by a one-game batch."""

    def __init__(self, args, device=None):
        seed = 0 if args["seed"] is None else args["seed"]  # rl_env.py:106-109
        self._vec = HanabiVecEnv(1, args["hanabi_name"], [seed], device=device, obs_dtype=torch.uint8)
        v = self._vec
        self.game = _GameView(v)
        self.state = None
        self.players = v.players
        self.action_space = [Discrete(v.num_actions) for _ in range(v.players)]
        self.observation_space = [[v.enc_len + v.players] for _ in range(v.players)]
        self.share_observation_space = [[v.own_len + v.enc_len + v.players] for _ in range(v.players)]
        self._action = torch.zeros(1, dtype=torch.int32, device=v.device)

    def vectorized_observation_shape(self):
        return [self._vec.enc_len]

    def vectorized_share_observation_shape(self):
        return [self._vec.own_len + self._vec.enc_len]

    def num_moves(self):
        return self._vec.num_actions

    def _tuple(self, g, l, a):
        host = torch.cat((g[0], l[0], a[0])).cpu().numpy()
        gd, ld = self._vec.global_dim, self._vec.local_dim
        share_obs = host[:gd].astype(np.int64).tolist()
        obs = host[gd:gd + ld].astype(np.int64).tolist()
        legal = list(host[gd + ld:].astype(np.float64))
        return share_obs, obs, legal

    def reset(self, choose=True):
        """rl_env.py:148-267 -> (share_obs, obs, available_actions)."""
        if not choose:
            # the reference's choose=False branch references undefined names and cannot run
            raise NotImplementedError("reset(choose=False) is broken in the reference (rl_env.py:264-266)")
        g, l, a = self._vec.reset_all()
        self.state = _StateView(self._vec)
        return self._tuple(g, l, a)

    def step(self, action):
        """rl_env.py:292-442 -> (share_obs, obs, reward, done, {'score'}, available_actions)."""
        if isinstance(action, dict):
            uid = move_uid(self._vec, action)
        elif isinstance(action, int) and not isinstance(action, bool):
            assert action != -1  # rl_env.py:404-411 ends in `assert False`
            uid = action
        else:
            raise ValueError("Expected action as dict or int, got: {}".format(action))  # rl_env.py:415
        if self.state is None:
            raise RuntimeError("step() before reset()")
        self._action.fill_(uid)
        g, l, a, reward, done, score = self._vec.step_all(self._action)
        try:
            self._vec.check()
        except _lib.IllegalMoveError as e:
            if isinstance(action, dict):  # rl_env.py:569-572
                raise AssertionError("Illegal action: {}".format(action)) from e
            raise
        share_obs, obs, legal = self._tuple(g, l, a)
        rds = torch.stack((reward[0], done[0].int(), score[0])).cpu().tolist()
        return share_obs, obs, int(rds[0]), bool(rds[1]), {"score": int(rds[2])}, legal

    def close(self):
        pass
