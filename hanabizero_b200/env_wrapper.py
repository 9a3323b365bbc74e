"""Host-side view the reference's actors hold on one game: drop-in for
config/hanabi_control/env_wrapper.py:6-34 (`HanabiControlWrapper`) and its base core/game.py:26-46 (`Game`).

The wrapper only selects which observation the agent sees — the global / MDP one (own hand ‖ canonical
encoding ‖ turn, 785 or 193 wide) or the local / POMDP one (660 / 173) — and hands every field back as a
numpy array, as the reference does.  All game logic stays in the CUDA kernels behind `HanabiEnv`.
"""
import numpy as np

_VIEWS = {"global": 0, "local": 1}   # index into (share_obs, obs) as returned by HanabiEnv


class Game:
    """core/game.py:26-46: the minimal environment handle the self-play code passes around."""

    def __init__(self, env, action_space_size, discount, config=None):
        self.env, self.action_space_size = env, action_space_size
        self.discount, self.config = discount, config

    def legal_actions(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError()

    def close(self, *args, **kwargs):
        self.env.close(*args, **kwargs)

    def render(self, *args, **kwargs):
        self.env.render(*args, **kwargs)


class HanabiControlWrapper(Game):
    def __init__(self, env, discount, cvt_string=False, mdp="global"):
        if mdp not in _VIEWS:   # the reference falls through and returns None for anything else
            raise ValueError(f"mdp must be one of {sorted(_VIEWS)}, got {mdp!r}")
        super().__init__(env, env.num_moves(), discount)
        self.cvt_string, self.mdp = cvt_string, mdp

    def _view(self, pair):
        return np.array(pair[_VIEWS[self.mdp]])

    def legal_actions(self):
        return list(range(self.action_space_size))

    def reset(self, **kwargs):
        *views, legal = self.env.reset()
        return self._view(views), np.array(legal)

    def step(self, action):
        *views, reward, done, info, legal = self.env.step(action)
        return (self._view(views),) + tuple(np.array(x) for x in (reward, done, info, legal))

    def close(self):
        pass
