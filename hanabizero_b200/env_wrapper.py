"""Drop-in for config/hanabi_control/env_wrapper.py:6-34 (HanabiControlWrapper) and the `Game`
base it derives from (core/game.py:26-46): selects the global (MDP, 785/193 wide) or local
(POMDP, 660/173 wide) observation and wraps every return value in numpy arrays."""
import numpy as np


class Game:
    """core/game.py:26-46."""

    def __init__(self, env, action_space_size, discount, config=None):
        self.env = env
        self.action_space_size = action_space_size
        self.discount = discount
        self.config = config

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
        super().__init__(env, env.num_moves(), discount)
        self.cvt_string = cvt_string
        if mdp not in ("global", "local"):
            raise ValueError("mdp must be 'global' or 'local'")  # the reference silently returns None
        self.mdp = mdp

    def legal_actions(self):
        return list(range(self.action_space_size))

    def step(self, action):
        global_state, state, reward, done, info, legal_actions = self.env.step(action)
        obs = global_state if self.mdp == "global" else state
        return np.array(obs), np.array(reward), np.array(done), np.array(info), np.array(legal_actions)

    def reset(self, **kwargs):
        global_state, state, legal_actions = self.env.reset()
        obs = global_state if self.mdp == "global" else state
        return np.array(obs), np.array(legal_actions)

    def close(self):
        pass
