"""The representation / dynamics / prediction network — stays in PyTorch (north-star: it is the only
dense contraction on the path; cuBLAS owns it).  Same architecture and the same parameter names as
the reference networks so that a reference checkpoint's state_dict loads unchanged:
  MuZeroNet      (Hanabi-Small)  /root/reference/config/hanabi_control/model.py:127-235
  MuZeroNetFull  (Hanabi-Full)   /root/reference/config/hanabi_control/model.py:237-335
  initial/recurrent_inference    /root/reference/core/model.py:61-84
  inverse value/reward transform /root/reference/core/config.py:204-232

What is new is the hand-off: `initial_inference_device` / `recurrent_inference_device` return CUDA
tensors (value and reward already passed through the inverse categorical transform on the device),
so the search loop never leaves HBM.  `initial_inference` / `recurrent_inference` keep the
reference's contract (numpy on the host in eval mode) for callers that want it.
"""
from typing import NamedTuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

FEATURE_SIZE = 512


class NetworkOutput(NamedTuple):  # core/model.py:9-14
    value: object
    reward: object
    policy_logits: object
    hidden_state: object


class _ResBlock(nn.Module):
    """fc1-bn1-fc2-bn2 with the skip joined either after bn1 (ResMLP, model.py:7-30) or after bn2
    (NewResMLP, model.py:32-56)."""

    def __init__(self, dim, skip_after_first):
        super().__init__()
        self.fc1, self.bn1 = nn.Linear(dim, dim), nn.BatchNorm1d(dim)
        self.fc2, self.bn2 = nn.Linear(dim, dim), nn.BatchNorm1d(dim)
        self.skip_after_first = skip_after_first

    def forward(self, x):
        y = self.bn1(self.fc1(x))
        if self.skip_after_first:
            y = y + x
        y = self.bn2(self.fc2(F.relu(y)))
        if not self.skip_after_first:
            y = y + x
        return F.relu(y)


class _Dynamics(nn.Module):
    """Three Linear+BN layers over [state ‖ one-hot action] with a residual from the state, joined
    after bn1 (DynamicNet, model.py:60-91) or after bn3 (NewDynamicNet, model.py:93-125)."""

    def __init__(self, state_dim, action_dim, skip_after_first):
        super().__init__()
        self.state_dim = state_dim
        self.fc1, self.bn1 = nn.Linear(state_dim + action_dim, state_dim), nn.BatchNorm1d(state_dim)
        self.fc2, self.bn2 = nn.Linear(state_dim, state_dim), nn.BatchNorm1d(state_dim)
        self.fc3, self.bn3 = nn.Linear(state_dim, state_dim), nn.BatchNorm1d(state_dim)
        self.skip_after_first = skip_after_first

    def forward(self, state_action):
        state = state_action[:, :self.state_dim]
        y = self.bn1(self.fc1(state_action))
        if self.skip_after_first:
            y = y + state
        y = F.relu(self.bn2(self.fc2(F.relu(y))))
        y = self.bn3(self.fc3(y))
        if not self.skip_after_first:
            y = y + state
        return F.relu(y)


def _head(in_dim, hidden, out_dim, depth):
    layers = [nn.Linear(in_dim, hidden), nn.BatchNorm1d(hidden), nn.ReLU()]
    for _ in range(depth - 1):
        layers += [nn.Linear(hidden, hidden), nn.BatchNorm1d(hidden), nn.ReLU()]
    layers.append(nn.Linear(hidden, out_dim))
    return nn.Sequential(*layers)


class HanabiMuZeroNet(nn.Module):
    """Both reference nets; `full` selects MuZeroNetFull."""

    def __init__(self, input_size, action_space_n, reward_support_size, value_support_size,
                 full=True, support_delta=1.0, zero_heads=True):
        super().__init__()
        self.full = full
        self.action_space_n = action_space_n
        self.feature_size = FEATURE_SIZE
        self.support_delta = float(support_delta)
        f = self.feature_size
        if full:
            self.init_size, self.hidden_size = 1024, 256
            self._representation = nn.Sequential(
                nn.Linear(input_size, self.init_size), nn.BatchNorm1d(self.init_size), nn.ReLU(),
                _ResBlock(self.init_size, False),
                nn.Linear(self.init_size, f), nn.BatchNorm1d(f), nn.ReLU(), _ResBlock(f, False))
            self._dynamics_state = _Dynamics(f, action_space_n, False)
            h = self.hidden_size
            self._dynamics_reward = _head(f, h, reward_support_size, 2)
            self._prediction_actor = nn.Sequential(nn.Linear(f, h), nn.BatchNorm1d(h), nn.ReLU(),
                                                   _ResBlock(h, False), nn.Linear(h, action_space_n))
            self._prediction_value = _head(f, h, value_support_size, 2)
        else:
            self.hidden_size = 128
            h = self.hidden_size
            self._representation = nn.Sequential(nn.Linear(input_size, f), nn.BatchNorm1d(f), nn.ReLU(),
                                                 _ResBlock(f, True))
            self._dynamics_state = _Dynamics(f, action_space_n, True)
            self._dynamics_reward = _head(f, h, reward_support_size, 1)
            self._prediction_actor = _head(f, h, action_space_n, 1)
            self._prediction_value = _head(f, h, value_support_size, 1)
        if zero_heads:  # model.py:151-156, 271-276
            for head in (self._prediction_value, self._dynamics_reward, self._prediction_actor):
                nn.init.zeros_(head[-1].weight)
                nn.init.zeros_(head[-1].bias)
        half_v, half_r = (value_support_size - 1) // 2, (reward_support_size - 1) // 2
        self.register_buffer("_value_support", torch.arange(-half_v, half_v + 1, dtype=torch.float32) * support_delta,
                             persistent=False)
        self.register_buffer("_reward_support", torch.arange(-half_r, half_r + 1, dtype=torch.float32) * support_delta,
                             persistent=False)

    def randomize_heads(self, std=0.02, seed=0):
        """The reference zero-initialises all three output heads, which makes every pUCT pick a
        many-way tie.  Synthetic benchmarks re-draw them N(0, std)."""
        gen = torch.Generator().manual_seed(seed)
        for head in (self._prediction_value, self._dynamics_reward, self._prediction_actor):
            head[-1].weight.data.copy_(torch.randn(head[-1].weight.shape, generator=gen) * std)
            head[-1].bias.data.copy_(torch.randn(head[-1].bias.shape, generator=gen) * std)
        return self

    # -- the three functions -----------------------------------------------------------------------
    def representation(self, obs_history):
        return self._representation(obs_history)

    def dynamics(self, state, action):
        one_hot = torch.zeros(action.shape[0], self.action_space_n, dtype=state.dtype, device=state.device)
        one_hot.scatter_(1, action.view(-1, 1), 1.0)
        next_state = self._dynamics_state(torch.cat((state, one_hot), dim=1))
        return next_state, self._dynamics_reward(next_state)

    def prediction(self, state):
        return self._prediction_actor(state), self._prediction_value(state)

    def inverse_scalar_transform(self, logits, support):
        """core/config.py:210-232: expectation over the categorical support, then the inverse of
        h(x) = sign(x)(sqrt(|x|+1)-1) + eps*x, NaN -> 0.  Returns [N] float32."""
        delta, eps = self.support_delta, 0.001
        probs = torch.softmax(logits.float(), dim=1)
        value = (probs * support).sum(1) / delta
        out = ((torch.sqrt(1 + 4 * eps * (value.abs() + 1 + eps)) - 1) / (2 * eps)) ** 2 - 1
        out = torch.where(value < 0, -out, out) * delta  # sign[value < 0] = -1, else +1
        return torch.nan_to_num(out, nan=0.0, posinf=float("inf"), neginf=float("-inf"))

    # -- device-resident hand-off ----------------------------------------------------------------------
    @torch.no_grad()
    def initial_inference_device(self, obs):
        state = self.representation(obs)
        logits, value = self.prediction(state)
        return self.inverse_scalar_transform(value, self._value_support), logits.float(), state

    @torch.no_grad()
    def recurrent_inference_device(self, hidden_state, action):
        """-> (value [N], reward [N], policy_logits [N, A], next_hidden_state [N, F]), all CUDA."""
        state, reward = self.dynamics(hidden_state, action)
        logits, value = self.prediction(state)
        return (self.inverse_scalar_transform(value, self._value_support),
                self.inverse_scalar_transform(reward, self._reward_support), logits.float(), state)

    def recurrent_plan(self, dtype=torch.float32):
        """Folded/fused execution plan of recurrent_inference for eval mode (hanabizero_b200/plan.py),
        cached per dtype; call plan.refresh() (cheap, automatic in MCTS.run_multi) after weight updates."""
        from .plan import RecurrentPlan
        plans = self.__dict__.setdefault("_plans", {})
        if dtype not in plans:
            plans[dtype] = RecurrentPlan(self, dtype)
        return plans[dtype]

    def initial_plan(self, dtype, frame_dim, stack):
        """Folded execution plan of initial_inference for eval mode (hanabizero_b200/plan.py InitialPlan), cached."""
        from .plan import InitialPlan
        plans = self.__dict__.setdefault("_iplans", {})
        key = (dtype, int(frame_dim), int(stack))
        if key not in plans:
            plans[key] = InitialPlan(self, dtype, frame_dim, stack)
        return plans[key]

    # -- the reference's contract (core/model.py:61-84): numpy on the host in eval mode --------------------
    def initial_inference(self, obs):
        if self.training:
            state = self.representation(obs)
            logits, value = self.prediction(state)
            return NetworkOutput(value, [0.0] * obs.size(0), logits, state)
        value, logits, state = self.initial_inference_device(obs)
        return NetworkOutput(value.view(-1, 1).cpu().numpy(), [0.0] * obs.size(0), logits.cpu().numpy(),
                             state.float().cpu().numpy())

    def recurrent_inference(self, hidden_state, action):
        if self.training:
            state, reward = self.dynamics(hidden_state, action)
            logits, value = self.prediction(state)
            return NetworkOutput(value, reward, logits, state)
        value, reward, logits, state = self.recurrent_inference_device(hidden_state, action)
        return NetworkOutput(value.view(-1, 1).cpu().numpy(), reward.view(-1, 1).cpu().numpy(),
                             logits.cpu().numpy(), state.float().cpu().numpy())

    def get_weights(self):
        return {k: v.cpu() for k, v in self.state_dict().items()}

    def set_weights(self, weights):
        self.load_state_dict(weights)


def MuZeroNetFull(input_size, action_space_n, reward_support_size=201, value_support_size=201, **kw):
    return HanabiMuZeroNet(input_size, action_space_n, reward_support_size, value_support_size, full=True, **kw)


def MuZeroNet(input_size, action_space_n, reward_support_size=51, value_support_size=51, **kw):
    return HanabiMuZeroNet(input_size, action_space_n, reward_support_size, value_support_size, full=False, **kw)
