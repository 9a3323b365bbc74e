"""Adapters giving the CUDA path (through hanabizero_b200's façade -> C ABI) the same driver API as
oracle.loader.TreeEngine / HanabiGameCPU, so parity tests run the same script on both."""
import numpy as np
import torch

from hanabizero_b200 import cytree


class GpuTreeEngine:
    def __init__(self, num, actions, sims, delta=0.006):
        self.num, self.actions, self.sims = num, actions, sims
        self.roots = cytree.Roots(num, actions, sims)
        self.mm = cytree.MinMaxStatsList(num)
        self.mm.set_delta(delta)
        self.results = None

    def prepare(self, frac, noises, rewards, logits, masks):
        if noises is None:
            self.roots.prepare_no_noise(rewards, logits, masks)
        else:
            self.roots.prepare(frac, noises, rewards, logits, masks)

    def traverse(self, pb_c_base, pb_c_init, discount):
        self.results = cytree.ResultsWrapper(self.num)
        ix, iy, la = cytree.multi_traverse(self.roots, pb_c_base, pb_c_init, discount, self.mm,
                                           self.results, as_tensor=True)
        return ix.cpu().numpy(), iy.cpu().numpy(), la.cpu().numpy()

    def backprop(self, x, discount, rewards, values, logits):
        cytree.multi_back_propagate(x, discount, rewards, values, logits, self.mm, self.results)

    def stats(self):
        visits, values = self.roots.get_stats_tensors()
        return visits.cpu().numpy(), values.cpu().numpy(), self.mm.tensor(self.roots.device).cpu().numpy()

    def root_priors(self):
        return self.roots.export(1)["root_priors"].cpu().numpy()

    def path_lens(self):
        return self.roots.export(1)["path_len"].cpu().numpy()

    def trajectories(self, max_len):
        out = np.full((self.num, max_len), -1, np.int32)
        for i, tr in enumerate(self.roots.get_trajectories()):
            out[i, :len(tr)] = tr[:max_len]
        return out

    def expanded_stats_all(self, cap):
        e = self.roots.export(cap)
        return e["reward"].cpu().numpy(), e["value_sum"].cpu().numpy(), e["visits"].cpu().numpy()


class GpuHanabiGame:
    """One game of the CUDA env with the driver API of oracle.loader.HanabiGameCPU."""

    def __init__(self, preset, seed):
        from hanabizero_b200.hanabi_env import HanabiVecEnv
        self.v = HanabiVecEnv(1, "Hanabi-Full" if preset == 0 else "Hanabi-Small", [seed])
        v = self.v
        self.enc_len, self.own_len, self.players, self.actions = v.enc_len, v.own_len, v.players, v.num_actions
        self.colors, self.ranks, self.hand_size = v.colors, v.ranks, v.hand_size
        self.local_dim, self.global_dim = v.local_dim, v.global_dim

    @staticmethod
    def _np(t):
        return t[0].cpu().numpy().astype(np.int32)

    def reset(self):
        g, l, a = self.v.reset_all()
        return self._np(g), self._np(l), self._np(a)

    def step(self, action):
        g, l, a, r, d, s = self.v.step_all(torch.tensor([action], dtype=torch.int32, device=self.v.device))
        self.v.check()
        return self._np(g), self._np(l), self._np(a), int(r[0]), bool(d[0]), int(s[0])

    def dump(self):
        return self.v.dump()[0].cpu().numpy()
