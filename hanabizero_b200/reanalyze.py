"""Reanalyze caller of the search — SURVEY.md §8f row N4.

`reanalyze_policies` is the MCTS part of BatchWorker._prepare_policy_re
(/root/reference/core/reanalyze_worker.py:308-371): the learner's target generation re-searches
`batch * (num_unroll_steps + 1)` stored positions with the current target model and turns the root visit
counts into policy targets, all-zero for positions past the end of their trajectory.  Same engine as
self-play (roots with all-zero legal masks included), nothing leaves the device until the targets do.
"""
import numpy as np
import torch

from . import _lib, cytree
from ._lib import check, ptr
from .mcts import MCTS
from .selfplay import dirichlet_noise


@torch.no_grad()
def reanalyze_policies(config, model, obs, legal_actions, policy_mask, num_unroll_steps, noises=None, mcts=None,
                       as_tensor=False, noise_seed=0, noise_step=0):
    """obs [B, obs_dim * stack] (float; host or CUDA), legal_actions [B, A] 0/1, policy_mask [B] (0 = position out
    of its trajectory), B = batch * (num_unroll_steps + 1) in trajectory-major order.  `noises` [B, A] replaces
    the Dirichlet draw (the reference multiplies its draw by the legal mask, :343; so is this one); without it the
    draw is selfplay.dirichlet_noise(seed=noise_seed, step=noise_step), reproducible on the host.
    Returns batch_policies_re: float64 [batch, num_unroll_steps + 1, A] (numpy, or a CUDA tensor with as_tensor).
    """
    lib = _lib.load()
    dev = next(model.parameters()).device
    A = int(config.action_space_size) if hasattr(config, "action_space_size") else int(np.shape(legal_actions)[-1])
    obs = cytree.as_device(obs, torch.float32, dev)
    B = obs.shape[0]
    per = int(num_unroll_steps) + 1
    if B % per:
        raise ValueError(f"{B} positions is not a multiple of num_unroll_steps + 1 = {per}")
    legal = cytree.as_device(legal_actions, torch.float32, dev, (B, A))
    mask = cytree.as_device(policy_mask, torch.uint8, dev, (B,))
    model.eval()
    amp = getattr(config, "amp_type", "none") == "torch_amp"
    with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
        _, logits, hidden = model.initial_inference_device(obs)       # :321-336 (value prefix of a root is 0)
    if noises is None:
        noises = dirichlet_noise(B, A, getattr(config, "root_dirichlet_alpha", 0.3), dev, noise_seed, noise_step)
    noises = cytree.as_device(noises, torch.float32, dev, (B, A)) * legal                       # :343
    roots = cytree.Roots(B, A, int(config.num_simulations), device=dev)
    roots.prepare(config.root_exploration_fraction, noises, torch.zeros(B, device=dev), logits.float(), legal.int())
    (mcts or MCTS(config)).run_multi(roots, model, hidden)                                       # :347
    visits = roots.get_distributions_tensor()
    out = torch.empty(B, A, dtype=torch.float64, device=dev)
    check(lib.hz_visit_policy(torch.cuda.current_stream(dev).cuda_stream, ptr(visits), ptr(mask), B, A, ptr(out), None))
    out = out.view(B // per, per, A)
    return out if as_tensor else out.cpu().numpy()
