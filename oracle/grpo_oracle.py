"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's inline GRPO arithmetic.

``/root/reference/fastvideo/train_grpo_flux.py`` (``TR``) cannot be imported in this image
(accelerate / diffusers / HPSv2 / open_clip / ImageReward are absent, and the checked-in
``main`` is broken — SURVEY.md §2.1), and the arithmetic below is *inline* in
``train_one_step`` rather than in functions.  It is therefore restated here, statement by
statement, with the same torch calls in the same order:

    group_advantages          TR:439-501   (advantage_aggr / reward_aggr / no-group)
    grpo_loss                 TR:560-583   (clamped advantage, ratio, clipped surrogate, KL)
    sample_slices             TR:400-415   (which transitions are kept for training)
    gather_cat                TR:332-338
    pack / unpack / image_ids TR:80-115

Pinning: this file has no executable counterpart to diff against (the reference code is
inline), so it is pinned by (a) tests/test_oracle_pin.py::test_grpo_oracle_source_pin which
extracts the reference's own statements TR:439-501 / TR:560-583 with ``ast`` when the tree is
present and executes them against this restatement on random inputs, and (b) the sanity
values SURVEY.md §8(c) recorded from the reference.  Golden vectors produced by (a) are
committed under tests/golden/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""
from __future__ import annotations

from typing import Dict, Optional, Union

import torch


def _group_stats(group_rewards: torch.Tensor, trimmed_ratio: float):
    if trimmed_ratio > 0:                                                 # TR:451-457
        srt = torch.sort(group_rewards)[0]
        n = len(srt)
        trim = min(int(n * trimmed_ratio), n - 1)
        kept = srt[trim:]
        return kept.mean(), kept.std() + 1e-8
    return group_rewards.mean(), group_rewards.std() + 1e-8               # TR:459-460


def _per_group(rewards: torch.Tensor, num_generations: int, trimmed_ratio: float) -> torch.Tensor:
    n = len(rewards) // num_generations                                   # TR:444
    adv = torch.zeros_like(rewards)
    for g in range(n):
        lo, hi = g * num_generations, (g + 1) * num_generations
        gr = rewards[lo:hi]
        mean, std = _group_stats(gr, trimmed_ratio)
        adv[lo:hi] += (gr - mean) / std                                   # TR:461 / TR:489
    return adv


def group_advantages(rewards: Union[torch.Tensor, Dict[str, torch.Tensor]], num_generations: int,
                     reward_weights: Optional[Dict[str, float]] = None, trimmed_ratio: float = 0.0,
                     use_group: bool = True, gathered: Optional[torch.Tensor] = None) -> torch.Tensor:
    """TR:439-501.  ``rewards`` is a dict {model: [local_B] fp32} for multi_reward_mix ==
    "advantage_aggr", or a tensor for "reward_aggr".  ``gathered`` is the all-gathered reward
    vector used only by the no-group path (TR:498)."""
    if use_group:
        if isinstance(rewards, dict):
            per_model = {k: _per_group(r, num_generations, trimmed_ratio) for k, r in rewards.items()}
            merged = torch.zeros_like(next(iter(rewards.values())))       # TR:465
            for k, a in per_model.items():
                merged += a * reward_weights[k]                           # TR:467
            return merged
        return _per_group(rewards, num_generations, trimmed_ratio)
    if isinstance(rewards, dict):
        raise ValueError("multi_reward_mix 'advantage_aggr' is not supported when use_group is False.")  # TR:496
    return (rewards - gathered.mean()) / (gathered.std() + 1e-8)          # TR:498


def grpo_loss(new_logp: torch.Tensor, old_logp: torch.Tensor, advantages: torch.Tensor, clip_range: float,
              adv_clip_max: float, kl_coeff: float, grad_accum: int, n_train_steps: int):
    """TR:560-583.  Returns (loss, policy_loss, kl_loss, clip_frac); differentiable in new_logp."""
    adv = torch.clamp(advantages, -adv_clip_max, adv_clip_max)            # TR:560-564
    ratio = torch.exp(new_logp - old_logp)                                # TR:566
    unclipped = -adv * ratio
    clipped = -adv * torch.clamp(ratio, 1.0 - clip_range, 1.0 + clip_range)
    clip_frac = torch.mean((torch.abs(ratio - 1.0) > clip_range).float())  # TR:574
    policy = torch.mean(torch.maximum(unclipped, clipped)) / (grad_accum * n_train_steps)  # TR:575-577
    kl = 0.5 * torch.mean((new_logp - old_logp) ** 2) / (grad_accum * n_train_steps)       # TR:578-582
    return policy + kl_coeff * kl, policy, kl, clip_frac                  # TR:583


def sample_slices(all_latents: torch.Tensor, all_log_probs: torch.Tensor):
    """TR:406-410: transitions 0..N-2 are kept (the last one is never trained)."""
    return all_latents[:, :-1][:, :-1], all_latents[:, 1:][:, :-1], all_log_probs[:, :-1]


def gather_cat(parts):
    """TR:332-338 with the collective replaced by the list of per-rank tensors."""
    return torch.cat(list(parts), dim=0)


def pack(latents, batch, channels, height, width):
    """TR:94-99."""
    t = latents.view(batch, channels, height // 2, 2, width // 2, 2)
    return t.permute(0, 2, 4, 1, 3, 5).reshape(batch, (height // 2) * (width // 2), channels * 4)


def unpack(latents, height, width, vae_scale_factor):
    """TR:102-115."""
    b, _, c = latents.shape
    hh = 2 * (int(height) // (vae_scale_factor * 2))
    ww = 2 * (int(width) // (vae_scale_factor * 2))
    t = latents.view(b, hh // 2, ww // 2, c // 4, 2, 2).permute(0, 3, 1, 4, 2, 5)
    return t.reshape(b, c // 4, hh, ww)


def image_ids(height, width, device=None, dtype=torch.float32):
    """TR:80-91."""
    ids = torch.zeros(height, width, 3)
    ids[..., 1] = ids[..., 1] + torch.arange(height)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(width)[None, :]
    return ids.reshape(height * width, 3).to(device=device, dtype=dtype)
