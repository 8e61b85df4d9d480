"""Reward → group-relative advantage → clipped-ratio GRPO loss, as library functions.

The reference has no API for this: the arithmetic is inline in ``train_one_step``
(``/root/reference/fastvideo/train_grpo_flux.py`` = ``TR``: gather TR:332-338 / TR:417-425, advantages
TR:439-501, loss TR:560-583, logging reductions TR:586-600).  The functions below keep the reference's
argument names (``num_generations``, ``trimmed_ratio``, ``reward_weights``, ``clip_range``,
``adv_clip_max``, ``kl_coeff``, ``gradient_accumulation_steps``) so the inline block can be replaced
line for line.

Each function is one kernel launch (csrc/grpo_kernels.cu) and never syncs the host.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.distributed as dist

from . import ops as _ops

RewardsLike = Union[torch.Tensor, Dict[str, torch.Tensor]]


def stack_rewards(rewards: RewardsLike) -> Tuple[torch.Tensor, Optional[List[str]]]:
    """dict {model_name: [local_B]} (advantage_aggr, TR:417-422) or tensor (reward_aggr, TR:423-425)
    → fp32 matrix [n_models, local_B] and the model-name order."""
    if isinstance(rewards, dict):
        names = list(rewards.keys())
        return torch.stack([rewards[k].to(torch.float32).reshape(-1) for k in names], dim=0), names
    r = rewards.to(torch.float32)
    return (r.reshape(1, -1) if r.dim() == 1 else r), None


def gather_tensor(tensor: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Drop-in for TR:332-338: all-gather a ``[local_B, ...]`` tensor and concatenate along dim 0 in rank order (one
    ``all_gather_into_tensor`` instead of a list all-gather + ``torch.cat``); identity without a process group."""
    if not dist.is_available() or not dist.is_initialized():
        return tensor
    world = dist.get_world_size(group)
    t = tensor.contiguous()
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out


def gather_rewards(rewards: RewardsLike, group: Optional[dist.ProcessGroup] = None) -> RewardsLike:
    """Replacement for the per-model ``gather_tensor`` list all-gathers (TR:332-338, TR:417-425): ONE
    ``all_gather_into_tensor`` of the ``[n_models, local_B]`` matrix.  Returns the same container type
    with every entry of length ``world * local_B`` in rank order (``torch.cat`` order of TR:338)."""
    mat, names = stack_rewards(rewards)
    if not dist.is_available() or not dist.is_initialized():
        gathered = mat
    else:
        world = dist.get_world_size(group)
        mat = mat.contiguous()
        buf = torch.empty((world * mat.shape[0], mat.shape[1]), dtype=mat.dtype, device=mat.device)   # rank-major rows
        dist.all_gather_into_tensor(buf, mat, group=group)
        gathered = buf.view(world, mat.shape[0], mat.shape[1]).permute(1, 0, 2).reshape(mat.shape[0], -1)   # [n_models, world*local_B]
    if names is None:
        return gathered.reshape(-1) if (isinstance(rewards, torch.Tensor) and rewards.dim() == 1) else gathered
    return {k: gathered[i] for i, k in enumerate(names)}


_weight_cache: Dict[Tuple, torch.Tensor] = {}


def _weights_on(device: torch.device, w: Tuple[float, ...]) -> torch.Tensor:
    """Reward weights as a cached device vector (one H2D copy per distinct weight tuple, not per step)."""
    key = (device.type, device.index, w)
    t = _weight_cache.get(key)
    if t is None:
        t = torch.tensor(w, dtype=torch.float32).to(device)
        if not torch.cuda.is_current_stream_capturing():                   # a copy recorded into a CUDA graph only runs on replay
            _weight_cache[key] = t
    return t


def compute_group_advantages(rewards: RewardsLike, num_generations: int,
                             reward_weights: Optional[Union[Dict[str, float], Sequence[float]]] = None,
                             trimmed_ratio: float = 0.0, use_group: bool = True,
                             gathered_rewards: Optional[torch.Tensor] = None) -> torch.Tensor:
    """TR:439-501.

    * ``rewards`` dict + ``reward_weights``  → ``multi_reward_mix == "advantage_aggr"`` (TR:441-468)
    * ``rewards`` tensor                     → ``"reward_aggr"`` (TR:470-491)
    * ``use_group=False``                    → global normalisation with the statistics of
      ``gathered_rewards`` (TR:495-499); a dict raises the reference's ValueError.

    Groups are consecutive runs of ``num_generations`` rank-local samples (TR:444-450)."""
    mat, names = stack_rewards(rewards)
    if not use_group:
        if names is not None:
            raise ValueError("multi_reward_mix 'advantage_aggr' is not supported when use_group is False.")   # TR:496
        stat = gathered_rewards if gathered_rewards is not None else mat.reshape(-1)
        return _ops.group_advantages(mat, None, 1, 0, use_group=False, stat_rewards=stat)
    weights = None
    if torch.is_tensor(reward_weights):                                     # already a device vector [n_models]
        weights = reward_weights
    elif names is not None or (reward_weights is not None and mat.shape[0] > 1):
        if reward_weights is None:
            raise ValueError("reward_weights is required for multi-reward (advantage_aggr) rewards")
        w = tuple(float(reward_weights[k]) for k in names) if isinstance(reward_weights, dict) else tuple(float(x) for x in reward_weights)
        weights = _weights_on(mat.device, w)
    trim = 0
    if trimmed_ratio > 0:                                                   # TR:451-454
        trim = min(int(num_generations * trimmed_ratio), num_generations - 1)
    return _ops.group_advantages(mat, weights, num_generations, trim, use_group=True)


class _GRPOLoss(torch.autograd.Function):
    """Forward + closed-form backward of TR:560-583 in one launch; backward is a scale by grad_out."""

    @staticmethod
    def forward(ctx, new_logp, old_logp, advantages, clip_range, adv_clip_max, kl_coeff, denom, stats_accum):
        stats, grad = _ops.grpo_loss_fwd_bwd(new_logp, old_logp, advantages, clip_range, adv_clip_max, kl_coeff, denom,
                                             want_grad=True, stats_accum=stats_accum)
        ctx.save_for_backward(grad)
        ctx.shape = new_logp.shape
        loss, policy, kl, clip_frac = stats[0], stats[1], stats[2], stats[3]
        ctx.mark_non_differentiable(policy, kl, clip_frac)
        return loss, policy, kl, clip_frac

    @staticmethod
    def backward(ctx, g_loss, *_):
        (grad,) = ctx.saved_tensors
        return (grad * g_loss).reshape(ctx.shape), None, None, None, None, None, None, None


def grpo_loss(new_log_probs: torch.Tensor, old_log_probs: torch.Tensor, advantages: torch.Tensor, clip_range: float,
              adv_clip_max: float, kl_coeff: float, gradient_accumulation_steps: int, num_train_timesteps: int,
              stats_accum: Optional[torch.Tensor] = None):
    """TR:560-583 → ``(loss, policy_loss, kl_loss, clip_frac)`` (0-dim fp32 device tensors).  ``loss`` is
    differentiable w.r.t. ``new_log_probs``.  ``stats_accum`` ([4] fp32 device tensor, optional) receives
    ``+= (loss, policy, kl, clip_frac)`` on the device — see ``reduce_step_stats``."""
    denom = float(gradient_accumulation_steps * num_train_timesteps)        # TR:576
    return _GRPOLoss.apply(new_log_probs, old_log_probs, advantages, clip_range, adv_clip_max, kl_coeff, denom, stats_accum)


def grpo_loss_and_grad(new_log_probs, old_log_probs, advantages, clip_range, adv_clip_max, kl_coeff,
                       gradient_accumulation_steps, num_train_timesteps, stats_accum=None):
    """Non-autograd form: returns ``(stats[4], dloss/dnew_log_probs [B])`` in one launch; the gradient
    vector feeds ``ops.logprob_backward`` directly (no host sync, TR:585)."""
    denom = float(gradient_accumulation_steps * num_train_timesteps)
    return _ops.grpo_loss_fwd_bwd(new_log_probs, old_log_probs, advantages, clip_range, adv_clip_max, kl_coeff, denom,
                                  want_grad=True, stats_accum=stats_accum)


def reduce_step_stats(stats_accum: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """ONE ``all_reduce(AVG)`` of the device-accumulated ``[loss, policy, kl, clip_frac]`` sums per
    ``train_one_step`` — replaces the four all-reduce + ``.item()`` pairs the reference issues per
    (sample, window step) at TR:586-600 (averaging commutes with the sum).  Returns the reduced tensor;
    the caller decides when to read it on the host."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if stats_accum.device.type == "cuda":
            dist.all_reduce(stats_accum, op=dist.ReduceOp.AVG, group=group)
        else:                                                               # gloo (CPU tests) has no AVG
            dist.all_reduce(stats_accum, op=dist.ReduceOp.SUM, group=group)
            stats_accum /= dist.get_world_size(group)
    return stats_accum


# ----------------------------------------------------------------------------------------------- partitioning
def partition_prompts(n_prompts: int, rank: int, world: int, drop_last: bool = False) -> List[int]:
    """Which prompts (= prompt groups: each prompt is repeated ``num_generations`` times, TR:368-384) a rank owns:
    ``rank, rank + world, ...`` — the order of ``DistributedSampler(shuffle=False)`` (TR:737-749), padded by wrapping
    around unless ``drop_last``.  Groups never cross ranks, so group statistics stay rank-local (TR:443-461) and the
    rollout needs no collective."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    if drop_last:
        total = (n_prompts // world) * world
        idx = list(range(total))
    else:
        total = ((n_prompts + world - 1) // world) * world
        idx = list(range(n_prompts))
        idx += idx[: total - n_prompts] if n_prompts else []
    return idx[rank:total:world]


def split_group_slice(values: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Extended mode: a prompt group split ACROSS ranks (e.g. group 24 over 8 GPUs = 3 samples each).
    ``values`` is a rank-major gathered vector ``[world * local_B]`` (what ``gather_tensor`` returns); the rank's own
    slice is returned."""
    local = values.shape[-1] // world
    return values[..., rank * local:(rank + 1) * local]


def compute_group_advantages_split(rewards: RewardsLike, num_generations: int,
                                   reward_weights: Optional[Union[Dict[str, float], Sequence[float]]] = None,
                                   trimmed_ratio: float = 0.0, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Extended mode: groups of ``num_generations`` samples laid out consecutively in the rank-major GATHERED order, so a
    group may span several ranks.  One all-gather of the reward matrix, the group kernel over the gathered matrix
    (``world * local_B`` scalars per model), then this rank's slice.  With whole groups per rank it equals
    ``compute_group_advantages`` on the local rewards."""
    gathered = gather_rewards(rewards, group)
    adv = compute_group_advantages(gathered, num_generations, reward_weights, trimmed_ratio)
    if dist.is_available() and dist.is_initialized():
        return split_group_slice(adv, dist.get_rank(group), dist.get_world_size(group)).contiguous()
    return adv
