"""The hot-path portions of the reference trainer, re-expressed as library functions over the fused kernels.

``TR`` = ``/root/reference/fastvideo/train_grpo_flux.py``.  The reference's ``main``/argparse/FSDP/optimizer/VAE/reward
models are out of scope; what is here is exactly the code between those pieces:

* ``grpo_one_step``            TR:118-181  — same signature; DiT forward with grad + differentiable log-prob
* ``sample_reference_model``   TR:184-329  — rollout orchestration, but the whole prompt group runs as ONE batch
                                             (the reference loops 12 batch-1 rollouts, TR:213,231); VAE decode + reward
                                             models are a caller-supplied callback
* ``train_window``             TR:503-615  — the (sample, window step) policy-update loop: DiT forward with grad, fused
                                             log-prob + loss forward, fused backward, ``pred.backward(grad)``; logging
                                             scalars stay on the device
* ``prepare_latent_image_ids`` TR:80-91

The ``args`` namespace is the reference's (same field names).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Sequence

import torch

from . import grpo as _grpo
from . import ops as _ops
from . import rollout as _rollout
from .sampling_utils import dance_grpo_step, dpm_step, flow_grpo_step, run_sample_step, sd3_time_shift


def prepare_latent_image_ids(batch_size, height, width, device, dtype):
    """TR:80-91 (host construction of a (h*w, 3) index table; one small H2D copy)."""
    ids = torch.zeros(height, width, 3)
    ids[..., 1] = ids[..., 1] + torch.arange(height)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(width)[None, :]
    return ids.reshape(height * width, 3).to(device=device, dtype=dtype)


def _forward(transformer, latents, encoder_hidden_states, pooled_prompt_embeds, text_ids, image_ids, timesteps):
    """The DiT call of TR:133-144 / SU:67-82 (opaque: stays on stock PyTorch / cuBLAS)."""
    with torch.autocast("cuda", torch.bfloat16):
        return transformer(
            hidden_states=latents,
            encoder_hidden_states=encoder_hidden_states,
            timestep=timesteps / 1000,
            guidance=torch.tensor([3.5], device=latents.device, dtype=torch.bfloat16),
            txt_ids=text_ids[:1].repeat(encoder_hidden_states.shape[1], 1),   # (L, 3); FLUX takes 2-D ids shared by the batch
            pooled_projections=pooled_prompt_embeds,
            img_ids=image_ids,
            joint_attention_kwargs=None,
            return_dict=False,
        )[0]


def grpo_one_step(args, latents, pre_latents, encoder_hidden_states, pooled_prompt_embeds, text_ids, image_ids, transformer,
                  timesteps, i, sigma_schedule):
    """TR:118-181: new log-prob of the stored transition ``latents -> pre_latents`` at step ``i`` (differentiable)."""
    transformer.train()
    pred = _forward(transformer, latents, encoder_hidden_states, pooled_prompt_embeds, text_ids,
                    image_ids.squeeze(0) if image_ids.dim() == 3 else image_ids, timesteps)
    if args.dpm_algorithm_type == "null" or ("dpmsolver" in args.dpm_algorithm_type and args.dpm_apply_strategy == "post"):
        if args.flow_grpo_sampling:
            _, _, log_prob, _, _ = flow_grpo_step(model_output=pred, latents=latents.to(torch.float32), eta=args.eta, sigmas=sigma_schedule,
                                                  index=i, prev_sample=pre_latents.to(torch.float32), determistic=False, return_mean=False)
        else:
            _, _, log_prob = dance_grpo_step(pred, latents.to(torch.float32), args.eta, sigma_schedule, i,
                                             prev_sample=pre_latents.to(torch.float32), grpo=True, sde_solver=True)
    else:   # dpm "all": fresh noise, the stored next latent is ignored (TR:169-180)
        _, _, log_prob = dpm_step(args, model_output=pred, sample=latents.to(torch.float32), step_index=i, timesteps=sigma_schedule[:-1],
                                  dpm_state=None, sde_solver=True, sigmas=sigma_schedule)
    return log_prob


def sample_reference_model(args, device, transformer, encoder_hidden_states, pooled_prompt_embeds, text_ids,
                           decode_and_score: Callable[[torch.Tensor], object], timesteps_train: Sequence[int], *,
                           input_latents: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None,
                           noises=None):
    """TR:184-329 for a whole batch of ``B = encoder_hidden_states.shape[0]`` samples at once.

    ``decode_and_score(latents (B, S, 64) fp32) -> rewards`` stands for unpack + VAE decode + reward models
    (TR:284-316; out of scope) and returns a tensor ``[B]`` (reward_aggr) or a dict of tensors (advantage_aggr).
    Returns ``(rewards, all_latents, all_log_probs, sigma_schedule, image_ids)`` like TR:329."""
    w, h = args.w, args.h
    sigma_schedule = torch.linspace(1, 0, args.sampling_steps + 1).to(device)
    sigma_schedule = sd3_time_shift(args.shift, sigma_schedule)                     # TR:200-202
    B = encoder_hidden_states.shape[0]
    latent_w, latent_h = w // 8, h // 8
    if input_latents is None:
        shape = (1 if getattr(args, "init_same_noise", False) else B, 16, latent_h, latent_w)
        input_latents = torch.randn(shape, device=device, dtype=torch.bfloat16, generator=generator)   # TR:223-243
        if input_latents.shape[0] == 1 and B > 1:
            input_latents = input_latents.expand(B, -1, -1, -1).contiguous()
    z = _ops.pack_latents(input_latents, B, 16, latent_h, latent_w)                  # TR:244
    image_ids = prepare_latent_image_ids(B, latent_h // 2, latent_w // 2, device, torch.bfloat16)
    if getattr(args, "training_strategy", "part") == "part":                         # TR:251-256
        determistic = _rollout.window_mask(args.sampling_steps, timesteps_train, "part")
    else:
        determistic = [False] * args.sampling_steps
    with torch.no_grad():
        _, latents, all_latents, all_log_probs = run_sample_step(
            args, z, range(args.sampling_steps), sigma_schedule, transformer, encoder_hidden_states, pooled_prompt_embeds,
            text_ids[:1], image_ids, True, determistic, noises=noises)
        rewards = decode_and_score(latents)
    return rewards, all_latents, all_log_probs, sigma_schedule, image_ids


def train_window(args, transformer, samples: Dict, advantages: torch.Tensor, sigma_schedule: torch.Tensor,
                 train_timesteps: Sequence[int], encoder_hidden_states, pooled_prompt_embeds, text_ids, image_ids, *,
                 micro_batch: int = 1, on_accumulated: Optional[Callable[[int], None]] = None,
                 stats_rows: Optional[torch.Tensor] = None) -> torch.Tensor:
    """TR:536-615 without host syncs: for every micro-batch of samples and every window step — DiT forward with grad on
    the stored latent, fused log-prob + clipped-ratio loss (forward and backward: two launches), ``pred.backward(grad)``.

    ``samples`` is ``rollout.make_samples(...)`` (TR:406-415).  ``on_accumulated(i)`` is called after sample ``i`` when
    ``(i + 1) % gradient_accumulation_steps == 0`` — the place of clip_grad_norm_/optimizer.step() (TR:605-609).
    Returns ``stats_rows`` ([B, 4]: per-sample sums of loss, policy_loss, kl_loss, clip_frac over the window)."""
    B = samples["latents"].shape[0]
    dev = samples["latents"].device
    if stats_rows is None:
        stats_rows = torch.zeros(B, 4, dtype=torch.float32, device=dev)
    cfg = _rollout.SamplerConfig(sampling_steps=args.sampling_steps, eta=args.eta, shift=args.shift,
                                 flow_grpo_sampling=args.flow_grpo_sampling)
    transformer.train()
    T = len(train_timesteps)
    for lo in range(0, B, micro_batch):
        hi = min(lo + micro_batch, B)
        for t in train_timesteps:
            lat = samples["latents"][lo:hi, t]
            pred = _forward(transformer, lat, encoder_hidden_states[lo:hi], pooled_prompt_embeds[lo:hi], text_ids[lo:hi], image_ids,
                            samples["timesteps"][lo:hi, t])
            _, _, grad = _rollout.policy_update(pred, lat, samples["next_latents"][lo:hi, t], samples["log_probs"][lo:hi, t],
                                                advantages[lo:hi], sigma_schedule, t, cfg, clip_range=args.clip_range,
                                                adv_clip_max=args.adv_clip_max, kl_coeff=args.kl_coeff,
                                                gradient_accumulation_steps=args.gradient_accumulation_steps,
                                                num_train_timesteps=T, stats_rows=stats_rows[lo:hi])
            pred.backward(grad)                                                      # TR:585 continues into the DiT
        if on_accumulated is not None:
            for i in range(lo, hi):
                if (i + 1) % args.gradient_accumulation_steps == 0:
                    on_accumulated(i)
    return stats_rows
