"""The hot-path portions of the reference trainer, re-expressed as library functions over the fused kernels.

``TR`` = ``/root/reference/fastvideo/train_grpo_flux.py``.  The reference's ``main``/argparse/FSDP/optimizer/VAE/reward
models are out of scope; what is here is exactly the code between those pieces:

* ``grpo_one_step``            TR:118-181  — same signature; DiT forward with grad + differentiable log-prob
* ``sample_reference_model``   TR:184-329  — rollout orchestration, but the whole prompt group runs as ONE batch
                                             (the reference loops 12 batch-1 rollouts, TR:213,231); VAE decode + reward
                                             models are a caller-supplied callback
* ``train_one_step``           TR:341-640  — the whole hot path of one training step as one call (rollout → reward exchange +
                                             advantages → re-ranging → policy-update loop → logging reduction); optional fused
                                             NVLink exchange (``peer.PeerExchange``)
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
                           noises=None, vae_input: bool = False):
    """TR:184-329 for a whole batch of ``B = encoder_hidden_states.shape[0]`` samples at once.

    ``decode_and_score(latents (B, S, 64) fp32) -> rewards`` stands for unpack + VAE decode + reward models
    (TR:284-316; out of scope) and returns a tensor ``[B]`` (reward_aggr) or a dict of tensors (advantage_aggr).
    ``vae_input=True``: the callback receives what TR:286-287 hands to ``vae.decode`` instead — ``unpack_latents(latents, h, w,
    8) / 0.3611 + 0.1159`` as fp32 ``(B, 16, h/8, w/8)`` — written by the LAST sampler step as a second output (no unpack
    launch, no extra pass over the final latent; bit-identical to the two torch ops on CUDA).
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
    dec = {"height": h, "width": w, "vae_scale_factor": 8, "divisor": 0.3611, "shift": 0.1159, "reciprocal": True} if vae_input else None
    with torch.no_grad():
        _, latents, all_latents, all_log_probs = run_sample_step(
            args, z, range(args.sampling_steps), sigma_schedule, transformer, encoder_hidden_states, pooled_prompt_embeds,
            text_ids[:1], image_ids, True, determistic, noises=noises, decode=dec)
        rewards = decode_and_score(dec["out"] if vae_input else latents)
    return rewards, all_latents, all_log_probs, sigma_schedule, image_ids


def train_window(args, transformer, samples: Dict, advantages: torch.Tensor, sigma_schedule: torch.Tensor,
                 train_timesteps: Sequence[int], encoder_hidden_states, pooled_prompt_embeds, text_ids, image_ids, *,
                 micro_batch: int = 1, on_accumulated: Optional[Callable[[int], None]] = None,
                 stats_rows: Optional[torch.Tensor] = None, order: Optional[Sequence[int]] = None,
                 perms: Optional[Sequence[Sequence[int]]] = None) -> torch.Tensor:
    """TR:536-615 without host syncs: for every micro-batch of samples and every window step — DiT forward with grad on
    the stored latent, fused log-prob + clipped-ratio loss (forward and backward: two launches), ``pred.backward(grad)``.

    ``samples`` is ``rollout.make_samples(...)`` (TR:406-415).  ``on_accumulated(i)`` is called after the ``i``-th trained
    sample when ``(i + 1) % gradient_accumulation_steps == 0`` — the place of clip_grad_norm_/optimizer.step() (TR:605-609).
    ``order``: which samples to train and in which order (``balance_pos_neg`` re-ranging, TR:528-535); default all, in
    place.  ``perms`` (host lists, ``training_strategy == "all"``, TR:503-509): ``samples`` columns were permuted per
    sample, so column ``t`` of sample ``i`` is sampler step ``perms[i][t]`` — needs ``micro_batch == 1``.
    Returns ``stats_rows`` ([B, 4]: per-sample sums of loss, policy_loss, kl_loss, clip_frac over the window)."""
    B = samples["latents"].shape[0]
    dev = samples["latents"].device
    if stats_rows is None:
        stats_rows = torch.zeros(B, 4, dtype=torch.float32, device=dev)
    cfg = _rollout.SamplerConfig(sampling_steps=args.sampling_steps, eta=args.eta, shift=args.shift,
                                 flow_grpo_sampling=args.flow_grpo_sampling)
    transformer.train()
    T = len(train_timesteps)
    if perms is not None and micro_batch != 1:
        raise ValueError("per-sample step permutations (training_strategy 'all') need micro_batch == 1")
    order = list(range(B)) if order is None else [int(i) for i in order]
    contiguous = order == list(range(B))
    # dpm_apply_strategy == "all" trains a DIFFERENT objective (TR:169-180): dpm_step draws fresh noise and scores its own
    # sample, ignoring the stored next latent — so it cannot use the fused stored-transition kernels.  It goes through
    # grpo_one_step (differentiable first-order DPM log-prob) + grpo_loss + autograd, one launch each.
    dpm_all = "dpmsolver" in getattr(args, "dpm_algorithm_type", "null") and getattr(args, "dpm_apply_strategy", "post") == "all"
    for lo in range(0, len(order), micro_batch):
        ids = order[lo:lo + micro_batch]
        # a run of consecutive samples is a view; a re-ranged micro-batch is gathered (one small index copy per tensor)
        sel = slice(ids[0], ids[-1] + 1) if (contiguous or len(ids) == 1) else torch.tensor(ids, device=dev)
        for t in train_timesteps:
            step = int(perms[ids[0]][t]) if perms is not None else t               # TR:553
            lat = samples["latents"][sel, t]
            rows = stats_rows[sel] if isinstance(sel, slice) else torch.zeros(len(ids), 4, dtype=torch.float32, device=dev)
            if dpm_all:
                for i in ids:                                                        # the reference's B == 1 evaluation (TR:542-585)
                    one = slice(i, i + 1)
                    new_lp = grpo_one_step(args, samples["latents"][one, t], samples["next_latents"][one, t], encoder_hidden_states[one],
                                           pooled_prompt_embeds[one], text_ids[one], image_ids, transformer, samples["timesteps"][one, t],
                                           step, sigma_schedule)
                    out = _grpo.grpo_loss(new_lp, samples["log_probs"][one, t], advantages[one], args.clip_range, args.adv_clip_max,
                                          args.kl_coeff, args.gradient_accumulation_steps, T)
                    out[0].backward()
                    stats_rows[i] += torch.stack([o.detach() for o in out])
                continue
            pred = _forward(transformer, lat, encoder_hidden_states[sel], pooled_prompt_embeds[sel], text_ids[sel], image_ids,
                            samples["timesteps"][sel, t])
            _, _, grad = _rollout.policy_update(pred, lat, samples["next_latents"][sel, t], samples["log_probs"][sel, t],
                                                advantages[sel], sigma_schedule, step, cfg, clip_range=args.clip_range,
                                                adv_clip_max=args.adv_clip_max, kl_coeff=args.kl_coeff,
                                                gradient_accumulation_steps=args.gradient_accumulation_steps,
                                                num_train_timesteps=T, stats_rows=rows)
            if not isinstance(sel, slice):
                stats_rows.index_add_(0, sel, rows)
            pred.backward(grad)                                                      # TR:585 continues into the DiT
        if on_accumulated is not None:
            for i in range(lo, lo + len(ids)):
                if (i + 1) % args.gradient_accumulation_steps == 0:
                    on_accumulated(i)
    return stats_rows


def train_one_step(args, device, transformer, decode_and_score: Callable[[torch.Tensor], object], timesteps_train: Sequence[int],
                   reward_weights, encoder_hidden_states, pooled_prompt_embeds, text_ids, *, exchange=None,
                   on_accumulated: Optional[Callable[[int], None]] = None, micro_batch: int = 1,
                   input_latents: Optional[torch.Tensor] = None, noises=None, generator: Optional[torch.Generator] = None,
                   rng=None, split_groups: bool = False, vae_input: bool = False):
    """The hot path of the reference's ``train_one_step`` (TR:341-640) as one call: prompt repetition (TR:369-384), batched
    rollout (TR:386-399), sample bookkeeping (TR:400-415), reward exchange + group-relative advantages (TR:417-501), step
    permutation / positive-negative re-ranging (TR:503-535), the (sample, window step) policy-update loop (TR:536-615) and
    the logging reduction (TR:586-600, 617-625).  What is NOT here is what SURVEY §8 leaves on stock PyTorch: the data
    loader (``next(loader)`` → pass its tensors), VAE decode + reward models (``decode_and_score``), and
    ``clip_grad_norm_`` / ``optimizer.step()`` / ``lr_scheduler.step()`` (``on_accumulated``, called where TR:605-609 runs them).

    ``exchange``: a :class:`mixgrpo_b200.peer.PeerExchange` — reward gather + advantages and the logging all-reduce then
    run as one fused NVLink kernel each; without it NCCL (or nothing, single process) is used.
    ``split_groups`` (SURVEY §8e extended mode): a prompt group is spread over several ranks — ``num_generations`` counts the
    samples of a group across ranks, consecutive in rank-major order, and the statistics come from the gathered rewards;
    each rank repeats its prompt ``num_generations // world`` times.
    ``vae_input``: ``decode_and_score`` receives the VAE-ready tensor written by the last sampler step (see
    ``sample_reference_model``).
    Returns ``(stats [4] = total_loss, policy_total_loss, kl_total_loss, total_clip_frac — rank-averaged device tensor,
    gathered_reward_res (per-model mean of the gathered rewards, device tensors), samples, advantages)``; nothing syncs the
    host except the optional ``advantage_rerange_strategy`` (which needs the advantages' signs, like the reference)."""
    G = int(args.num_generations)
    world = torch.distributed.get_world_size() if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 1
    local_G = G // world if split_groups else G
    if split_groups and (G % world != 0 or not getattr(args, "use_group", True)):
        raise ValueError("split_groups needs use_group and num_generations divisible by the world size")
    if getattr(args, "use_group", True):                                             # TR:369-384
        rep = lambda t: None if t is None else torch.repeat_interleave(t, local_G, dim=0)   # noqa: E731
        encoder_hidden_states, pooled_prompt_embeds, text_ids = rep(encoder_hidden_states), rep(pooled_prompt_embeds), rep(text_ids)
    rewards, all_latents, all_log_probs, sigma_schedule, image_ids = sample_reference_model(
        args, device, transformer, encoder_hidden_states, pooled_prompt_embeds, text_ids, decode_and_score, timesteps_train,
        input_latents=input_latents, generator=generator, noises=noises, vae_input=vae_input)
    samples = _rollout.make_samples(all_latents, all_log_probs, sigma_schedule, args.sampling_steps)        # TR:400-415
    rewards = {k: v.to(torch.float32) for k, v in rewards.items()} if isinstance(rewards, dict) else rewards.to(torch.float32)
    use_group = getattr(args, "use_group", True)
    trimmed = float(getattr(args, "trimmed_ratio", 0.0) or 0.0)
    if exchange is not None:                                                         # TR:417-501 in ONE launch
        advantages, gathered = exchange.gather_advantages(rewards, G, reward_weights, trimmed_ratio=trimmed,
                                                          mode=("split" if split_groups else "local") if use_group else "global")
    elif split_groups:
        gathered = _grpo.gather_rewards(rewards)
        advantages = _grpo.compute_group_advantages_split(rewards, G, reward_weights, trimmed_ratio=trimmed)
    else:
        gathered = _grpo.gather_rewards(rewards)
        advantages = _grpo.compute_group_advantages(rewards, G, reward_weights, trimmed_ratio=trimmed, use_group=use_group,
                                                    gathered_rewards=None if use_group else gathered)
    samples["rewards"], samples["advantages"] = rewards, advantages
    B = all_latents.shape[0]
    perms_host, order = None, None
    strategy = getattr(args, "training_strategy", "part")
    if strategy == "all":                                                            # TR:503-509, TR:518-525
        perms = _rollout.shuffle_timesteps(samples, generator=None)
        perms_host = perms.tolist()
        n_cols = samples["timesteps"].shape[1]
        frozen = int(getattr(args, "frozen_init_timesteps", 0) or 0)
        train_timesteps = range(frozen) if frozen > 0 else range(int(n_cols * float(getattr(args, "timestep_fraction", 1.0))))
        micro_batch = 1
    else:
        train_timesteps = list(timesteps_train)
        rerange = getattr(args, "advantage_rerange_strategy", "null")
        if rerange in ("random", "balance"):                                         # TR:527-535 (host decision, like the reference)
            tagged = [{"index": i, "advantages": a} for i, a in enumerate(advantages.tolist())]
            order = [d["index"] for d in _rollout.balance_pos_neg(tagged, use_random=rerange == "random", rng=rng)]
        elif rerange != "null":
            raise ValueError(f"advantage_rerange_strategy {rerange} is not supported.")
    rows = train_window(args, transformer, samples, advantages, sigma_schedule, train_timesteps, encoder_hidden_states,
                        pooled_prompt_embeds, text_ids, image_ids, micro_batch=micro_batch, on_accumulated=on_accumulated,
                        order=order, perms=perms_host)
    stats = rows.sum(dim=0)                                                          # TR:588-600 add these up one .item() at a time
    stats = exchange.allreduce_stats(stats) if exchange is not None else _grpo.reduce_step_stats(stats)
    if isinstance(gathered, dict):                                                   # TR:617-625
        gathered_res = {k: v.mean() for k, v in gathered.items()}
    else:
        gathered_res = gathered.mean()
    return stats, gathered_res, samples, advantages
