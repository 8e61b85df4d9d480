"""B200-native rollout and policy-update drivers built on the fused kernels.

The reference rolls a prompt group out as 12 sequential batch-1 loops (TR:213,231; TR =
/root/reference/fastvideo/train_grpo_flux.py) and updates the policy one (sample, window step) at a time
with ~15 scalar kernels, 4 all-reduces and 5 ``.item()`` syncs each (TR:536-600).  Here:

* ``rollout``            — the whole group in ONE batch: per step one model call + one fused kernel that
                           reads ``all_latents[:, i]`` and writes ``all_latents[:, i+1]`` in place.
* ``make_samples``       — TR:400-415 (views, no copies).
* ``policy_update``      — fused log-prob + loss forward, fused loss-grad + log-prob backward: two launches, zero
                           host syncs; returns ``grad_model_output`` for ``pred.backward(grad)``.
* ``balance_pos_neg`` / ``shuffle_timesteps`` — sample bookkeeping of TR:503-532 and
                           fastvideo/models/reward_model/utils.py:18-48.
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import coefs as _coefs
from . import grpo as _grpo
from . import ops as _ops
from ._cabi import SRC_DETERMINISTIC, SRC_GIVEN, SRC_NOISE, SRC_PHILOX
from .sampling_utils import DPMState, _dpm_order, _flash_schedule, _mode


@dataclass
class SamplerConfig:
    """The ``args`` fields the reference's operators read (SU:29-152, TR:145-180); defaults are the
    shipped MixGRPO configuration (scripts/finetune/finetune_flux_grpo_MixGRPO.sh:48-79)."""
    sampling_steps: int = 25
    eta: float = 0.7
    shift: float = 3.0
    flow_grpo_sampling: bool = True
    dpm_algorithm_type: str = "null"          # "null" | "dpmsolver" | "dpmsolver++"
    dpm_apply_strategy: str = "post"          # "post" | "all"
    dpm_post_compress_ratio: float = 0.4
    dpm_solver_order: int = 2
    dpm_solver_type: str = "midpoint"
    sample_strategy: str = "progressive"
    drop_last_sample: bool = False
    rounding: str = "auto"
    inkernel_noise: bool = False             # draw SDE noise inside the step kernel when `noises[i]` is None (no randn launch)
    # programmatic dependent launch: the model outputs / noise handed to a sampler step are NOT written by the launch issued
    # immediately before it on the stream (true when `model` returns precomputed tensors or when its last kernel is not the
    # producer of v; after a real DiT forward the attribute is inert anyway) -> v / noise loads overlap the previous step's drain
    pdl_early_v: bool = False
    # all_log_probs[:, i] of a DETERMINISTIC step is dead data in the reference itself: train_one_step only ever indexes the
    # SDE-window steps (TR:536-553), and with training_strategy "all" every step is an SDE step.  True (default) computes it
    # anyway, exactly like the reference (SU:201-208 runs on every step); False skips the reduction on ODE steps (their CTAs
    # retire right after their stores: -0.5 us per launch at (12,4096,64)) and leaves NaN in those columns.
    ode_log_probs: bool = True
    # The rollout's log-probs are only read once the rollout is over (SU:153-155): its step launches accumulate their per-sample
    # sums with a fire-and-forget reduction and ONE finalize launch writes all_log_probs (mixgrpo_logp_finalize) — same packed
    # integer sums, same bits, but no CTA waits an L2 round trip for a returning atomic before it retires.
    defer_log_probs: bool = True


def sigma_schedule(sampling_steps: int, shift: float, device=None) -> torch.Tensor:
    """TR:200-202: ``sd3_time_shift(shift, linspace(1, 0, N+1))``."""
    t = torch.linspace(1, 0, sampling_steps + 1)
    if device is not None:
        t = t.to(device)
    return (shift * t) / (1 + (shift - 1) * t)


def window_mask(sampling_steps: int, timesteps_train: Sequence[int], training_strategy: str = "part") -> List[bool]:
    """TR:251-256: ``determistic[i]`` is False exactly on the SDE-window steps."""
    if training_strategy == "all":
        return [False] * sampling_steps
    det = [True] * sampling_steps
    for i in timesteps_train:
        det[i] = False
    return det


def rollout(model: Callable[[torch.Tensor, float, int], torch.Tensor], z: torch.Tensor, sigmas: torch.Tensor,
            determistic: Sequence[bool], cfg: SamplerConfig, noises: Optional[Sequence[Optional[torch.Tensor]]] = None,
            want_x0: bool = False, generator: Optional[torch.Generator] = None, philox_state: Optional[_ops.PhiloxState] = None,
            decode: Optional[dict] = None):
    """Batched equivalent of SU:61-155.  ``model(latents, sigma_float, step) -> model_output`` stands for the
    DiT forward (SU:62-82).  Returns ``(z_final, latents, all_latents (B,N+1,...), all_log_probs (B,N), sigmas)``
    where ``sigmas`` is the schedule actually used (rebuilt in Flash "post" mode, SU:33-54).

    ``philox_state``: with ``cfg.inkernel_noise``, draw from this graph-safe device state (required under CUDA-graph
    capture; advanced once at the end of the rollout) instead of a torch generator.
    ``decode``: ``{"height": h, "width": w, "vae_scale_factor": 8, "divisor": 0.3611, "shift": 0.1159, "reciprocal": False}`` —
    the LAST sampler step also writes the VAE's input ``unpack_latents(latents, h, w, 8) / divisor + shift`` (TR:286-287)
    as a second output; it is returned in ``decode["out"]`` (fp32 (B, C, H', W')).  Saves the unpack launch and one
    read + write of the final latent."""
    mode = _mode(cfg.rounding)
    dpm_state = None
    last_sde = None
    flash = False
    if "dpmsolver" in cfg.dpm_algorithm_type:
        dpm_state = DPMState(order=cfg.dpm_solver_order)
        if cfg.dpm_apply_strategy == "post":
            assert cfg.sample_strategy == "progressive", "post strategy is only supported for progressive sampling"
            sigmas, last_sde = _flash_schedule(cfg, sigmas, determistic)
            flash = True
    n_steps = sigmas.size(0) - 1
    B, dev = z.shape[0], z.device
    traj = torch.empty((B, n_steps + 1) + tuple(z.shape[1:]), dtype=torch.float32, device=dev)
    # all_latents[:, 0] = float(z) (SU:26, SU:153): written by the FIRST sampler step itself when that is a flow-family step on
    # the vector path (it then reads the bf16 z directly), else by its own cast launch
    first_is_flow = cfg.flow_grpo_sampling and not ("dpmsolver" in cfg.dpm_algorithm_type and cfg.dpm_apply_strategy == "all")
    seed_in_step = (first_is_flow and (n_steps > 1 or (n_steps == 1 and decode is None)) and _ops.can_seed(z, traj[:, 0]))
    if not seed_in_step:
        _ops.cast_rows(z, traj[:, 0])
    acc = _ops.DeferredLogProbs(dev, n_steps, B, z[0].numel()) if (cfg.defer_log_probs and B > 0) else None
    if cfg.ode_log_probs or acc is not None:
        logps_t = torch.empty((n_steps, B), dtype=torch.float32, device=dev)                     # finalize also fills skipped rows with NaN
    else:
        logps_t = torch.full((n_steps, B), float("nan"), dtype=torch.float32, device=dev)       # skipped columns stay poisoned
    host_sig = _coefs.host_schedule(sigmas).tolist()
    x0 = None
    need_x0_last = cfg.drop_last_sample or want_x0
    early = 1 if cfg.pdl_early_v else 0
    dec_last = None
    if decode is not None:
        vsf = int(decode.get("vae_scale_factor", 8))
        hh, ww = 2 * (int(decode["height"]) // (vsf * 2)), 2 * (int(decode["width"]) // (vsf * 2))
        ch = z.shape[-1] // 4
        decode["out"] = torch.empty((B, ch, hh, ww), dtype=torch.float32, device=dev)
        dec_last = {"out": decode["out"], "divisor": decode.get("divisor", 1.0), "shift": decode.get("shift", 0.0),
                    "from_x0": bool(cfg.drop_last_sample), "reciprocal": bool(decode.get("reciprocal", False))}
    for i in range(n_steps):
        x, out = traj[:, i], traj[:, i + 1]
        dec = dec_last if i == n_steps - 1 else None
        v = model(z if i == 0 else x, host_sig[i], i)
        bf16_v = v.dtype == torch.bfloat16
        rnd = bf16_v and mode != "fp32"
        nz = noises[i] if noises is not None else None
        use_dpm = dpm_state is not None and (cfg.dpm_apply_strategy == "all" or (flash and i > last_sde))
        # x0 is only materialised when something downstream reads it: the DPM history, or the caller
        keep_x0 = (dpm_state is not None) or (need_x0_last and i == n_steps - 1)
        if use_dpm:
            sde = (not determistic[i]) if cfg.dpm_apply_strategy == "all" else False
            order = _dpm_order(cfg, i, n_steps, dpm_state)
            m1 = dpm_state.model_outputs[-1] if order >= 2 else None
            m2 = dpm_state.model_outputs[-2] if order == 3 else None
            k, _ = _coefs.dpm(sigmas, i, order, cfg.dpm_algorithm_type, cfg.dpm_solver_type, mode, bf16_v)
            if sde and nz is None:
                nz = torch.randn(v.shape, device=dev, dtype=torch.float32)
            lp_on = sde or cfg.ode_log_probs
            _, x0, _, _ = _ops.fused_step(_ops.DPM, v, x, k, src=SRC_NOISE if sde else SRC_DETERMINISTIC,
                                          noise=nz if sde else None, m1=m1, m2=m2, order=order, out_x_next=out,
                                          out_logp=logps_t[i] if lp_on else None, want_logp=lp_on, want_x0=True, round_like_torch=rnd,
                                          early=early, decode=dec, defer=acc.slot(i, k) if (acc is not None and lp_on) else None)
            dpm_state.update(x0)
            dpm_state.update_lower_order()
        else:
            fam = _ops.FLOW if cfg.flow_grpo_sampling else _ops.DANCE
            k, _ = (_coefs.flow if cfg.flow_grpo_sampling else _coefs.dance)(sigmas, i, cfg.eta, mode, bf16_v)
            ph = None
            if determistic[i]:
                src, nz = SRC_DETERMINISTIC, None
            elif nz is None and cfg.inkernel_noise:
                src = SRC_PHILOX
                ph = philox_state.take(v.numel()) if philox_state is not None else _ops.philox_from_generator(dev, v.numel(), generator)
            else:
                src = SRC_NOISE
                if nz is None:
                    nz = torch.randn(v.shape, device=dev, dtype=v.dtype if cfg.flow_grpo_sampling else torch.float32, generator=generator)
            lp_on = (not determistic[i]) or cfg.ode_log_probs
            seed0 = seed_in_step and i == 0
            _, x0, _, _ = _ops.fused_step(fam, v, z if seed0 else x, k, src=src, noise=nz, philox=ph, sde_solver=not determistic[i], out_x_next=out,
                                          out_logp=logps_t[i] if lp_on else None, want_logp=lp_on, want_x0=keep_x0, round_like_torch=rnd,
                                          early=early, decode=dec, defer=acc.slot(i, k) if (acc is not None and lp_on) else None,
                                          seed_out=x if seed0 else None)
            if flash and cfg.flow_grpo_sampling:               # SU:116-117, SU:127
                dpm_state.update(x0)
                dpm_state.update_lower_order()
    if acc is not None:
        acc.finalize(logps_t)                                                # ONE launch: every step's log-probs
    if philox_state is not None:
        philox_state.advance()                                               # the next rollout (or graph replay) draws fresh noise
    z_final = traj[:, n_steps]
    latents = x0 if cfg.drop_last_sample else z_final                        # SU:149-152
    return z_final, latents, traj, logps_t.t(), sigmas


def timestep_values(sigmas: torch.Tensor, sampling_steps: Optional[int] = None) -> List[int]:
    """``[int(sigma * 1000) for sigma in sigma_schedule][:sampling_steps]`` (TR:401, SU:63-65): the product is taken in
    fp32 like the reference's 0-dim tensor op — a python-double product differs by one for many (N, shift) pairs — then
    truncated.  One host pass over the cached schedule, no device sync per entry."""
    host = _coefs.host_schedule(sigmas)
    vals = (host * 1000).to(torch.int64).tolist()                            # fp32 multiply, truncation toward zero = int()
    return vals if sampling_steps is None else vals[:sampling_steps]


def make_samples(all_latents: torch.Tensor, all_log_probs: torch.Tensor, sigmas: torch.Tensor, sampling_steps: int) -> Dict:
    """TR:400-415: the last transition is never trained, so N-1 transitions remain (views only)."""
    B = all_latents.shape[0]
    tvals = timestep_values(sigmas, sampling_steps)                          # TR:401
    timesteps = torch.tensor([tvals] * B, dtype=torch.long).to(all_latents.device, non_blocking=True)
    return {
        "timesteps": timesteps[:, :-1],
        "latents": all_latents[:, :-1][:, :-1],
        "next_latents": all_latents[:, 1:][:, :-1],
        "log_probs": all_log_probs[:, :-1],
    }


def shuffle_timesteps(samples: Dict, generator: Optional[torch.Generator] = None):
    """TR:503-509 (training_strategy == "all"): independent random step order per sample.  Returns perms."""
    B, T = samples["timesteps"].shape
    dev = samples["timesteps"].device
    perms = torch.stack([torch.randperm(T, generator=generator) for _ in range(B)]).to(dev)
    rows = torch.arange(B, device=dev)[:, None]
    for key in ("timesteps", "latents", "next_latents", "log_probs"):
        samples[key] = samples[key][rows, perms]
    return perms


def balance_pos_neg(samples: List[dict], use_random: bool = False, rng: Optional[random.Random] = None) -> List[dict]:
    """fastvideo/models/reward_model/utils.py:18-48 — interleave positive- and negative-advantage samples.
    Samples with advantage exactly 0 are dropped, like the reference.  ``rng`` defaults to the global
    ``random`` module the reference uses."""
    rng = rng or random
    if use_random:
        return rng.sample(samples, len(samples))
    signs = [float(s["advantages"]) if not torch.is_tensor(s["advantages"]) else s["advantages"].item() for s in samples]
    pos = [s for s, a in zip(samples, signs) if a > 0]
    neg = [s for s, a in zip(samples, signs) if a < 0]
    pos = rng.sample(pos, len(pos))
    neg = rng.sample(neg, len(neg))
    small, large = (pos, neg) if len(pos) < len(neg) else (neg, pos)
    mixed = []
    for a, b in zip(small, large):
        mixed += [a, b]
    mixed.extend(large[len(small):])
    return mixed


def policy_update(v: torch.Tensor, latents: torch.Tensor, next_latents: torch.Tensor, old_log_probs: torch.Tensor,
                  advantages: torch.Tensor, sigmas: torch.Tensor, index: int, cfg: SamplerConfig, *, clip_range: float,
                  adv_clip_max: float, kl_coeff: float, gradient_accumulation_steps: int, num_train_timesteps: int,
                  stats_rows: Optional[torch.Tensor] = None, accumulate: bool = True, single_pass: bool = False):
    """One (samples, window step) policy update, TR:542-585 without autograd: given the model output ``v`` for
    the stored ``latents`` it returns ``(stats_rows, new_log_probs [B], grad_v)`` where ``grad_v`` is dloss/dv — hand it
    to ``v.backward(grad_v)`` to continue into the DiT.  ``stats_rows`` ([B,4] fp32, optional) accumulates each sample's
    (loss, policy_loss, kl_loss, clip_frac) (``accumulate=False`` overwrites instead); ``stats_rows.sum(0)`` is what
    TR:588-600 adds up.  Two launches, no sync.  ``single_pass=True`` uses mixgrpo_policy_step instead when the batch's
    residuals fit on chip (up to ~8 M latent scalars, e.g. (24, 4096, 64)): ONE launch that reads every latent byte once
    (12 instead of 22 B/elem from HBM, same bits).  It is opt-in because on B200 the pair is as fast or faster: the
    backward's re-reads hit the 126 MB L2, while the single pass pays a grid-wide dependency (profiles/r01_policy_step.md)."""
    if "dpmsolver" in cfg.dpm_algorithm_type and cfg.dpm_apply_strategy == "all":
        raise NotImplementedError("dpm_apply_strategy='all' trains dpm_step's own fresh-noise transition (TR:169-180), not the stored "
                                  "one: use trainer.train_window / trainer.grpo_one_step (autograd path)")
    mode = _mode(cfg.rounding)
    bf16_v = v.dtype == torch.bfloat16
    rnd = bf16_v and mode != "fp32"
    if cfg.flow_grpo_sampling:
        fam, (k, _) = _ops.FLOW, _coefs.flow(sigmas, index, cfg.eta, mode, bf16_v)
    else:
        fam, (k, _) = _ops.DANCE, _coefs.dance(sigmas, index, cfg.eta, mode, bf16_v)
    vd = v.detach()
    # The reference evaluates the loss one sample at a time (B == 1 per call, TR:536-585) and lets autograd accumulate
    # sum_i loss_i / (GA*T); the fused kernels do exactly that per sample: forward = log-prob + loss terms into the
    # sample's stats row, backward = dL/dlogp evaluated in place + closed-form chain.  Two launches, no loss kernel.
    denom = float(gradient_accumulation_steps * num_train_timesteps)
    if single_pass:
        res = _ops.policy_step(fam, vd, latents, next_latents, k, old_log_probs, advantages, clip_range, adv_clip_max, kl_coeff, denom,
                               stats_rows=stats_rows, round_like_torch=rnd, accumulate=accumulate)
        if res is not None:
            return stats_rows, res[0], res[1]
    new_lp = _ops.policy_forward(fam, vd, latents, next_latents, k, old_log_probs, advantages, clip_range, adv_clip_max, kl_coeff,
                                 denom, stats_rows=stats_rows, round_like_torch=rnd, accumulate=accumulate)
    grad_v = _ops.policy_backward(fam, vd, latents, next_latents, new_lp, k, old_log_probs, advantages, clip_range, adv_clip_max,
                                  kl_coeff, denom, round_like_torch=rnd, early_loads=True)   # launched right after the forward
    return stats_rows, new_lp, grad_v



def policy_update_window(vs: Sequence[torch.Tensor], all_latents: torch.Tensor, steps: Sequence[int], old_log_probs: torch.Tensor,
                         advantages: torch.Tensor, sigmas: torch.Tensor, cfg: SamplerConfig, *, clip_range: float, adv_clip_max: float,
                         kl_coeff: float, gradient_accumulation_steps: int, stats_rows: Optional[torch.Tensor] = None,
                         accumulate: bool = True, early_loads: bool = False):
    """The whole SDE window's policy update in TWO launches (mixgrpo_policy_fwd_multi / _bwd_multi): for window step
    ``t = steps[j]`` with model output ``vs[j]``, the stored transition ``all_latents[:, t] -> all_latents[:, t+1]`` and old
    log-probs ``old_log_probs[:, t]`` (TR:536-585 — the reference walks the (sample, step) pairs one by one; given their model
    outputs they are independent).  Returns ``(stats_rows [J,B,4], new_log_probs [J,B], [grad_v_j])``; hand ``grad_v_j`` to
    ``vs[j].backward``.  Bit-identical to ``policy_update`` per step; falls back to it for ragged / unaligned tensors or
    more than 8 steps.  ``early_loads``: no input was written by the launch immediately before this call on the stream."""
    if "dpmsolver" in cfg.dpm_algorithm_type and cfg.dpm_apply_strategy == "all":
        raise NotImplementedError("dpm_apply_strategy='all' trains dpm_step's own transition: use trainer.train_window")
    J = len(steps)
    B, dev = all_latents.shape[0], all_latents.device
    mode = _mode(cfg.rounding)
    bf16_v = vs[0].dtype == torch.bfloat16
    rnd = bf16_v and mode != "fp32"
    fam = _ops.FLOW if cfg.flow_grpo_sampling else _ops.DANCE
    table = _coefs.flow if cfg.flow_grpo_sampling else _coefs.dance
    ks = [table(sigmas, int(t), cfg.eta, mode, bf16_v)[0] for t in steps]
    denom = float(gradient_accumulation_steps * J)                           # TR:576
    if stats_rows is None:
        stats_rows = torch.zeros((J, B, 4), dtype=torch.float32, device=dev)
        accumulate = False
    vd = [v.detach() for v in vs]
    xs = [all_latents[:, int(t)] for t in steps]
    xns = [all_latents[:, int(t) + 1] for t in steps]
    olds = [old_log_probs[:, int(t)] for t in steps]
    new_lp = None
    if J <= _ops.POLICY_MAX_ITEMS:
        new_lp = _ops.policy_forward_multi(fam, vd, xs, xns, ks, olds, advantages, clip_range, adv_clip_max, kl_coeff, denom,
                                           stats_rows=[stats_rows[j] for j in range(J)], round_like_torch=rnd, accumulate=accumulate,
                                           early_loads=early_loads)
    if new_lp is not None:
        grads = _ops.policy_backward_multi(fam, vd, xs, xns, new_lp, ks, olds, advantages, clip_range, adv_clip_max, kl_coeff, denom,
                                           round_like_torch=rnd, early_loads=True)        # right after the forward, which wrote only [J,B] floats
        if grads is not None:
            return stats_rows, new_lp, grads
    new_lp = torch.empty((J, B), dtype=torch.float32, device=dev)
    grads = []
    for j, t in enumerate(steps):
        _, lp, g = policy_update(vs[j], xs[j], xns[j], olds[j], advantages, sigmas, int(t), cfg, clip_range=clip_range, adv_clip_max=adv_clip_max,
                                 kl_coeff=kl_coeff, gradient_accumulation_steps=gradient_accumulation_steps, num_train_timesteps=J,
                                 stats_rows=stats_rows[j], accumulate=accumulate)
        new_lp[j].copy_(lp)
        grads.append(g)
    return stats_rows, new_lp, grads
