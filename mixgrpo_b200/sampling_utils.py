"""Drop-in replacement for the reference's sampler / log-prob operator module.

    reference:  from fastvideo.utils.sampling_utils import (dance_grpo_step, dpm_step,
                    flow_grpo_step, run_sample_step, sd3_time_shift)         # train_grpo_flux.py:71-73
    here:       from mixgrpo_b200.sampling_utils import (...same names...)

Same names, positional/keyword signatures, return tuples, dtypes and exceptions as
``/root/reference/fastvideo/utils/sampling_utils.py`` (``SU``); each operator is ONE fused sm_100a
kernel launch (csrc/step_kernel.cuh) instead of ~35 eager launches and a host sync, and the log-prob
is differentiable w.r.t. ``model_output`` through a closed-form backward kernel
(csrc/bwd_kernels.cu).  Additions are keyword-only: ``noise=`` (explicit noise; the reference draws it
internally) and ``rounding=`` (see coefs.py).

Numerics: with ``rounding="auto"`` (default) bf16 model outputs reproduce the reference's CUDA
type-promotion roundings, so outputs match the reference on identical inputs and noise; fp32 model
outputs need no emulation.
"""
from __future__ import annotations

import math
import os
from ctypes import addressof as _addressof
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import coefs as _coefs
from . import ops as _ops
from ._cabi import SRC_DETERMINISTIC, SRC_GIVEN, SRC_NOISE, SRC_PHILOX

#: draw SDE noise inside the step kernel (counter-based Philox, no randn launch, no noise tensor) when the caller
#: supplies none; off by default so a drop-in run consumes torch's RNG stream exactly like torch.randn would
INKERNEL_NOISE = os.environ.get("MIXGRPO_INKERNEL_NOISE", "0") == "1"

#: run_sample_step: step launches only accumulate their log-prob sums, one finalize launch writes all_log_probs (same bits)
DEFER_LOG_PROBS = os.environ.get("MIXGRPO_DEFER_LOG_PROBS", "1") == "1"

#: rounding mode used when callers do not pass ``rounding=`` ("ref_cuda" | "ref_cpu" | "fp32")
DEFAULT_ROUNDING = os.environ.get("MIXGRPO_ROUNDING", "ref_cuda")


def _mode(rounding: Optional[str]) -> str:
    m = DEFAULT_ROUNDING if rounding in (None, "auto") else rounding
    if m not in _coefs.MODES:
        raise ValueError(f"rounding must be one of {_coefs.MODES} or 'auto', got {rounding!r}")
    return m


def sd3_time_shift(shift, t):
    """SU:9-10 (schedule construction; a 26-element tensor op, left to torch)."""
    return (shift * t) / (1 + (shift - 1) * t)


_scale_cache = {}


def _scale_tensor(scale: float, device: torch.device) -> torch.Tensor:
    """0-dim fp32 device tensor holding ``std_dev_t*sqrt(-dt)`` (the 5-tuple's last entry, SU:210), cached per (device,
    value): 25 distinct values per schedule, so the steady state launches no fill kernel for an output no caller reads
    (TR:149, SU:85,118).  Treat it as read-only."""
    key = (device.type, device.index, scale)
    t = _scale_cache.get(key)
    if t is None:
        if torch.cuda.is_current_stream_capturing():
            # a fill recorded into a CUDA graph only runs on replay and its memory belongs to the graph's pool: never cache it
            return torch.full((), scale, dtype=torch.float32, device=device)
        if len(_scale_cache) > 4096:
            _scale_cache.clear()
        t = torch.full((), scale, dtype=torch.float32, device=device)
        _scale_cache[key] = t
    return t


def _randn(shape, generator, device, dtype):
    """Noise source when the caller gives none (SU:189-194 ``randn_tensor``): torch's Philox on the
    target device.  A CPU generator draws on the CPU then moves, like diffusers' randn_tensor."""
    if generator is not None and generator.device.type != torch.device(device).type:
        return torch.randn(shape, generator=generator, device=generator.device, dtype=dtype).to(device)
    return torch.randn(shape, generator=generator, device=device, dtype=dtype)


def _refuse_silent_detach(model_output: torch.Tensor, what: str) -> None:
    """The reference's log-prob is differentiable on every branch (autograd through the mean).  Branches without a
    backward kernel here must not hand back a graph-less log-prob to a caller that will call ``.backward()``."""
    if torch.is_grad_enabled() and model_output.requires_grad:
        raise NotImplementedError(
            f"mixgrpo_b200: {what} has no backward kernel (the reference only trains the stored-transition log-prob, "
            "TR:149-180); call it under torch.no_grad() or pass a detached model_output")


class _DPMFirstOrderLogProb(torch.autograd.Function):
    """dpm_step(dpm_state=None, sde_solver=True) as the dpm_apply_strategy == "all" training path calls it (TR:169-180):
    first-order update with fresh noise; the log-prob is differentiable w.r.t. model_output through the mean only
    (``prev_sample.detach()``, SU:376-378).  Backward = mixgrpo_logprob_bwd family 2 on the sample this forward drew."""

    @staticmethod
    def forward(ctx, v, x, noise, k, rnd):
        xn, x0, logp, _ = _ops.fused_step(_ops.DPM, v, x, k, src=SRC_NOISE, noise=noise, order=1, round_like_torch=rnd)
        ctx.save_for_backward(v, x, xn)
        ctx.k, ctx.rnd = k, rnd
        ctx.mark_non_differentiable(xn, x0)
        return logp, xn, x0

    @staticmethod
    def backward(ctx, g_logp, _g_xn, _g_x0):
        v, x, xn = ctx.saved_tensors
        grad_v = None
        if ctx.needs_input_grad[0] and g_logp is not None:
            grad_v = _ops.logprob_backward(_ops.DPM, v, x, xn, g_logp, ctx.k, ctx.rnd)
        return grad_v, None, None, None, None


class _TransitionLogProb(torch.autograd.Function):
    """log p(x_next | x, v) with the closed-form gradient w.r.t. v (policy-update path, TR:149-168)."""

    @staticmethod
    def forward(ctx, v, x, x_next, k, family, rnd, want_mean):
        xn, x0, logp, mean = _ops.fused_step(family, v, x, k, src=SRC_GIVEN, x_next=x_next, want_x0=True,
                                             want_mean=want_mean, round_like_torch=rnd)
        ctx.save_for_backward(v, x, x_next)
        ctx.k, ctx.family, ctx.rnd = k, family, rnd
        ctx.mark_non_differentiable(x0)
        if mean is None:
            mean = x0.new_empty(0)
        ctx.mark_non_differentiable(mean)
        return logp, x0, mean

    @staticmethod
    def backward(ctx, g_logp, _g_x0, _g_mean):
        v, x, x_next = ctx.saved_tensors
        grad_v = None
        if ctx.needs_input_grad[0] and g_logp is not None:
            grad_v = _ops.logprob_backward(ctx.family, v, x, x_next, g_logp, ctx.k, ctx.rnd)
        return grad_v, None, None, None, None, None, None


def _transition_logprob(v, x, x_next, k, family, rnd, want_mean):
    """Differentiable stored-transition log-prob: the C++ ``torch::autograd::Function`` of the compiled binding (forward and
    backward never re-enter the interpreter) or, through the ctypes loader, the Python ``autograd.Function`` above."""
    tb = _ops.binding()
    if tb is not None:
        _ops.launch_count += 1 if v.shape[0] else 0                     # the forward; C++ counts its own backward launches (ops.total_launches)
        return tb.transition_logprob(v, x, x_next, _addressof(k), family, rnd, want_mean)
    return _TransitionLogProb.apply(v, x, x_next, k, family, rnd, want_mean)


def flow_grpo_step(
    model_output: torch.Tensor,
    latents: torch.Tensor,
    eta: float,
    sigmas: torch.Tensor,
    index: int,
    prev_sample: torch.Tensor,
    generator: Optional[torch.Generator] = None,
    determistic: bool = False,
    *,
    noise=None,
    rounding: Optional[str] = None,
    return_mean: bool = True,
):
    """SU:157-210.  Returns ``(prev_sample, pred_original_sample, log_prob, prev_sample_mean,
    std_dev_t*sqrt(-dt))``.  ``return_mean=False`` skips writing the (unused by every reference
    caller, TR:149, SU:85,118) mean tensor and returns None in its place.  ``noise``: a tensor (explicit noise),
    ``"philox"`` (drawn inside the kernel from ``generator`` / the device default generator) or None (torch.randn)."""
    if prev_sample is not None and generator is not None:            # SU:180-184
        raise ValueError(
            "Cannot pass both generator and prev_sample. Please make sure that either `generator` or"
            " `prev_sample` stays `None`."
        )
    _ops._require_cuda(model_output, "model_output")                  # no CPU fallback: fail before any device query
    _ops._require_cuda(latents, "latents")
    mode = _mode(rounding)
    bf16_v = model_output.dtype == torch.bfloat16
    k, scale = _coefs.flow(sigmas, index, eta, mode, bf16_v)
    rnd = bf16_v and mode != "fp32"
    # 5th output (SU:210): 0-dim device tensor like the reference; a plain float when the caller opted out
    scale_t = _scale_tensor(scale, model_output.device) if return_mean else scale
    if prev_sample is None:
        # rollout: the reference draws noise even when the step is deterministic (SU:188-195); only the
        # stochastic branch needs it here
        _refuse_silent_detach(model_output, "flow_grpo_step in rollout mode (prev_sample=None)")
        if determistic:
            xn, x0, logp, mean = _ops.fused_step(_ops.FLOW, model_output, latents, k, src=SRC_DETERMINISTIC,
                                                 want_mean=return_mean, round_like_torch=rnd)
        elif (isinstance(noise, str) and noise == "philox") or (noise is None and INKERNEL_NOISE):
            ph = _ops.philox_from_generator(model_output.device, model_output.numel(), generator)
            xn, x0, logp, mean = _ops.fused_step(_ops.FLOW, model_output, latents, k, src=SRC_PHILOX, philox=ph,
                                                 want_mean=return_mean, round_like_torch=rnd)
        else:
            if noise is None:
                noise = _randn(model_output.shape, generator, model_output.device, model_output.dtype)
            xn, x0, logp, mean = _ops.fused_step(_ops.FLOW, model_output, latents, k, src=SRC_NOISE, noise=noise,
                                                 want_mean=return_mean, round_like_torch=rnd)
        return xn, x0, logp, mean, scale_t
    if determistic:
        # prev_sample given AND determistic: the reference overwrites prev_sample with the Euler step
        # (SU:198-199) and scores that
        _refuse_silent_detach(model_output, "flow_grpo_step(determistic=True)")
        xn, x0, logp, mean = _ops.fused_step(_ops.FLOW, model_output, latents, k, src=SRC_DETERMINISTIC,
                                             want_mean=return_mean, round_like_torch=rnd)
        return xn, x0, logp, mean, scale_t
    if torch.is_grad_enabled() and model_output.requires_grad:
        logp, x0, mean = _transition_logprob(model_output, latents, prev_sample, k, _ops.FLOW, rnd, return_mean)
        return prev_sample, x0, logp, (mean if return_mean else None), scale_t
    xn, x0, logp, mean = _ops.fused_step(_ops.FLOW, model_output, latents, k, src=SRC_GIVEN, x_next=prev_sample,
                                         want_mean=return_mean, round_like_torch=rnd)
    return prev_sample, x0, logp, mean, scale_t


def dance_grpo_step(
    model_output: torch.Tensor,
    latents: torch.Tensor,
    eta: float,
    sigmas: torch.Tensor,
    index: int,
    prev_sample: torch.Tensor,
    grpo: bool,
    sde_solver: bool,
    *,
    noise: Optional[torch.Tensor] = None,
    rounding: Optional[str] = None,
):
    """SU:212-253 (DanceGRPO ``flux_step``).  ``grpo=True`` returns ``(prev_sample, pred_original_sample,
    log_prob)``; ``grpo=False`` returns ``(prev_sample_mean, pred_original_sample)``.  The log-prob omits the
    normalisation constants exactly like the reference (SU:247 is a dangling statement)."""
    _ops._require_cuda(model_output, "model_output")
    _ops._require_cuda(latents, "latents")
    mode = _mode(rounding)
    bf16_v = model_output.dtype == torch.bfloat16
    k, _std = _coefs.dance(sigmas, index, eta, mode, bf16_v)
    rnd = bf16_v and mode != "fp32"
    if prev_sample is None or not sde_solver:
        _refuse_silent_detach(model_output, "dance_grpo_step in rollout mode or with sde_solver=False")
    if not grpo:
        mean, x0, _, _ = _ops.fused_step(_ops.DANCE, model_output, latents, k, src=SRC_DETERMINISTIC, sde_solver=sde_solver,
                                         want_mean=False, want_logp=False, round_like_torch=rnd)
        return mean, x0                                               # x_next == mean on this path
    if prev_sample is None:
        if sde_solver:
            if noise is None:
                noise = torch.randn(model_output.shape, device=model_output.device, dtype=torch.float32)   # SU:238
            xn, x0, logp, _ = _ops.fused_step(_ops.DANCE, model_output, latents, k, src=SRC_NOISE, noise=noise,
                                              sde_solver=True, round_like_torch=rnd)
        else:
            xn, x0, logp, _ = _ops.fused_step(_ops.DANCE, model_output, latents, k, src=SRC_DETERMINISTIC,
                                              sde_solver=False, round_like_torch=rnd)
        return xn, x0, logp
    if torch.is_grad_enabled() and model_output.requires_grad and sde_solver:
        logp, x0, _ = _transition_logprob(model_output, latents, prev_sample, k, _ops.DANCE, rnd, False)
        return prev_sample, x0, logp
    xn, x0, logp, _ = _ops.fused_step(_ops.DANCE, model_output, latents, k, src=SRC_GIVEN, x_next=prev_sample,
                                      sde_solver=sde_solver, round_like_torch=rnd)
    return prev_sample, x0, logp


@dataclass
class DPMState:
    """SU:255-271 — history of the last ``order`` x0 predictions."""
    order: int
    model_outputs: List[torch.Tensor] = None
    lower_order_nums = 0

    def __post_init__(self):
        self.model_outputs = [None] * self.order

    def update(self, model_output: torch.Tensor):
        for i in range(self.order - 1):
            self.model_outputs[i] = self.model_outputs[i + 1]
        self.model_outputs[-1] = model_output

    def update_lower_order(self):
        if self.lower_order_nums < self.order:
            self.lower_order_nums += 1


def convert_model_output(model_output, sample, sigmas, step_index, *, rounding: Optional[str] = None) -> torch.Tensor:
    """SU:387-396: x0 = sample - sigma*model_output (one launch, no log-prob)."""
    mode = _mode(rounding)
    bf16_v = model_output.dtype == torch.bfloat16
    k = _coefs.x0_only(sigmas, step_index, mode, bf16_v)
    _, x0, _, _ = _ops.fused_step(_ops.FLOW, model_output, sample, k, src=SRC_GIVEN, x_next=sample, want_x0=True,
                                  want_logp=False, round_like_torch=bf16_v and mode != "fp32")
    return x0


def _dpm_order(args, step_index: int, n_timesteps: int, dpm_state: Optional[DPMState]) -> int:
    """Order selection of SU:308-309, 327-367."""
    lower_order_final = step_index == n_timesteps - 1
    lower_order_second = (step_index == n_timesteps - 2) and n_timesteps < 15
    if not dpm_state:
        return 1
    if args.dpm_solver_order == 1 or dpm_state.lower_order_nums < 1 or lower_order_final:
        return 1
    if args.dpm_solver_order == 2 or dpm_state.lower_order_nums < 2 or lower_order_second:
        return 2
    return 3


def dpm_step(
    args,
    model_output: torch.Tensor,
    sample: torch.Tensor,
    step_index: int,
    timesteps: list,
    sigmas: torch.Tensor,
    dpm_state: DPMState = None,
    generator=None,
    variance_noise: Optional[torch.Tensor] = None,
    sde_solver: bool = False,
    *,
    rounding: Optional[str] = None,
):
    """SU:273-385: multistep DPM-Solver(++) transition in x0-prediction form + its log-prob.
    Returns ``(prev_sample, x0_pred, log_prob)``; ``dpm_state`` is updated like the reference's."""
    _ops._require_cuda(model_output, "model_output")
    _ops._require_cuda(sample, "sample")
    mode = _mode(rounding)
    bf16_v = model_output.dtype == torch.bfloat16
    order = _dpm_order(args, step_index, len(timesteps), dpm_state)
    # history BEFORE this step's x0 is pushed: m1 = last x0, m2 = the one before (SU:486, SU:603)
    m1 = m2 = None
    if order >= 2:
        m1 = dpm_state.model_outputs[-1]
    if order == 3:
        m2 = dpm_state.model_outputs[-2]
    if order == 3 and args.dpm_algorithm_type == "dpmsolver":
        assert not sde_solver, "SDE solver is not supported for DPMSolver"          # SU:630
    k, _scale = _coefs.dpm(sigmas, step_index, order, args.dpm_algorithm_type, getattr(args, "dpm_solver_type", "midpoint"),
                           mode, bf16_v)
    rnd = bf16_v and mode != "fp32"
    if sde_solver:
        if variance_noise is None:                                    # SU:318-321
            variance_noise = _randn(model_output.shape, generator, model_output.device, torch.float32)
        src, nz = SRC_NOISE, variance_noise
    else:
        src, nz = SRC_DETERMINISTIC, None
    if torch.is_grad_enabled() and model_output.requires_grad:
        if not (sde_solver and order == 1):
            _refuse_silent_detach(model_output, f"dpm_step(order={order}, sde_solver={sde_solver})")
        logp, xn, x0 = _DPMFirstOrderLogProb.apply(model_output, sample, nz, k, rnd)      # TR:169-180
    else:
        xn, x0, logp, _ = _ops.fused_step(_ops.DPM, model_output, sample, k, src=src, noise=nz, m1=m1, m2=m2, order=order,
                                          round_like_torch=rnd)
    if dpm_state is not None:
        dpm_state.update(x0)                                          # SU:313-314
        dpm_state.update_lower_order()                                # SU:369-370
    return xn, x0, logp


def _flash_schedule(args, sigma_schedule: torch.Tensor, determistic):
    """Post-window schedule compression of MixGRPO-Flash, SU:33-54.  Host arithmetic on a 26-entry
    table; returns (new schedule on sigma_schedule.device, last SDE index)."""
    last = None
    for i in range(len(determistic) - 1, -1, -1):
        if not determistic[i]:
            last = i
            break
    n = sigma_schedule.size(0)
    num_post = int(max((n - 1 - last) * args.dpm_post_compress_ratio, 1))     # SU:44
    t_post = torch.linspace(1, 0, n)[last + 1].item()                         # SU:47
    tail = sd3_time_shift(args.shift, torch.linspace(t_post, 0, num_post).to(sigma_schedule.device))
    return torch.cat([sigma_schedule[: last + 1], tail], dim=0), last


def run_sample_step(
    args,
    z,
    progress_bar,
    sigma_schedule,
    transformer,
    encoder_hidden_states,
    pooled_prompt_embeds,
    text_ids,
    image_ids,
    grpo_sample,
    determistic,
    *,
    noises: Optional[list] = None,
    rounding: Optional[str] = None,
    decode: Optional[dict] = None,
):
    """SU:12-155 — the rollout loop: per step one transformer forward (opaque callable, bf16 autocast)
    and ONE fused sampler-step kernel that writes the next latent straight into its slot of the
    ``all_latents (B, N+1, S, 64)`` fp32 trajectory buffer (the reference appends to a list and pays a
    ``torch.stack`` copy at SU:153).  Returns ``(z, latents, all_latents, all_log_probs)``.

    ``noises`` (keyword-only, optional): per-step explicit noise tensors (entries may be None).
    ``decode`` (keyword-only, optional): ``{"height": h, "width": w, "vae_scale_factor": 8, "divisor": 0.3611, "shift": 0.1159,
    "reciprocal": True}`` — the LAST sampler step also writes the VAE's input ``unpack_latents(latents, h, w, 8) / 0.3611 +
    0.1159`` (TR:102-115, TR:286-287) as a second output, returned in ``decode["out"]`` (fp32 (B, 16, h/8, w/8)): the unpack
    launch and one read + write of the final latent disappear."""
    flash = False
    dpm_state = None
    last_sde = None
    if "dpmsolver" in args.dpm_algorithm_type:
        dpm_state = DPMState(order=args.dpm_solver_order)
        if args.dpm_apply_strategy == "post":
            assert args.sample_strategy == "progressive", "post strategy is only supported for progressive sampling"
            sigma_schedule, last_sde = _flash_schedule(args, sigma_schedule, determistic)
            progress_bar = range(0, sigma_schedule.size(0) - 1)
            flash = True

    n_steps = sigma_schedule.size(0) - 1
    B = z.shape[0]
    dev = z.device
    traj = torch.empty((B, n_steps + 1) + tuple(z.shape[1:]), dtype=torch.float32, device=dev)
    # all_latents[:, 0] = float(z): written by the first sampler step itself when that is a flow step on the vector path
    first_is_flow = args.flow_grpo_sampling and not ("dpmsolver" in args.dpm_algorithm_type and args.dpm_apply_strategy == "all")
    can_seed = first_is_flow and (n_steps > 1 or (n_steps == 1 and decode is None)) and _ops.can_seed(z, traj[:, 0])
    seed_in_step = False                         # decided at the first iteration (progress_bar may be a tqdm: no indexing)
    logps_t = torch.empty((n_steps, B), dtype=torch.float32, device=dev)   # step-major so each kernel writes a row
    # all_log_probs is read after the loop only (SU:153-155): the step launches accumulate, ONE finalize launch writes every row
    acc = _ops.DeferredLogProbs(dev, n_steps, B, z[0].numel()) if (DEFER_LOG_PROBS and B > 0) else None
    mode = _mode(rounding)
    host_sig = _coefs.host_schedule(sigma_schedule)
    guidance = torch.tensor([3.5], device=dev, dtype=torch.bfloat16)
    txt_ids = text_ids.repeat(encoder_hidden_states.shape[1], 1)
    pred_original = None
    cur = z
    steps_done = 0
    dec_last = None
    if decode is not None:
        vsf = int(decode.get("vae_scale_factor", 8))
        hh, ww = 2 * (int(decode["height"]) // (vsf * 2)), 2 * (int(decode["width"]) // (vsf * 2))
        decode["out"] = torch.empty((B, z.shape[-1] // 4, hh, ww), dtype=torch.float32, device=dev)
        dec_last = {"out": decode["out"], "divisor": decode.get("divisor", 1.0), "shift": decode.get("shift", 0.0),
                    "from_x0": bool(args.drop_last_sample), "reciprocal": bool(decode.get("reciprocal", False))}
    for i in progress_bar:
        if steps_done == 0:
            seed_in_step = can_seed and i == 0
            if not seed_in_step:
                _ops.cast_rows(z, traj[:, 0])
        dec = dec_last if i == n_steps - 1 else None
        timestep_value = int(host_sig[i] * 1000)                                  # SU:63-65 without the sync
        timesteps = torch.full([encoder_hidden_states.shape[0]], timestep_value, device=dev, dtype=torch.long)
        transformer.eval()
        with torch.autocast("cuda", torch.bfloat16):
            pred = transformer(
                hidden_states=cur,
                encoder_hidden_states=encoder_hidden_states,
                timestep=timesteps / 1000,
                guidance=guidance,
                txt_ids=txt_ids,
                pooled_projections=pooled_prompt_embeds,
                img_ids=image_ids,
                joint_attention_kwargs=None,
                return_dict=False,
            )[0]
        x = traj[:, i]
        out = traj[:, i + 1]
        nz = noises[i] if noises is not None else None
        bf16_v = pred.dtype == torch.bfloat16
        rnd = bf16_v and mode != "fp32"
        use_dpm = "dpmsolver" in args.dpm_algorithm_type and (args.dpm_apply_strategy == "all" or (flash and i > last_sde))
        if use_dpm:
            sde = (not determistic[i]) if args.dpm_apply_strategy == "all" else False
            order = _dpm_order(args, i, n_steps, dpm_state)
            m1 = dpm_state.model_outputs[-1] if order >= 2 else None
            m2 = dpm_state.model_outputs[-2] if order == 3 else None
            k, _ = _coefs.dpm(sigma_schedule, i, order, args.dpm_algorithm_type, args.dpm_solver_type, mode, bf16_v)
            if sde and nz is None:
                nz = torch.randn(pred.shape, device=dev, dtype=torch.float32)
            _, pred_original, lp, _ = _ops.fused_step(_ops.DPM, pred, x, k, src=SRC_NOISE if sde else SRC_DETERMINISTIC,
                                                      noise=nz if sde else None, m1=m1, m2=m2, order=order, out_x_next=out,
                                                      out_logp=logps_t[i], round_like_torch=rnd, decode=dec, defer=acc.slot(i, k) if acc is not None else None)
            dpm_state.update(pred_original)
            dpm_state.update_lower_order()
        elif args.flow_grpo_sampling:
            k, _ = _coefs.flow(sigma_schedule, i, args.eta, mode, bf16_v)
            seed0 = seed_in_step and i == 0
            if determistic[i]:
                _, pred_original, lp, _ = _ops.fused_step(_ops.FLOW, pred, z if seed0 else x, k, src=SRC_DETERMINISTIC, out_x_next=out,
                                                          out_logp=logps_t[i], round_like_torch=rnd, decode=dec, defer=acc.slot(i, k) if acc is not None else None,
                                                          seed_out=x if seed0 else None)
            else:
                if nz is None:
                    nz = torch.randn(pred.shape, device=dev, dtype=pred.dtype)
                _, pred_original, lp, _ = _ops.fused_step(_ops.FLOW, pred, z if seed0 else x, k, src=SRC_NOISE, noise=nz, out_x_next=out,
                                                          out_logp=logps_t[i], round_like_torch=rnd, decode=dec, defer=acc.slot(i, k) if acc is not None else None,
                                                          seed_out=x if seed0 else None)
            if flash:                                                               # SU:116-117, 127
                dpm_state.update(pred_original)
                dpm_state.update_lower_order()
        else:
            k, _ = _coefs.dance(sigma_schedule, i, args.eta, mode, bf16_v)
            if determistic[i]:
                _, pred_original, lp, _ = _ops.fused_step(_ops.DANCE, pred, x, k, src=SRC_DETERMINISTIC, sde_solver=False,
                                                          out_x_next=out, out_logp=logps_t[i], round_like_torch=rnd, decode=dec, defer=acc.slot(i, k) if acc is not None else None)
            else:
                if nz is None:
                    nz = torch.randn(pred.shape, device=dev, dtype=torch.float32)
                _, pred_original, lp, _ = _ops.fused_step(_ops.DANCE, pred, x, k, src=SRC_NOISE, noise=nz, sde_solver=True,
                                                          out_x_next=out, out_logp=logps_t[i], round_like_torch=rnd, decode=dec, defer=acc.slot(i, k) if acc is not None else None)
        cur = out
        steps_done += 1
    if steps_done == 0:
        _ops.cast_rows(z, traj[:, 0])
    if acc is not None:
        acc.finalize(logps_t)

    z_out = traj[:, n_steps]
    latents = pred_original if args.drop_last_sample else z_out.to(pred_original.dtype)   # SU:149-152
    return z_out, latents, traj, logps_t.t()


def _sigma_to_alpha_sigma_t(sigma):
    """SU:641-644."""
    return 1 - sigma, sigma


__all__ = ["sd3_time_shift", "run_sample_step", "flow_grpo_step", "dance_grpo_step", "DPMState", "dpm_step",
           "convert_model_output"]
