"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference sampler/log-prob operators.

This is the parity oracle for the CUDA path in ``mixgrpo_b200``.  It restates, with
plain PyTorch tensor ops executed in the reference's operator order (so that every
fp32/bf16 rounding point is reproduced), the algorithms of
``/root/reference/fastvideo/utils/sampling_utils.py`` (abbreviated ``SU`` below):

    sd3_time_shift            SU:9-10
    flow_step                 SU:157-210   (flow_grpo_step — MixGRPO default)
    dance_step                SU:212-253   (dance_grpo_step — DanceGRPO flux_step)
    x0_from_velocity          SU:387-396   (convert_model_output)
    dpm_first_order           SU:398-447
    dpm_second_order          SU:449-561
    dpm_third_order           SU:563-639
    dpm_step / History        SU:255-385
    rollout                   SU:12-155    (run_sample_step incl. the Flash "post" schedule)

Differences from the reference are deliberate and limited to the *interface*:
noise is always an explicit argument (the reference draws it from a generator), and
``args`` namespaces are replaced by keyword arguments.

Pinning: tests/test_oracle_pin.py executes the unmodified reference (oracle/ref_loader.py)
on the same inputs and requires bit-identical tensors on CPU; tests/golden/*.npz hold
outputs of the *reference itself* (tools/make_golden.py) for boxes without the tree.
The functions are device-agnostic: run on CUDA tensors they reproduce the reference's
CUDA type-promotion behaviour (0-dim CUDA scalars are cast to the tensor dtype).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product path never does.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch

LOG_SQRT_2PI_ARG = 2 * math.pi


def sd3_time_shift(shift, t):
    """SU:9-10."""
    return (shift * t) / (1 + (shift - 1) * t)


def _gauss_logp(x_next, mean, scale):
    """Mean over non-batch dims of the diagonal-Gaussian log-density (SU:201-208, SU:376-383).

    ``scale`` is a 0-dim tensor; the three terms are evaluated per element in the
    reference's order before the mean.
    """
    per_elem = (
        -((x_next.detach() - mean) ** 2) / (2 * (scale ** 2))
        - torch.log(scale)
        - torch.log(torch.sqrt(2 * torch.as_tensor(math.pi)))
    )
    return per_elem.mean(dim=tuple(range(1, per_elem.ndim)))


def flow_step(v, x, eta, sigmas, index, x_next=None, noise=None, determistic=False):
    """Flow-GRPO SDE/ODE transition + log-prob, SU:157-210.

    v: model output (bf16 or fp32), x: latents (fp32), x_next: stored next latents for
    the policy-update path or None for rollout (then ``noise`` — dtype of v — is required,
    SU:188-195).  Returns (x_next, x0, logp, mean, scale) like SU:210.
    """
    dev = v.device
    s_cur = sigmas[index].to(dev)
    s_nxt = sigmas[index + 1].to(dev)
    s_guard = sigmas[1].item()                                   # SU:172
    dt = s_nxt - s_cur                                           # negative, SU:173
    x0 = x - s_cur * v                                           # SU:175
    std = torch.sqrt(s_cur / (1 - torch.where(s_cur == 1, s_guard, s_cur))) * eta   # SU:177
    mean = x * (1 + std ** 2 / (2 * s_cur) * dt) + v * (1 + std ** 2 * (1 - s_cur) / (2 * s_cur)) * dt  # SU:186
    scale = std * torch.sqrt(-1 * dt)
    if x_next is None:
        assert noise is not None and noise.dtype == v.dtype
        x_next = mean + std * torch.sqrt(-1 * dt) * noise        # SU:195
    if determistic:
        x_next = x + dt * v                                      # SU:198-199
    logp = _gauss_logp(x_next, mean, scale)
    return x_next, x0, logp, mean, scale


def gauss_logp_fp64(x_next, mean, scale, constants=True):
    """fp64 "truth" of the log-prob reduction (SU:201-208) for GIVEN fp32 tensors ``x_next`` / ``mean`` and the fp32
    scalar ``scale``: every element widened to double before it is squared and averaged.  Used for error budgeting
    (SURVEY §7 stage 1): the reference's fp32 ``mean()`` and the kernel's fixed-point reduction are both compared to it."""
    s = torch.as_tensor(scale, dtype=torch.float32).double()
    d = x_next.double() - mean.double()
    q = (-(d ** 2) / (2 * s ** 2)).reshape(d.shape[0], -1).mean(dim=1)
    if constants:
        q = q - torch.log(s) - math.log(math.sqrt(2 * math.pi))
    return q


def dance_step(v, x, eta, sigmas, index, x_next=None, noise=None, grpo=True, sde_solver=True):
    """DanceGRPO flux_step, SU:212-253.  ``noise`` (fp32, SU:238 randn_like) is used
    only when rolling out with sde_solver.  The log-prob constants are NOT subtracted
    (SU:247 is a dangling expression statement) — preserved."""
    s_cur = sigmas[index]
    ds = sigmas[index + 1] - s_cur
    mean = x + ds * v                                            # SU:224
    x0 = x - s_cur * v                                           # SU:226
    delta = s_cur - sigmas[index + 1]
    std = eta * math.sqrt(delta)                                 # python float, SU:229
    if sde_solver:
        score = -(x - x0 * (1 - s_cur)) / s_cur ** 2             # SU:232
        drift = -0.5 * eta ** 2 * score                          # SU:233
        mean = mean + drift * ds                                 # SU:234
    if grpo and x_next is None:
        if sde_solver:
            assert noise is not None
            x_next = mean + noise.to(mean.dtype) * std           # SU:238
        else:
            x_next = mean
    if not grpo:
        return mean, x0
    logp = -((x_next.detach().to(torch.float32) - mean.to(torch.float32)) ** 2) / (2 * (std ** 2))  # SU:244-246
    logp = logp.mean(dim=tuple(range(1, logp.ndim)))
    return x_next, x0, logp


def x0_from_velocity(v, x, sigmas, step_index):
    """SU:387-396."""
    return x - sigmas[step_index] * v


class History:
    """Multistep history of x0 predictions (SU:255-271, DPMState)."""

    def __init__(self, order: int):
        self.order = order
        self.model_outputs: List[Optional[torch.Tensor]] = [None] * order
        self.lower_order_nums = 0

    def update(self, x0):
        for k in range(self.order - 1):
            self.model_outputs[k] = self.model_outputs[k + 1]
        self.model_outputs[-1] = x0

    def update_lower_order(self):
        if self.lower_order_nums < self.order:
            self.lower_order_nums += 1


def _lam(sig):
    return torch.log(1 - sig) - torch.log(sig)                   # SU:424-425, 641-644


def dpm_first_order(algo, m0, sigmas, i, x, noise, sde):
    """SU:398-447.  Returns (x_next, mean, std, dt_sqrt)."""
    sg_t, sg_s = sigmas[i + 1], sigmas[i]
    al_t, al_s = 1 - sg_t, 1 - sg_s
    h = _lam(sg_t) - _lam(sg_s)
    if algo == "dpmsolver++":
        mean = (sg_t / sg_s * torch.exp(-h)) * x + (al_t * (1 - torch.exp(-2.0 * h))) * m0
        std, dt_sqrt = sg_t, torch.sqrt(1.0 - torch.exp(-2 * h))
        if sde:
            assert noise is not None
            out = mean + std * dt_sqrt * noise
        else:
            out = (sg_t / sg_s) * x - (al_t * (torch.exp(-h) - 1.0)) * m0
    elif algo == "dpmsolver":
        mean = (al_t / al_s) * x - 2.0 * (sg_t * (torch.exp(h) - 1.0)) * m0
        std, dt_sqrt = sg_t, torch.sqrt(torch.exp(2 * h) - 1.0)
        if sde:
            assert noise is not None
            out = mean + std * dt_sqrt * noise
        else:
            out = (al_t / al_s) * x - (sg_t * (torch.exp(h) - 1.0)) * m0
    else:
        raise ValueError(algo)
    return out, mean, std, dt_sqrt


def dpm_second_order(algo, solver_type, hist, sigmas, i, x, noise, sde):
    """SU:449-561."""
    sg_t, sg_0, sg_1 = sigmas[i + 1], sigmas[i], sigmas[i - 1]
    al_t, al_0 = 1 - sg_t, 1 - sg_0
    l_t, l_0, l_1 = _lam(sg_t), _lam(sg_0), _lam(sg_1)
    m0, m1 = hist[-1], hist[-2]
    h, h_0 = l_t - l_0, l_0 - l_1
    r0 = h_0 / h
    D0, D1 = m0, (1.0 / r0) * (m0 - m1)
    if algo == "dpmsolver++":
        base = (sg_t / sg_0 * torch.exp(-h)) * x + (al_t * (1 - torch.exp(-2.0 * h))) * D0
        if solver_type == "midpoint":
            mean = base + 0.5 * (al_t * (1 - torch.exp(-2.0 * h))) * D1
        elif solver_type == "heun":
            mean = base + (al_t * ((1.0 - torch.exp(-2.0 * h)) / (-2.0 * h) + 1.0)) * D1
        else:
            raise ValueError(solver_type)
        std, dt_sqrt = sg_t, torch.sqrt(1.0 - torch.exp(-2 * h))
        if sde:
            assert noise is not None
            out = mean + std * dt_sqrt * noise
        elif solver_type == "midpoint":
            out = (sg_t / sg_0) * x - (al_t * (torch.exp(-h) - 1.0)) * D0 - 0.5 * (al_t * (torch.exp(-h) - 1.0)) * D1
        else:
            out = (sg_t / sg_0) * x - (al_t * (torch.exp(-h) - 1.0)) * D0 + (al_t * ((torch.exp(-h) - 1.0) / h + 1.0)) * D1
    elif algo == "dpmsolver":
        lead = (al_t / al_0) * x - 2.0 * (sg_t * (torch.exp(h) - 1.0)) * D0
        if solver_type == "midpoint":
            mean = lead - (sg_t * (torch.exp(h) - 1.0)) * D1
        elif solver_type == "heun":
            mean = lead - 2.0 * (sg_t * ((torch.exp(h) - 1.0) / h - 1.0)) * D1
        else:
            raise ValueError(solver_type)
        std, dt_sqrt = sg_t, torch.sqrt(torch.exp(2 * h) - 1.0)
        if sde:
            assert noise is not None
            out = mean + std * dt_sqrt * noise
        elif solver_type == "midpoint":
            out = (al_t / al_0) * x - (sg_t * (torch.exp(h) - 1.0)) * D0 - 0.5 * (sg_t * (torch.exp(h) - 1.0)) * D1
        else:
            out = (al_t / al_0) * x - (sg_t * (torch.exp(h) - 1.0)) * D0 - (sg_t * ((torch.exp(h) - 1.0) / h - 1.0)) * D1
    else:
        raise ValueError(algo)
    return out, mean, std, dt_sqrt


def dpm_third_order(algo, hist, sigmas, i, x, noise, sde):
    """SU:563-639.  For algo == "dpmsolver" the reference returns names it never
    assigned (prev_mean/std_dev_t/dt_sqrt → UnboundLocalError, SU:629-639); we raise the same."""
    sg_t, sg_0, sg_1, sg_2 = sigmas[i + 1], sigmas[i], sigmas[i - 1], sigmas[i - 2]
    al_t, al_0 = 1 - sg_t, 1 - sg_0
    l_t, l_0, l_1, l_2 = _lam(sg_t), _lam(sg_0), _lam(sg_1), _lam(sg_2)
    m0, m1, m2 = hist[-1], hist[-2], hist[-3]
    h, h_0, h_1 = l_t - l_0, l_0 - l_1, l_1 - l_2
    r0, r1 = h_0 / h, h_1 / h
    D0 = m0
    D1_0, D1_1 = (1.0 / r0) * (m0 - m1), (1.0 / r1) * (m1 - m2)
    D1 = D1_0 + (r0 / (r0 + r1)) * (D1_0 - D1_1)
    D2 = (1.0 / (r0 + r1)) * (D1_0 - D1_1)
    if algo == "dpmsolver++":
        mean = ((sg_t / sg_0 * torch.exp(-h)) * x
                + (al_t * (1.0 - torch.exp(-2.0 * h))) * D0
                + (al_t * ((1.0 - torch.exp(-2.0 * h)) / (-2.0 * h) + 1.0)) * D1
                + (al_t * ((1.0 - torch.exp(-2.0 * h) - 2.0 * h) / (2.0 * h) ** 2 - 0.5)) * D2)
        std, dt_sqrt = sg_t, torch.sqrt(1.0 - torch.exp(-2 * h))
        if sde:
            assert noise is not None
            out = mean + std * dt_sqrt * noise
        else:
            out = ((sg_t / sg_0) * x
                   - (al_t * (torch.exp(-h) - 1.0)) * D0
                   + (al_t * ((torch.exp(-h) - 1.0) / h + 1.0)) * D1
                   - (al_t * ((torch.exp(-h) - 1.0 + h) / h ** 2 - 0.5)) * D2)
        return out, mean, std, dt_sqrt
    elif algo == "dpmsolver":
        assert not sde, "SDE solver is not supported for DPMSolver"
        raise UnboundLocalError("reference SU:639 returns prev_mean before assignment for dpmsolver order 3")
    raise ValueError(algo)


def dpm_step(v, x, step_index, n_timesteps, sigmas, *, algo="dpmsolver++", solver_order=2,
             solver_type="midpoint", history: Optional[History] = None, noise=None, sde_solver=False):
    """SU:273-385.  ``n_timesteps`` = len(timesteps) of the reference call (= len(sigmas)-1).
    Returns (x_next, x0, logp)."""
    final = step_index == n_timesteps - 1                        # SU:308
    second_last = (step_index == n_timesteps - 2) and n_timesteps < 15
    x0 = x0_from_velocity(v, x, sigmas, step_index)              # SU:311
    if history is not None:
        history.update(x0)
    x = x.to(torch.float32)
    nz = noise.to(device=x0.device, dtype=torch.float32) if sde_solver else None   # SU:318-325
    if history:
        if solver_order == 1 or history.lower_order_nums < 1 or final:
            out, mean, std, dts = dpm_first_order(algo, x0, sigmas, step_index, x, nz, sde_solver)
        elif solver_order == 2 or history.lower_order_nums < 2 or second_last:
            out, mean, std, dts = dpm_second_order(algo, solver_type, history.model_outputs, sigmas, step_index, x, nz, sde_solver)
        else:
            out, mean, std, dts = dpm_third_order(algo, history.model_outputs, sigmas, step_index, x, nz, sde_solver)
    else:
        out, mean, std, dts = dpm_first_order(algo, x0, sigmas, step_index, x, nz, sde_solver)
    if history is not None:
        history.update_lower_order()
    out = out.to(x0.dtype)                                       # SU:373
    logp = _gauss_logp(out, mean, std * dts)                     # SU:376-383
    return out, x0, logp


def flash_schedule(sigma_schedule, determistic: Sequence[bool], shift, compress_ratio):
    """Rebuild the schedule after the SDE window for dpm_apply_strategy == "post", SU:33-54.
    Returns (new_schedule, last_sde_index)."""
    n = sigma_schedule.size(0)
    last = None
    for k in range(len(determistic) - 1, -1, -1):
        if not determistic[k]:
            last = k
            break
    n_post = int(max((n - 1 - last) * compress_ratio, 1))        # SU:44
    t_post = torch.linspace(1, 0, n)[last + 1].item()            # SU:47
    tail = sd3_time_shift(shift, torch.linspace(t_post, 0, n_post).to(sigma_schedule.device))
    return torch.cat([sigma_schedule[: last + 1], tail], dim=0), last


def rollout(model, z, sigma_schedule, determistic, noises, *, eta, shift=3.0, flow_grpo_sampling=True,
            dpm_algorithm_type="null", dpm_apply_strategy="post", dpm_post_compress_ratio=0.4,
            dpm_solver_order=2, dpm_solver_type="midpoint", drop_last_sample=False):
    """run_sample_step, SU:12-155, with ``model(z, sigma, i) -> v`` standing in for the
    FLUX forward (SU:62-82) and ``noises[i]`` the explicit noise of step i.
    Returns (z, latents, all_latents, all_log_probs)."""
    all_latents, all_logp = [z], []
    hist = None
    last_sde = None
    if "dpmsolver" in dpm_algorithm_type:
        hist = History(dpm_solver_order)
        if dpm_apply_strategy == "post":
            sigma_schedule, last_sde = flash_schedule(sigma_schedule, determistic, shift, dpm_post_compress_ratio)
    x0 = None
    for i in range(sigma_schedule.size(0) - 1):
        v = model(z, sigma_schedule[i], i)
        zf = z.to(torch.float32)
        if dpm_algorithm_type == "null":
            if flow_grpo_sampling:
                z, x0, lp, _, _ = flow_step(v, zf, eta, sigma_schedule, i, None, noises[i], determistic[i])
            else:
                z, x0, lp = dance_step(v, zf, eta, sigma_schedule, i, None, noises[i], True, not determistic[i])
        elif dpm_apply_strategy == "all":
            z, x0, lp = dpm_step(v, zf, i, sigma_schedule.size(0) - 1, sigma_schedule, algo=dpm_algorithm_type,
                                 solver_order=dpm_solver_order, solver_type=dpm_solver_type, history=hist,
                                 noise=noises[i], sde_solver=not determistic[i])
        elif i <= last_sde:
            if flow_grpo_sampling:
                hist.update(x0_from_velocity(v, zf, sigma_schedule, i))      # SU:116-117
                z, x0, lp, _, _ = flow_step(v, zf, eta, sigma_schedule, i, None, noises[i], determistic[i])
                hist.update_lower_order()                                   # SU:127
            else:
                z, x0, lp = dance_step(v, zf, eta, sigma_schedule, i, None, noises[i], True, not determistic[i])
        else:
            z, x0, lp = dpm_step(v, zf, i, sigma_schedule.size(0) - 1, sigma_schedule, algo=dpm_algorithm_type,
                                 solver_order=dpm_solver_order, solver_type=dpm_solver_type, history=hist,
                                 noise=None, sde_solver=False)
        all_latents.append(z)
        all_logp.append(lp)
    latents = x0 if drop_last_sample else z.to(x0.dtype)          # SU:149-152
    return z, latents, torch.stack(all_latents, dim=1), torch.stack(all_logp, dim=1)
