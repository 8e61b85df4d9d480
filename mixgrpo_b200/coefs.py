"""Per-step scalar tables for the fused step kernels, computed on the HOST.

The reference evaluates ~20 zero-dim torch ops per sampler call to get ``std_dev_t``, the mean's two
coefficients, the noise scale etc. (SU:170-177, SU:186, SU:421-431, SU:472-504, SU:586-617;
SU = /root/reference/fastvideo/utils/sampling_utils.py) and syncs the host each call
(``sigmas[1].item()`` SU:172, ``math.sqrt(tensor)`` SU:229).  The schedule is known before the loop,
so here the same fp32 expressions — same operator order, evaluated with zero-dim fp32 torch CPU
tensors so each op rounds exactly as the reference's — run once per (schedule, step) on the host, are
cached, and reach the kernel by value in a ``mixgrpo_step_coefs`` block.  One D2H copy per *schedule*
replaces one sync per *step*.

Rounding modes.  With bf16 ``model_output`` torch's type promotion rounds some
products to bf16 and casts the zero-dim scalar factor to bf16 first:
  ``fp32``      no emulation: scalars stay fp32, kernel intermediates stay fp32 (fastest, most accurate)
  ``ref_cpu``   what the reference does on CPU tensors: a zero-dim factor is cast to bf16 when it is the
                LEFT operand (``sigma*v``, ``scale*noise``, ``dt*v``) and kept fp32 when on the right
                (``v*c_v``, ``(...)*dt``)
  ``ref_cuda``  what the reference does on CUDA tensors: zero-dim CUDA tensors are always cast to the
                common dtype (bf16) — all five factors are rounded
With fp32 ``model_output`` the three modes coincide.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Tuple

import torch

from ._cabi import StepCoefs

MODES = ("fp32", "ref_cpu", "ref_cuda")

_SCHED_CACHE_MAX = 64
_sched_cache: "OrderedDict[Tuple, Tuple]" = OrderedDict()   # key -> (strong ref, host fp32 tensor, python list)
_coef_cache: Dict[Tuple, Tuple] = {}


def _host_entry(sigmas: torch.Tensor):
    """(fp32 CPU tensor, python list) of a sigma schedule.

    Cached per live tensor: the entry keeps a strong reference to ``sigmas`` so its storage address
    cannot be recycled for a different schedule while the key (address, version, ...) is in the
    cache; ``_version`` catches in-place edits.  A CUDA schedule therefore costs one D2H copy for
    the whole rollout instead of one ``.item()`` per step (SU:172)."""
    key = (sigmas.device.type, sigmas.device.index, sigmas.data_ptr(), sigmas._version, sigmas.numel(),
           sigmas.dtype, sigmas.stride())
    hit = _sched_cache.get(key)
    if hit is not None:
        _sched_cache.move_to_end(key)
        return hit[1], hit[2]
    host = sigmas.detach().to(device="cpu", dtype=torch.float32).contiguous()
    if host.data_ptr() == sigmas.data_ptr():
        host = host.clone()
    entry = (sigmas, host, host.tolist())
    _sched_cache[key] = entry
    while len(_sched_cache) > _SCHED_CACHE_MAX:
        _sched_cache.popitem(last=False)
    return entry[1], entry[2]


def host_schedule(sigmas: torch.Tensor) -> torch.Tensor:
    return _host_entry(sigmas)[0]


def _bf(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def _f(t) -> float:
    return float(t)


def _pack(two_var, log_scale, log_norm, cs) -> StepCoefs:
    k = StepCoefs()
    k.two_var, k.log_scale, k.log_norm = _f(two_var), _f(log_scale), _f(log_norm)
    for i, v in enumerate(cs):
        k.c[i] = _f(v)
    return k


def _log_norm() -> torch.Tensor:
    return torch.log(torch.sqrt(2 * torch.as_tensor(math.pi)))            # SU:204


def flow(sigmas: torch.Tensor, index: int, eta: float, mode: str, bf16_v: bool):
    """Scalars of flow_grpo_step (SU:170-177, 186, 195, 199, 201-204).  Returns (StepCoefs, scale_f32)."""
    sig, vals = _host_entry(sigmas)
    key = ("flow", vals[index], vals[index + 1], vals[1], float(eta), mode, bool(bf16_v))
    hit = _coef_cache.get(key)
    if hit is not None:
        return hit
    s, s_next = sig[index], sig[index + 1]
    s_guard = sig[1].item()                                               # SU:172
    dt = s_next - s
    std = torch.sqrt(s / (1 - torch.where(s == 1, s_guard, s))) * eta     # SU:177
    c_x = 1 + std ** 2 / (2 * s) * dt                                     # SU:186
    c_v = 1 + std ** 2 * (1 - s) / (2 * s)
    scale = std * torch.sqrt(-1 * dt)
    two_var = 2 * (scale ** 2)                                            # SU:202
    emulate = bf16_v and mode != "fp32"
    left = _bf if emulate else (lambda t: t)                              # zero-dim factor on the left
    right = _bf if (emulate and mode == "ref_cuda") else (lambda t: t)    # ... on the right
    k = _pack(two_var, torch.log(scale), _log_norm(),
              [left(s), c_x, right(c_v), right(dt), left(scale), left(dt)])
    out = (k, float(scale))
    _remember(key, out)
    return out


def x0_only(sigmas: torch.Tensor, index: int, mode: str, bf16_v: bool) -> StepCoefs:
    """Only c[0] = sigma[index] (convert_model_output, SU:393-394); everything else is inert."""
    sig, _ = _host_entry(sigmas)
    s = sig[index]
    s0 = _bf(s) if (bf16_v and mode != "fp32") else s
    return _pack(1.0, 0.0, 0.0, [s0, 1.0, 0.0, 0.0, 0.0, 0.0])


def dance(sigmas: torch.Tensor, index: int, eta: float, mode: str, bf16_v: bool):
    """Scalars of dance_grpo_step (SU:222-234, 238, 244-246).  Returns (StepCoefs, std python float)."""
    sig, vals = _host_entry(sigmas)
    key = ("dance", vals[index], vals[index + 1], float(eta), mode, bool(bf16_v))
    hit = _coef_cache.get(key)
    if hit is not None:
        return hit
    s = sig[index]
    ds = sig[index + 1] - s
    delta = s - sig[index + 1]
    std = eta * math.sqrt(delta)                                          # python double, SU:229
    emulate = bf16_v and mode != "fp32"
    left = _bf if emulate else (lambda t: t)
    # python scalars multiplying / dividing an fp32 tensor are narrowed to fp32 by torch
    k = _pack(torch.tensor(2 * (std ** 2), dtype=torch.float32), 0.0, 0.0,
              [left(s), left(ds), 1 - s, s ** 2, torch.tensor(-0.5 * eta ** 2, dtype=torch.float32), ds,
               torch.tensor(std, dtype=torch.float32)])
    out = (k, std)
    _remember(key, out)
    return out


def _lam(sg):
    return torch.log(1 - sg) - torch.log(sg)                              # SU:424-425


def dpm(sigmas: torch.Tensor, index: int, order: int, algo: str, solver_type: str, mode: str, bf16_v: bool):
    """Scalars of dpm_step's order-``order`` update (SU:421-447, 472-561, 586-639) with the signs of
    the reference's subtractions folded into the coefficients:  a - (s)*D == a + (-s)*D exactly.
    Returns (StepCoefs, scale_f32)."""
    sig, vals = _host_entry(sigmas)
    lo = index - (order - 1)
    key = ("dpm", tuple(vals[lo:index + 2]), int(order), algo, solver_type, mode, bool(bf16_v))
    hit = _coef_cache.get(key)
    if hit is not None:
        return hit
    if algo not in ("dpmsolver++", "dpmsolver"):
        raise ValueError(algo)
    if order >= 2 and solver_type not in ("midpoint", "heun"):
        raise ValueError(solver_type)
    i = index
    sg_t, sg_0 = sig[i + 1], sig[i]
    al_t, al_0 = 1 - sg_t, 1 - sg_0
    l_t, l_0 = _lam(sg_t), _lam(sg_0)
    h = l_t - l_0
    zero = torch.zeros((), dtype=torch.float32)
    k1 = k2 = k3 = k4 = zero
    m = [zero] * 4      # mean coefficients  (x, D0, D1, D2)
    o = [zero] * 4      # ODE coefficients
    if order >= 2:
        l_1 = _lam(sig[i - 1])
        h_0 = l_0 - l_1
        r0 = h_0 / h
        k1 = 1.0 / r0                                                     # SU:490 / SU:608
    if order == 3:
        l_2 = _lam(sig[i - 2])
        h_1 = l_1 - l_2
        r1 = h_1 / h
        k2 = 1.0 / r1
        k3 = r0 / (r0 + r1)                                               # SU:609
        k4 = 1.0 / (r0 + r1)                                              # SU:610
    if algo == "dpmsolver++":
        m[0] = sg_t / sg_0 * torch.exp(-h)
        m[1] = al_t * (1 - torch.exp(-2.0 * h))
        o[0] = sg_t / sg_0
        o[1] = -(al_t * (torch.exp(-h) - 1.0))
        if order == 2 and solver_type == "midpoint":                      # SU:494-498, 516-520
            m[2] = 0.5 * (al_t * (1 - torch.exp(-2.0 * h)))
            o[2] = -(0.5 * (al_t * (torch.exp(-h) - 1.0)))
        elif order == 2:                                                  # heun, SU:500-504, 522-526
            m[2] = al_t * ((1.0 - torch.exp(-2.0 * h)) / (-2.0 * h) + 1.0)
            o[2] = al_t * ((torch.exp(-h) - 1.0) / h + 1.0)
        elif order == 3:                                                  # SU:612-628
            m[1] = al_t * (1.0 - torch.exp(-2.0 * h))
            m[2] = al_t * ((1.0 - torch.exp(-2.0 * h)) / (-2.0 * h) + 1.0)
            m[3] = al_t * ((1.0 - torch.exp(-2.0 * h) - 2.0 * h) / (2.0 * h) ** 2 - 0.5)
            o[2] = al_t * ((torch.exp(-h) - 1.0) / h + 1.0)
            o[3] = -(al_t * ((torch.exp(-h) - 1.0 + h) / h ** 2 - 0.5))
        dt_sqrt = torch.sqrt(1.0 - torch.exp(-2 * h))
    else:  # "dpmsolver"
        if order == 3:
            # reference returns unassigned names here (SU:629-639)
            raise UnboundLocalError("dpmsolver order 3: the reference never assigns prev_mean/std_dev_t/dt_sqrt (SU:639)")
        m[0] = al_t / al_0
        m[1] = -(2.0 * (sg_t * (torch.exp(h) - 1.0)))
        o[0] = al_t / al_0
        o[1] = -(sg_t * (torch.exp(h) - 1.0))
        if order == 2 and solver_type == "midpoint":                      # SU:529-533, 549-553
            m[2] = -(sg_t * (torch.exp(h) - 1.0))
            o[2] = -(0.5 * (sg_t * (torch.exp(h) - 1.0)))
        elif order == 2:                                                  # SU:535-539, 555-559
            m[2] = -(2.0 * (sg_t * ((torch.exp(h) - 1.0) / h - 1.0)))
            o[2] = -(sg_t * ((torch.exp(h) - 1.0) / h - 1.0))
        dt_sqrt = torch.sqrt(torch.exp(2 * h) - 1.0)
    scale = sg_t * dt_sqrt                                                # SU:434 std_dev_t * dt_sqrt
    emulate = bf16_v and mode != "fp32"
    sig_x0 = _bf(sg_0) if emulate else sg_0                               # SU:394 sigma_t * model_output
    # the same factor as autograd's mul backward sees it (`grad * sigma`: sigma is then the RIGHT operand, which torch
    # keeps in fp32 on CPU and casts to bf16 on CUDA) — c[14], read by the order-1 log-prob backward only
    sig_bwd = _bf(sg_0) if (emulate and mode == "ref_cuda") else sg_0
    k = _pack(2 * (scale ** 2), torch.log(scale), _log_norm(),
              [sig_x0, k1, k2, k3, k4, *m, *o, scale, sig_bwd])
    out = (k, float(scale))
    _remember(key, out)
    return out


def _remember(key, value):
    if len(_coef_cache) > 8192:
        _coef_cache.clear()
    _coef_cache[key] = value
