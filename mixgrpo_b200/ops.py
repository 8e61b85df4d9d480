"""Tensor-level wrappers over the C ABI (``include/mixgrpo_b200.h``): argument validation, output
allocation through torch's caching allocator, workspace management, stream plumbing.

Everything here launches hand-written sm_100a kernels from ``csrc/``.  CPU tensors are rejected — there
is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Tuple

import torch

from . import _cabi
from ._cabi import (BF16, F32, FLAG_DEFER_LOGP, FLAG_PDL_EARLY_LOADS, FLAG_PDL_EARLY_V, FLAG_ROUND_LIKE_TORCH, POLICY_MAX_ITEMS, SRC_DETERMINISTIC,
                    SRC_GIVEN, SRC_NOISE, SRC_PHILOX, LossArgs, PhiloxArgs, PolicyItem, StepCoefs, StepExt)

FLOW, DANCE, DPM = 0, 1, 2

_workspaces: Dict[Tuple[int, int], torch.Tensor] = {}

#: number of kernels launched through this module (bench.py reports it as ``gpu_launches``)
launch_count = 0

_binding_mod = None
_binding_tried = False


def binding():
    """The compiled binding ``_lib/_torchbind.so`` (csrc_bind/torch_bind.cpp: pybind11 + torch::autograd over the SAME C ABI)
    or ``None`` — then every call goes through ctypes (``_cabi``), which is only a slower way to reach the same entry points.
    ``MIXGRPO_BINDING=ctypes`` forces that; a missing / stale binding is rebuilt when a C++ compiler is here (one minute)
    unless ``MIXGRPO_NO_REBUILD=1``."""
    global _binding_mod, _binding_tried
    if _binding_tried:
        return _binding_mod
    _binding_tried = True
    if os.environ.get("MIXGRPO_BINDING", "compiled") == "ctypes":
        return None
    from . import _build
    _cabi.lib()                                           # libmixgrpo_b200.so first: the binding links against it
    try:
        if not _build.binding_is_fresh() and os.environ.get("MIXGRPO_NO_REBUILD") != "1":
            _build.build_binding()
        if _build.BIND_LIB.exists():
            import importlib.util
            spec = importlib.util.spec_from_file_location("_torchbind", str(_build.BIND_LIB))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            if mod.abi_version() == _cabi.ABI_VERSION:
                _binding_mod = mod
    except Exception as e:  # noqa: BLE001  (no compiler / torch headers: the ctypes loader still reaches every kernel)
        import warnings
        warnings.warn(f"mixgrpo_b200: compiled binding unavailable ({type(e).__name__}: {e}); using the ctypes loader")
        _binding_mod = None
    return _binding_mod


def total_launches() -> int:
    """``launch_count`` plus the backward launches the compiled binding's C++ autograd Function issued (those never pass
    through this module)."""
    return launch_count + (_binding_mod.autograd_backward_launches() if _binding_mod is not None else 0)


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor) or t.device.type != "cuda":
        raise RuntimeError(f"mixgrpo_b200: `{name}` must be a CUDA tensor — this package has no CPU fallback "
                           f"(got {getattr(t, 'device', type(t))})")


def _dtype_code(t: torch.Tensor, name: str) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"mixgrpo_b200: `{name}` must be float32 or bfloat16, got {t.dtype}")


def _stream_ptr(device: torch.device) -> int:
    """Raw cudaStream_t of torch's current stream on ``device`` (torch._C._cuda_getCurrentRawStream: no Stream object)."""
    idx = device.index
    return torch._C._cuda_getCurrentRawStream(idx if idx is not None else torch.cuda.current_device())


class _on_device:
    """``with _on_device(dev)`` that costs nothing when ``dev`` is already current (the usual case: one process
    per GPU) — the C ABI launches on the CURRENT device's context."""
    __slots__ = ("ctx",)

    def __init__(self, device: torch.device):
        idx = device.index
        self.ctx = None if (idx is None or idx == torch.cuda.current_device()) else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _rows(t: torch.Tensor, name: str) -> Tuple[torch.Tensor, int]:
    """Return (tensor, batch stride in elements) for a (B, ...) tensor whose per-sample block is
    contiguous; anything else is made contiguous (one copy)."""
    if t.dim() < 1:
        raise ValueError(f"mixgrpo_b200: `{name}` needs a batch dimension")
    if t.dim() == 1:
        t = t.unsqueeze(1)
    if t.shape[0] == 1:
        return (t if t[0].is_contiguous() else t.contiguous()), t[0].numel()
    if t[0].is_contiguous():
        return t, t.stride(0)
    t = t.contiguous()
    return t, t.stride(0)


def _workspace(device: torch.device, B: int, n: int, stream: Optional[int] = None) -> torch.Tensor:
    need = _cabi.lib().mixgrpo_step_workspace_bytes(B, n)
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream if stream is not None else _stream_ptr(device))
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 1 << 16), dtype=torch.uint8, device=device)   # counters must start at zero
        if not torch.cuda.is_current_stream_capturing():
            _workspaces[key] = ws
        # else: allocated inside a CUDA-graph capture (no eager warm-up on this stream) — the zero fill is part of the graph
        # and the memory belongs to the graph's pool, so it must not be handed to later eager calls
    return ws


_deferred_bufs: Dict[Tuple[int, int, int], torch.Tensor] = {}


class DeferredLogProbs:
    """Records for ``n_launches`` step launches that only ACCUMULATE their log-prob sums (MIXGRPO_FLAG_DEFER_LOGP) plus the one
    ``finalize`` launch that turns them into log-probs.  A rollout's log-probs are not read until the rollout is over
    (SU:153-155), so none of its step launches needs the returning atomic that makes a CTA wait an L2 round trip before it
    retires; the sums are the same packed integers either way, so the finalized values are bit-identical to the immediate
    path.  Zeroed once; ``finalize`` leaves the records zeroed (one cached buffer per device, stream and size)."""

    def __init__(self, device: torch.device, n_launches: int, B: int, n: int):
        self.device, self.n_launches, self.B = device, int(n_launches), int(B)
        self.stride = int(_cabi.lib().mixgrpo_deferred_workspace_bytes(B, n))     # 8 spread sub-records per sample
        st = _stream_ptr(device)
        key = (device.index if device.index is not None else torch.cuda.current_device(), st, self.stride * self.n_launches)
        buf = _deferred_bufs.get(key)
        if buf is None:
            buf = torch.zeros(max(self.stride * self.n_launches, 8), dtype=torch.uint8, device=device)
            if not torch.cuda.is_current_stream_capturing():
                _deferred_bufs[key] = buf            # never evicted: a captured graph may hold its address (77 KB per key for a 25-step rollout of 12 samples)
        self.buf = buf
        self.log_scale = (C.c_float * self.n_launches)()
        self.log_norm = (C.c_float * self.n_launches)()
        self.active = (C.c_int * self.n_launches)()

    def slot(self, i: int, coefs: StepCoefs) -> Tuple[int, int]:
        """(device pointer, bytes) of launch ``i``'s records; remembers the step's two log-prob constants for ``finalize``."""
        self.log_scale[i], self.log_norm[i], self.active[i] = coefs.log_scale, coefs.log_norm, 1
        return self.buf.data_ptr() + i * self.stride, self.stride

    def finalize(self, out: torch.Tensor) -> torch.Tensor:
        """out[i, b] (fp32 ``[n_launches, B]``, row stride >= B) = launch i's log-prob of sample b; NaN rows for launches that
        never took a slot.  One launch."""
        global launch_count
        if out.dtype != torch.float32 or out.dim() != 2 or out.shape[0] != self.n_launches or out.shape[1] != self.B or out.stride(1) != 1:
            raise ValueError("mixgrpo_b200: finalize needs an fp32 [n_launches, B] tensor with contiguous rows")
        if self.B == 0 or self.n_launches == 0:
            return out
        with _on_device(self.device):
            rc = _cabi.lib().mixgrpo_logp_finalize(self.buf.data_ptr(), self.stride, self.n_launches, self.B, self.log_scale, self.log_norm, self.active,
                                                   out.data_ptr(), out.stride(0), _stream_ptr(self.device))
        _cabi.check(rc, "logp_finalize")
        launch_count += 1
        return out


class PhiloxState:
    """CUDA-graph-safe state for in-kernel noise: a device vector ``{seed, base offset}`` the step kernels read at RUN time
    (``mixgrpo_philox_args.device_state``).  A launch captured into a graph keeps only its position relative to the base
    (``take``), and ``advance`` — one tiny launch, captured with the rollout — moves the base, so every replay draws fresh
    noise.  Eagerly it behaves like a private generator.  (Host-read seeds/offsets, ``philox_from_generator``, would be
    frozen into the graph: every replay would repeat the same exploration noise.)"""

    def __init__(self, device, seed: int = 0, offset: int = 0):
        self.device = torch.device(device)
        vals = [int(seed) & 0x7FFFFFFFFFFFFFFF, int(offset) & 0x7FFFFFFFFFFFFFFF]
        self.state = torch.tensor(vals, dtype=torch.int64).to(self.device)
        self._pending = 0                       # numbers handed out since the last advance (host-known, constant per graph)

    def take(self, numel: int) -> Tuple["PhiloxState", int]:
        """Reserve ``numel`` normals: returns ``(self, relative offset)`` for ``fused_step(philox=...)``."""
        rel = self._pending
        self._pending += 4 * ((int(numel) + 3) // 4)
        return self, rel

    def advance(self) -> None:
        """``base += everything taken since the last advance`` on the device (stream-ordered, capturable)."""
        global launch_count
        if self._pending == 0:
            return
        with _on_device(self.device):
            rc = _cabi.lib().mixgrpo_philox_advance(self.state.data_ptr(), self._pending, _stream_ptr(self.device))
        _cabi.check(rc, "philox_advance")
        launch_count += 1
        self._pending = 0


def philox_from_generator(device: torch.device, numel: int, generator: Optional[torch.Generator] = None) -> Tuple[int, int]:
    """(seed, offset) for MIXGRPO_SRC_PHILOX taken from a torch CUDA generator (the device default when None), whose
    Philox offset is advanced past the ``numel`` normals consumed — later torch.randn calls draw fresh numbers, and
    ``torch.manual_seed`` / ``generator.manual_seed`` reproduce the in-kernel noise too.

    Not capturable: the values are read on the host and would be frozen into a CUDA graph (identical noise on every
    replay) — raises under stream capture; use a :class:`PhiloxState` there."""
    if torch.cuda.is_current_stream_capturing():
        raise RuntimeError("mixgrpo_b200: in-kernel noise seeded from a torch generator cannot be captured into a CUDA graph "
                           "(seed/offset are host values: every replay would draw the same noise); pass a "
                           "mixgrpo_b200.ops.PhiloxState (rollout(..., philox_state=...))")
    if generator is None or generator.device.type != "cuda":
        idx = device.index if device.index is not None else torch.cuda.current_device()
        seed_src = generator
        generator = torch.cuda.default_generators[idx]
        if seed_src is not None:                       # a CPU generator only lends its seed
            return int(seed_src.initial_seed()), int(torch.randint(0, 2 ** 31, (1,), generator=seed_src).item()) * 4
    seed, off = int(generator.initial_seed()), int(generator.get_offset())
    generator.set_offset(off + 4 * ((numel + 3) // 4))
    return seed, off


def fused_step(family: int, v: torch.Tensor, x: torch.Tensor, coefs: StepCoefs, *, src: int,
               noise: Optional[torch.Tensor] = None, x_next: Optional[torch.Tensor] = None,
               m1: Optional[torch.Tensor] = None, m2: Optional[torch.Tensor] = None, order: int = 1,
               sde_solver: bool = True, out_x_next: Optional[torch.Tensor] = None, want_x0: bool = True,
               want_mean: bool = False, want_logp: bool = True, round_like_torch: bool = False,
               out_logp: Optional[torch.Tensor] = None, philox=None,
               out_x0: Optional[torch.Tensor] = None, early: int = 0, decode: Optional[dict] = None,
               defer: Optional[Tuple[int, int]] = None, seed_out: Optional[torch.Tensor] = None):
    """One fused sampler step + log-prob launch.  Returns (x_next, x0, logp, mean); entries not
    requested are None; with ``src == SRC_GIVEN`` x_next is the tensor passed in.

    ``philox``: ``(seed, offset)`` host values or ``(PhiloxState, relative offset)`` (graph-safe).
    ``early``: 1 = model output / noise were not written by the immediately preceding launch on this stream
    (MIXGRPO_FLAG_PDL_EARLY_V), 2 = no streamed input was (MIXGRPO_FLAG_PDL_EARLY_LOADS).
    ``decode``: ``{"out": fp32 (B,C,H,W), "divisor": 0.3611, "shift": 0.1159, "from_x0": False, "reciprocal": False}`` —
    the VAE's input (unpack + de-normalise, TR:102-115, TR:286-287) written by this launch as a second output.
    ``defer``: ``DeferredLogProbs.slot(i, coefs)`` — the launch only accumulates its log-prob sums there (no log-prob is
    returned; ``want_logp`` is ignored).
    ``seed_out``: the rollout's first step — ``x`` is the bf16 initial latent itself and its fp32 widening (``all_latents[:, 0]``,
    SU:26 / SU:153) is written to this fp32 view as well: no separate cast launch (flow family; see ``can_seed``)."""
    global launch_count
    tb = binding()
    if tb is not None:                                    # compiled binding: same checks, same C-ABI call, no interpreter time
        has_ph, seed, off, state = False, 0, 0, 0
        if src == SRC_PHILOX and philox is not None:
            has_ph = True
            if isinstance(philox[0], PhiloxState):
                off, state = int(philox[1]) & 0xFFFFFFFFFFFFFFFF, philox[0].state.data_ptr()
            else:
                seed, off = int(philox[0]) & 0xFFFFFFFFFFFFFFFF, int(philox[1]) & 0xFFFFFFFFFFFFFFFF
        d_ptr, d_bytes = defer if defer is not None else (0, 0)
        if defer is not None:
            want_logp, out_logp = False, None
        if decode is None:
            res = tb.fused_step(family, v, x, C.addressof(coefs), src, noise, x_next, m1, m2, order, sde_solver, out_x_next, want_x0, want_mean,
                                want_logp, round_like_torch, out_logp, out_x0, early, has_ph, seed, off, state, None, 1.0, 0.0, False, False,
                                d_ptr, d_bytes, seed_out)
        else:
            res = tb.fused_step(family, v, x, C.addressof(coefs), src, noise, x_next, m1, m2, order, sde_solver, out_x_next, want_x0, want_mean,
                                want_logp, round_like_torch, out_logp, out_x0, early, has_ph, seed, off, state, decode["out"],
                                float(decode.get("divisor", 1.0)), float(decode.get("shift", 0.0)), bool(decode.get("from_x0", False)),
                                bool(decode.get("reciprocal", False)), d_ptr, d_bytes, None)
        launch_count += 1 if v.shape[0] else 0
        return res
    lib = _cabi.lib()
    if defer is not None:
        want_logp, out_logp = False, None
    _require_cuda(v, "model_output")
    _require_cuda(x, "latents")
    if seed_out is not None:
        if decode is not None or x.dtype != torch.bfloat16 or not x.is_contiguous() or seed_out.dtype != torch.float32 or seed_out.shape != x.shape \
                or not seed_out[0].is_contiguous():
            raise ValueError("mixgrpo_b200: seed_out needs a contiguous bf16 `latents`, an fp32 view of the same shape, and no decode output")
    elif x.dtype != torch.float32:
        x = x.to(torch.float32)
    vd = _dtype_code(v, "model_output")
    if v.shape != x.shape:
        raise ValueError(f"mixgrpo_b200: model_output {tuple(v.shape)} and latents {tuple(x.shape)} differ")
    v = v if v.is_contiguous() else v.contiguous()
    B = v.shape[0]
    dev = v.device
    if B == 0:
        # empty batch: every reference op is a no-op on empty tensors and the per-sample log-prob is an empty [0] vector
        f32 = lambda: torch.empty(v.shape, dtype=torch.float32, device=dev)   # noqa: E731
        return ((x_next if src == SRC_GIVEN else (out_x_next if out_x_next is not None else f32())), f32() if want_x0 else None,
                (out_logp if out_logp is not None else torch.empty((0,), dtype=torch.float32, device=dev)) if want_logp else None,
                f32() if want_mean else None)
    n = v[0].numel()
    x, x_bs = (x, n) if seed_out is not None else _rows(x, "latents")
    noise_p = in_p = m1_p = m2_p = None
    in_bs = n
    keep = [v, x]
    if src == SRC_NOISE:
        if noise is None:
            raise ValueError("mixgrpo_b200: rollout step needs explicit `noise`")
        _require_cuda(noise, "noise")
        want = v.dtype if family == FLOW else torch.float32      # SU:193 vs SU:238 / SU:320
        if noise.dtype != want or not noise.is_contiguous():
            noise = noise.to(want).contiguous()
        if noise.shape != v.shape:
            raise ValueError("mixgrpo_b200: noise shape mismatch")
        noise_p = noise.data_ptr()
        keep.append(noise)
    elif src == SRC_PHILOX:
        if philox is None:
            raise ValueError("mixgrpo_b200: in-kernel noise needs philox=(seed, offset)")
        if isinstance(philox[0], PhiloxState):
            pa = PhiloxArgs(0, int(philox[1]) & 0xFFFFFFFFFFFFFFFF, philox[0].state.data_ptr())
        else:
            pa = PhiloxArgs(int(philox[0]) & 0xFFFFFFFFFFFFFFFF, int(philox[1]) & 0xFFFFFFFFFFFFFFFF, None)
        keep.append(pa)
        noise_p = C.cast(C.pointer(pa), C.c_void_p)          # HOST pointer, read during the call
    elif src == SRC_GIVEN:
        _require_cuda(x_next, "prev_sample")
        if x_next.dtype != torch.float32:
            x_next = x_next.to(torch.float32)
        if x_next.shape != v.shape:
            raise ValueError("mixgrpo_b200: prev_sample shape mismatch")
        x_next, in_bs = _rows(x_next, "prev_sample")
        in_p = x_next.data_ptr()
        keep.append(x_next)
    if family == DPM and order >= 2:
        m1 = m1.contiguous()
        m1_p = m1.data_ptr()
        keep.append(m1)
        if order == 3:
            m2 = m2.contiguous()
            m2_p = m2.data_ptr()
            keep.append(m2)

    out = None
    out_p, out_bs = None, n
    if src != SRC_GIVEN:
        if out_x_next is None:
            out = torch.empty(v.shape, dtype=torch.float32, device=dev)
            out_p = out.data_ptr()
        else:
            if out_x_next.dtype != torch.float32 or out_x_next.shape != v.shape or not out_x_next[0].is_contiguous():
                raise ValueError("mixgrpo_b200: out_x_next must be fp32, same shape, contiguous per sample")
            out = out_x_next
            out_p, out_bs = out.data_ptr(), (out.stride(0) if B > 1 else n)
    x0 = None
    if want_x0:
        if out_x0 is not None:
            if out_x0.dtype != torch.float32 or out_x0.shape != v.shape or not out_x0.is_contiguous():
                raise ValueError("mixgrpo_b200: out_x0 must be a contiguous fp32 tensor of the model output's shape")
            x0 = out_x0
        else:
            x0 = torch.empty(v.shape, dtype=torch.float32, device=dev)
    mean = torch.empty(v.shape, dtype=torch.float32, device=dev) if want_mean else None
    logp = None
    if want_logp:
        if out_logp is not None:
            if out_logp.dtype != torch.float32 or out_logp.numel() != B or not out_logp.is_contiguous():
                raise ValueError("mixgrpo_b200: out_logp must be a contiguous fp32 [B] tensor")
            logp = out_logp
        else:
            logp = torch.empty((B,), dtype=torch.float32, device=dev)
    st = _stream_ptr(dev)
    ws = _workspace(dev, B, n, st) if want_logp else None
    flags = (FLAG_ROUND_LIKE_TORCH if (round_like_torch and vd == BF16) else 0) | (FLAG_PDL_EARLY_LOADS if early == 2 else (FLAG_PDL_EARLY_V if early == 1 else 0))
    if defer is not None:
        flags |= FLAG_DEFER_LOGP
    ext = None
    if decode is not None:
        d_out = decode["out"]
        _require_cuda(d_out, "decode['out']")
        if d_out.dtype != torch.float32 or d_out.dim() != 4 or d_out.shape[0] != B or d_out[0].numel() != n or not d_out.is_contiguous():
            raise ValueError("mixgrpo_b200: decode['out'] must be a contiguous fp32 (B, C, H, W) tensor with C*H*W == elements per sample")
        xe = StepExt(d_out.data_ptr(), d_out.shape[1], d_out.shape[2], d_out.shape[3], float(decode.get("divisor", 1.0)),
                     float(decode.get("shift", 0.0)), 1 if decode.get("from_x0", False) else 0, 1 if decode.get("reciprocal", False) else 0,
                     0, None, 0)
        keep.append(xe)
        ext = C.byref(xe)
    elif seed_out is not None:
        xe = StepExt(None, 0, 0, 0, 1.0, 0.0, 0, 0, 1, seed_out.data_ptr(), seed_out.stride(0) if B > 1 else n)
        keep.append(xe)
        ext = C.byref(xe)
    common_out = (out_p, out_bs, x0.data_ptr() if want_x0 else None, mean.data_ptr() if want_mean else None,
                  logp.data_ptr() if want_logp else None, (defer[0] if defer is not None else (ws.data_ptr() if want_logp else None)),
                  (defer[1] if defer is not None else (ws.numel() if want_logp else 0)), B, n, C.byref(coefs))
    with _on_device(dev):
        if family == FLOW:
            rc = lib.mixgrpo_flow_step(v.data_ptr(), vd, x.data_ptr(), x_bs, noise_p, in_p, in_bs, *common_out, src, flags, st, ext)
        elif family == DANCE:
            rc = lib.mixgrpo_dance_step(v.data_ptr(), vd, x.data_ptr(), x_bs, noise_p, in_p, in_bs, *common_out, src,
                                        1 if sde_solver else 0, flags, st, ext)
        elif family == DPM:
            rc = lib.mixgrpo_dpm_step(v.data_ptr(), vd, x.data_ptr(), x_bs, noise_p, m1_p, m2_p, order, *common_out, src, flags, st, ext)
        else:
            raise ValueError(family)
    _cabi.check(rc, ("flow_step", "dance_step", "dpm_step")[family])
    launch_count += 1
    del keep
    return (x_next if src == SRC_GIVEN else out), x0, logp, mean


def can_seed(z: torch.Tensor, dst: torch.Tensor) -> bool:
    """Whether ``fused_step(..., x=z, seed_out=dst)`` covers these tensors (bf16 contiguous ``z``, fp32 ``dst`` rows, the
    256-bit vector path); otherwise seed the trajectory with ``cast_rows`` first."""
    if z.dtype != torch.bfloat16 or not z.is_contiguous() or z.dim() < 2 or z.shape[0] == 0 or dst.dtype != torch.float32 or dst.shape != z.shape:
        return False
    n = z[0].numel()
    bs = dst.stride(0) if z.shape[0] > 1 else n
    return n % 8 == 0 and bs % 8 == 0 and dst[0].is_contiguous() and z.data_ptr() % 32 == 0 and dst.data_ptr() % 32 == 0


def logprob_backward(family: int, v: torch.Tensor, x: torch.Tensor, x_next: torch.Tensor, grad_logp: torch.Tensor,
                     coefs: StepCoefs, round_like_torch: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """grad of sum_b grad_logp[b]*logp[b] w.r.t. model_output; dtype = model_output.dtype."""
    global launch_count
    tb = binding()
    if tb is not None:
        launch_count += 1 if v.shape[0] else 0
        return tb.logprob_backward(family, v, x, x_next, grad_logp, C.addressof(coefs), round_like_torch, out)
    lib = _cabi.lib()
    for t, nm in ((v, "model_output"), (x, "latents"), (x_next, "prev_sample"), (grad_logp, "grad_log_prob")):
        _require_cuda(t, nm)
    vd = _dtype_code(v, "model_output")
    v = v if v.is_contiguous() else v.contiguous()
    if v.shape[0] == 0:
        return out if out is not None else torch.empty_like(v)
    B, n = v.shape[0], v[0].numel()
    x, x_bs = _rows(x.to(torch.float32), "latents")
    x_next, in_bs = _rows(x_next.to(torch.float32), "prev_sample")
    g = grad_logp.to(torch.float32).contiguous()
    if g.numel() != B:
        raise ValueError("mixgrpo_b200: grad_log_prob must have one entry per sample")
    grad_v = out if out is not None else torch.empty_like(v)
    if grad_v.dtype != v.dtype or grad_v.shape != v.shape or not grad_v.is_contiguous():
        raise ValueError("mixgrpo_b200: `out` must match model_output's dtype/shape and be contiguous")
    flags = FLAG_ROUND_LIKE_TORCH if (round_like_torch and vd == BF16) else 0
    with _on_device(v.device):
        rc = lib.mixgrpo_logprob_bwd(family, v.data_ptr(), vd, x.data_ptr(), x_bs, x_next.data_ptr(), in_bs, g.data_ptr(),
                                     grad_v.data_ptr(), B, n, C.byref(coefs), flags, _stream_ptr(v.device))
    _cabi.check(rc, "logprob_bwd")
    launch_count += 1
    return grad_v


def _loss_args(old_logp, advantages, stats_rows, clip_range, adv_clip_max, kl_coeff, denom, B, device, accumulate=True):
    keep = []
    def vec(t, name):
        _require_cuda(t, name)
        t = t.detach().to(torch.float32).contiguous().view(-1)
        if t.numel() != B:
            raise ValueError(f"mixgrpo_b200: `{name}` must have one entry per sample")
        keep.append(t)
        return t.data_ptr()
    la = LossArgs()
    la.old_logp, la.advantages = vec(old_logp, "old_log_probs"), vec(advantages, "advantages")
    la.stats_rows = None
    if stats_rows is not None:
        if stats_rows.dtype != torch.float32 or tuple(stats_rows.shape) != (B, 4) or not stats_rows.is_contiguous() or stats_rows.device != device:
            raise ValueError("mixgrpo_b200: stats_rows must be a contiguous fp32 [B, 4] tensor on the same device")
        la.stats_rows = stats_rows.data_ptr()
    la.clip_range, la.adv_clip_max, la.kl_coeff, la.denom = float(clip_range), float(adv_clip_max), float(kl_coeff), float(denom)
    la.accumulate = 1 if accumulate else 0
    return la, keep


def policy_forward(family: int, v: torch.Tensor, x: torch.Tensor, x_next: torch.Tensor, coefs: StepCoefs, old_logp: torch.Tensor,
                   advantages: torch.Tensor, clip_range: float, adv_clip_max: float, kl_coeff: float, denom: float,
                   stats_rows: Optional[torch.Tensor] = None, round_like_torch: bool = False,
                   out_logp: Optional[torch.Tensor] = None, accumulate: bool = True, early_loads: bool = False) -> torch.Tensor:
    """Fused policy-update forward (mixgrpo_policy_fwd): new log-probs [B]; per-sample loss terms += stats_rows.
    ``early_loads``: no input tensor was written by the launch immediately before this one on the stream."""
    global launch_count
    tb = binding()
    if tb is not None:
        launch_count += 1 if v.shape[0] else 0
        return tb.policy_forward(family, v, x, x_next, C.addressof(coefs), old_logp, advantages, float(clip_range), float(adv_clip_max), float(kl_coeff),
                                 float(denom), stats_rows, round_like_torch, out_logp, accumulate, early_loads)
    lib = _cabi.lib()
    for t, nm in ((v, "model_output"), (x, "latents"), (x_next, "prev_sample")):
        _require_cuda(t, nm)
    vd = _dtype_code(v, "model_output")
    v = v if v.is_contiguous() else v.contiguous()
    if v.shape[0] == 0:
        return out_logp if out_logp is not None else torch.empty((0,), dtype=torch.float32, device=v.device)
    B, n, dev = v.shape[0], v[0].numel(), v.device
    x, x_bs = _rows(x.to(torch.float32), "latents")
    x_next, in_bs = _rows(x_next.to(torch.float32), "prev_sample")
    la, keep = _loss_args(old_logp, advantages, stats_rows, clip_range, adv_clip_max, kl_coeff, denom, B, dev, accumulate)
    logp = out_logp if out_logp is not None else torch.empty((B,), dtype=torch.float32, device=dev)
    ws = _workspace(dev, B, n)
    flags = (FLAG_ROUND_LIKE_TORCH if (round_like_torch and vd == BF16) else 0) | (FLAG_PDL_EARLY_LOADS if early_loads else 0)
    with _on_device(dev):
        rc = lib.mixgrpo_policy_fwd(family, v.data_ptr(), vd, x.data_ptr(), x_bs, x_next.data_ptr(), in_bs, logp.data_ptr(), ws.data_ptr(),
                                    ws.numel(), B, n, C.byref(coefs), C.byref(la), flags, _stream_ptr(dev))
    _cabi.check(rc, "policy_fwd")
    launch_count += 1
    del keep
    return logp


def policy_backward(family: int, v: torch.Tensor, x: torch.Tensor, x_next: torch.Tensor, new_logp: torch.Tensor, coefs: StepCoefs,
                    old_logp: torch.Tensor, advantages: torch.Tensor, clip_range: float, adv_clip_max: float, kl_coeff: float,
                    denom: float, round_like_torch: bool = False, early_loads: bool = False) -> torch.Tensor:
    """Fused policy-update backward (mixgrpo_policy_bwd): d loss / d model_output, dtype = model_output.dtype.
    ``early_loads``: the caller launched ``policy_forward`` on the same inputs immediately before (nothing in between
    writes v / x / x_next), so the kernel may load them while that launch drains (MIXGRPO_FLAG_PDL_EARLY_LOADS)."""
    global launch_count
    tb = binding()
    if tb is not None:
        launch_count += 1 if v.shape[0] else 0
        return tb.policy_backward(family, v, x, x_next, new_logp, C.addressof(coefs), old_logp, advantages, float(clip_range), float(adv_clip_max),
                                  float(kl_coeff), float(denom), round_like_torch, early_loads)
    lib = _cabi.lib()
    for t, nm in ((v, "model_output"), (x, "latents"), (x_next, "prev_sample"), (new_logp, "new_log_probs")):
        _require_cuda(t, nm)
    vd = _dtype_code(v, "model_output")
    v = v if v.is_contiguous() else v.contiguous()
    if v.shape[0] == 0:
        return torch.empty_like(v)
    B, n, dev = v.shape[0], v[0].numel(), v.device
    x, x_bs = _rows(x.to(torch.float32), "latents")
    x_next, in_bs = _rows(x_next.to(torch.float32), "prev_sample")
    nl = new_logp.detach().to(torch.float32).contiguous().view(-1)
    la, keep = _loss_args(old_logp, advantages, None, clip_range, adv_clip_max, kl_coeff, denom, B, dev)
    grad_v = torch.empty_like(v)
    flags = (FLAG_ROUND_LIKE_TORCH if (round_like_torch and vd == BF16) else 0) | (FLAG_PDL_EARLY_LOADS if early_loads else 0)
    with _on_device(dev):
        rc = lib.mixgrpo_policy_bwd(family, v.data_ptr(), vd, x.data_ptr(), x_bs, x_next.data_ptr(), in_bs, nl.data_ptr(), C.byref(la),
                                    grad_v.data_ptr(), B, n, C.byref(coefs), flags, _stream_ptr(dev))
    _cabi.check(rc, "policy_bwd")
    launch_count += 1
    del keep
    return grad_v


def _policy_items(family, vs, xs, x_nexts, coefs_list, old_logps, logps, stats_rows, grads, B, n, dev):
    if not (0 < len(vs) <= POLICY_MAX_ITEMS):
        raise ValueError(f"mixgrpo_b200: 1..{POLICY_MAX_ITEMS} items per batched policy launch, got {len(vs)}")
    items = (PolicyItem * len(vs))()
    keep = []
    for j in range(len(vs)):
        v = vs[j]
        for t, nm in ((v, "model_output"), (xs[j], "latents"), (x_nexts[j], "prev_sample")):
            _require_cuda(t, nm)
        if v.dtype != vs[0].dtype or v.shape != vs[0].shape:
            raise ValueError("mixgrpo_b200: every item of a batched policy launch must have the same shape and dtype")
        v = v if v.is_contiguous() else v.contiguous()
        x, x_bs = _rows(xs[j].to(torch.float32), "latents")
        xn, in_bs = _rows(x_nexts[j].to(torch.float32), "prev_sample")
        ol = old_logps[j].detach().to(torch.float32).contiguous().view(-1)
        if ol.numel() != B or logps[j].numel() != B or logps[j].dtype != torch.float32 or not logps[j].is_contiguous():
            raise ValueError("mixgrpo_b200: log-prob vectors must be contiguous fp32 [B]")
        it = items[j]
        it.v, it.x, it.x_next, it.x_bs, it.in_bs = v.data_ptr(), x.data_ptr(), xn.data_ptr(), x_bs, in_bs
        it.logp, it.old_logp = logps[j].data_ptr(), ol.data_ptr()
        it.stats_rows = None
        if stats_rows is not None and stats_rows[j] is not None:
            r = stats_rows[j]
            if r.dtype != torch.float32 or tuple(r.shape) != (B, 4) or not r.is_contiguous() or r.device != dev:
                raise ValueError("mixgrpo_b200: stats_rows must be contiguous fp32 [B, 4] tensors on the same device")
            it.stats_rows = r.data_ptr()
        it.grad_v = grads[j].data_ptr() if grads is not None else None
        it.coefs = coefs_list[j]
        keep += [v, x, xn, ol]
    return items, keep


def policy_forward_multi(family: int, vs, xs, x_nexts, coefs_list, old_logps, advantages: torch.Tensor, clip_range: float,
                         adv_clip_max: float, kl_coeff: float, denom: float, stats_rows=None, round_like_torch: bool = False,
                         out_logps: Optional[torch.Tensor] = None, accumulate: bool = True, early_loads: bool = False):
    """The window's policy forwards as ONE launch (mixgrpo_policy_fwd_multi): item j = (vs[j], xs[j], x_nexts[j], coefs_list[j],
    old_logps[j], stats_rows[j]).  Returns the new log-probs ``[n_items, B]``, or ``None`` — nothing launched — for ragged /
    unaligned tensors (call ``policy_forward`` per item then).  Bit-identical to the per-item launches."""
    global launch_count
    lib = _cabi.lib()
    v0 = vs[0]
    _require_cuda(v0, "model_output")
    vd = _dtype_code(v0, "model_output")
    B, n, dev = v0.shape[0], v0[0].numel(), v0.device
    J = len(vs)
    logps = out_logps if out_logps is not None else torch.empty((J, B), dtype=torch.float32, device=dev)
    if B == 0:
        return logps
    items, keep = _policy_items(family, vs, xs, x_nexts, coefs_list, old_logps, [logps[j] for j in range(J)], stats_rows, None, B, n, dev)
    _require_cuda(advantages, "advantages")
    adv = advantages.detach().to(torch.float32).contiguous().view(-1)
    if adv.numel() != B:
        raise ValueError("mixgrpo_b200: `advantages` must have one entry per sample")
    ws = _workspace(dev, B * J, n)
    flags = (FLAG_ROUND_LIKE_TORCH if (round_like_torch and vd == BF16) else 0) | (FLAG_PDL_EARLY_LOADS if early_loads else 0)
    with _on_device(dev):
        rc = lib.mixgrpo_policy_fwd_multi(family, vd, items, J, adv.data_ptr(), float(clip_range), float(adv_clip_max), float(kl_coeff),
                                          float(denom), 1 if accumulate else 0, ws.data_ptr(), ws.numel(), B, n, flags, _stream_ptr(dev))
    del keep
    if rc == _cabi.EUNSUPPORTED:
        return None
    _cabi.check(rc, "policy_fwd_multi")
    launch_count += 1
    return logps


def policy_backward_multi(family: int, vs, xs, x_nexts, new_logps: torch.Tensor, coefs_list, old_logps, advantages: torch.Tensor,
                          clip_range: float, adv_clip_max: float, kl_coeff: float, denom: float, round_like_torch: bool = False,
                          early_loads: bool = False, out_grads=None):
    """The window's policy backwards as ONE launch (mixgrpo_policy_bwd_multi): ``grads[j] = dloss/d vs[j]``; ``None`` when the
    batched entry point does not cover the tensors."""
    global launch_count
    lib = _cabi.lib()
    v0 = vs[0]
    vd = _dtype_code(v0, "model_output")
    B, n, dev = v0.shape[0], v0[0].numel(), v0.device
    J = len(vs)
    grads = out_grads if out_grads is not None else [torch.empty_like(v0, memory_format=torch.contiguous_format) for _ in range(J)]
    if B == 0:
        return grads
    for g in grads:
        if g.dtype != v0.dtype or g.shape != v0.shape or not g.is_contiguous():
            raise ValueError("mixgrpo_b200: gradient buffers must match the model output's dtype/shape and be contiguous")
    nl = new_logps.detach()
    items, keep = _policy_items(family, vs, xs, x_nexts, coefs_list, old_logps, [nl[j] for j in range(J)], None, grads, B, n, dev)
    adv = advantages.detach().to(torch.float32).contiguous().view(-1)
    flags = (FLAG_ROUND_LIKE_TORCH if (round_like_torch and vd == BF16) else 0) | (FLAG_PDL_EARLY_LOADS if early_loads else 0)
    with _on_device(dev):
        rc = lib.mixgrpo_policy_bwd_multi(family, vd, items, J, adv.data_ptr(), float(clip_range), float(adv_clip_max), float(kl_coeff),
                                          float(denom), B, n, flags, _stream_ptr(dev))
    del keep
    if rc == _cabi.EUNSUPPORTED:
        return None
    _cabi.check(rc, "policy_bwd_multi")
    launch_count += 1
    return grads


def policy_step(family: int, v: torch.Tensor, x: torch.Tensor, x_next: torch.Tensor, coefs: StepCoefs, old_logp: torch.Tensor,
                advantages: torch.Tensor, clip_range: float, adv_clip_max: float, kl_coeff: float, denom: float,
                stats_rows: Optional[torch.Tensor] = None, round_like_torch: bool = False, accumulate: bool = True):
    """Single-pass policy update (mixgrpo_policy_step): ``(new_log_probs [B], grad_model_output)`` in ONE launch that reads
    v / x / x_next once (12 B/elem instead of 22).  Returns ``None`` — nothing launched — when the shape is not covered
    (residuals do not fit on chip, ragged or unaligned tensors); the caller then uses ``policy_forward`` + ``policy_backward``."""
    global launch_count
    lib = _cabi.lib()
    for t, nm in ((v, "model_output"), (x, "latents"), (x_next, "prev_sample")):
        _require_cuda(t, nm)
    vd = _dtype_code(v, "model_output")
    v = v if v.is_contiguous() else v.contiguous()
    if v.shape[0] == 0:
        return torch.empty((0,), dtype=torch.float32, device=v.device), torch.empty_like(v)
    B, n, dev = v.shape[0], v[0].numel(), v.device
    x, x_bs = _rows(x.to(torch.float32), "latents")
    x_next, in_bs = _rows(x_next.to(torch.float32), "prev_sample")
    la, keep = _loss_args(old_logp, advantages, stats_rows, clip_range, adv_clip_max, kl_coeff, denom, B, dev, accumulate)
    logp = torch.empty((B,), dtype=torch.float32, device=dev)
    grad_v = torch.empty_like(v)
    ws = _workspace(dev, B, n)
    flags = FLAG_ROUND_LIKE_TORCH if (round_like_torch and vd == BF16) else 0
    with _on_device(dev):
        rc = lib.mixgrpo_policy_step(family, v.data_ptr(), vd, x.data_ptr(), x_bs, x_next.data_ptr(), in_bs, logp.data_ptr(),
                                     grad_v.data_ptr(), ws.data_ptr(), ws.numel(), B, n, C.byref(coefs), C.byref(la), flags, _stream_ptr(dev))
    del keep
    if rc == _cabi.EUNSUPPORTED:
        return None
    _cabi.check(rc, "policy_step")
    launch_count += 1
    return logp, grad_v


def cast_rows(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """dst[b] = float32(src[b]) for a (B, ...) bf16|f32 contiguous ``src`` and an fp32 ``dst`` view whose per-sample
    block is contiguous (e.g. ``all_latents[:, 0]``): seeds the trajectory buffer (SU:26, SU:153)."""
    global launch_count
    _require_cuda(src, "src")
    _require_cuda(dst, "dst")
    code = _dtype_code(src, "src")
    src = src if src.is_contiguous() else src.contiguous()
    if dst.dtype != torch.float32 or dst.shape != src.shape or not dst[0].is_contiguous():
        raise ValueError("mixgrpo_b200: dst must be fp32, same shape, contiguous per sample")
    B, n = src.shape[0], src[0].numel()
    with _on_device(src.device):
        rc = _cabi.lib().mixgrpo_cast_rows(src.data_ptr(), code, dst.data_ptr(), dst.stride(0) if B > 1 else n, B, n, _stream_ptr(src.device))
    _cabi.check(rc, "cast_rows")
    launch_count += 1
    return dst


def group_advantages(rewards: torch.Tensor, weights: Optional[torch.Tensor], num_generations: int, trim_size: int = 0,
                     use_group: bool = True, stat_rewards: Optional[torch.Tensor] = None) -> torch.Tensor:
    """rewards [n_models, local_B] fp32 -> advantages [local_B] fp32 (TR:439-501)."""
    global launch_count
    lib = _cabi.lib()
    _require_cuda(rewards, "rewards")
    r = rewards.to(torch.float32)
    if r.dim() == 1:
        r = r.unsqueeze(0)
    r = r.contiguous()
    n_models, local_B = r.shape
    if local_B == 0:
        return torch.empty((0,), dtype=torch.float32, device=r.device)
    w_p = None
    if weights is not None:
        weights = weights.to(device=r.device, dtype=torch.float32).contiguous()
        if weights.numel() != n_models:
            raise ValueError("mixgrpo_b200: one weight per reward model required")
        w_p = weights.data_ptr()
    s_p, n_stat = None, 0
    if stat_rewards is not None:
        stat_rewards = stat_rewards.to(device=r.device, dtype=torch.float32).contiguous()
        s_p, n_stat = stat_rewards.data_ptr(), stat_rewards.numel()
    # every entry is written unless a ragged tail (local_B % G) exists, which the reference leaves at zero (TR:445)
    full = (not use_group) or (local_B % int(num_generations) == 0)
    adv = (torch.empty if full else torch.zeros)((local_B,), dtype=torch.float32, device=r.device)
    with _on_device(r.device):
        rc = lib.mixgrpo_group_advantages(r.data_ptr(), w_p, n_models, local_B, int(num_generations), int(trim_size),
                                          1 if use_group else 0, s_p, n_stat, adv.data_ptr(), _stream_ptr(r.device))
    _cabi.check(rc, "group_advantages")
    launch_count += 1
    return adv


def grpo_loss_fwd_bwd(new_logp: torch.Tensor, old_logp: torch.Tensor, advantages: torch.Tensor, clip_range: float,
                      adv_clip_max: float, kl_coeff: float, denom: float, want_grad: bool = True,
                      stats_accum: Optional[torch.Tensor] = None):
    """Returns (stats[4] = loss, policy, kl, clip_frac ; grad_new_logp or None).  TR:560-583."""
    global launch_count
    lib = _cabi.lib()
    for t, nm in ((new_logp, "new_log_probs"), (old_logp, "old_log_probs"), (advantages, "advantages")):
        _require_cuda(t, nm)
    nl = new_logp.detach().to(torch.float32).contiguous().view(-1)
    ol = old_logp.detach().to(torch.float32).contiguous().view(-1)
    ad = advantages.detach().to(torch.float32).contiguous().view(-1)
    B = nl.numel()
    if ol.numel() != B or ad.numel() != B:
        raise ValueError("mixgrpo_b200: new/old log-probs and advantages must have the same length")
    stats = torch.empty((4,), dtype=torch.float32, device=nl.device)
    grad = torch.empty((B,), dtype=torch.float32, device=nl.device) if want_grad else None
    acc_p = None
    if stats_accum is not None:
        if stats_accum.dtype != torch.float32 or stats_accum.numel() != 4 or not stats_accum.is_contiguous():
            raise ValueError("mixgrpo_b200: stats_accum must be a contiguous fp32 [4] tensor")
        acc_p = stats_accum.data_ptr()
    with _on_device(nl.device):
        rc = lib.mixgrpo_grpo_loss(nl.data_ptr(), ol.data_ptr(), ad.data_ptr(), B, float(clip_range), float(adv_clip_max),
                                   float(kl_coeff), float(denom), stats.data_ptr(), grad.data_ptr() if want_grad else None,
                                   acc_p, _stream_ptr(nl.device))
    _cabi.check(rc, "grpo_loss")
    launch_count += 1
    return stats, grad


def pack_latents(latents: torch.Tensor, batch_size: int, num_channels: int, height: int, width: int) -> torch.Tensor:
    """(B,C,H,W) -> (B,(H/2)(W/2),4C), TR:94-99."""
    global launch_count
    _require_cuda(latents, "latents")
    code = _dtype_code(latents, "latents")
    src = latents.contiguous().view(batch_size, num_channels, height, width)
    dst = torch.empty((batch_size, (height // 2) * (width // 2), num_channels * 4), dtype=latents.dtype, device=latents.device)
    with _on_device(latents.device):
        rc = _cabi.lib().mixgrpo_pack_latents(src.data_ptr(), dst.data_ptr(), code, batch_size, num_channels, height, width,
                                              _stream_ptr(latents.device))
    _cabi.check(rc, "pack_latents")
    launch_count += 1
    return dst


def unpack_latents(latents: torch.Tensor, height: int, width: int, vae_scale_factor: int, divisor: float = 1.0,
                   shift: float = 0.0) -> torch.Tensor:
    """(B,S,4C) -> (B,C,H',W') with H' = 2*(height // (2*vae_scale_factor)), TR:102-115; optional fused
    ``x / divisor + shift`` (TR:287)."""
    global launch_count
    _require_cuda(latents, "latents")
    code = _dtype_code(latents, "latents")
    b, _, ch = latents.shape
    hh = 2 * (int(height) // (vae_scale_factor * 2))
    ww = 2 * (int(width) // (vae_scale_factor * 2))
    src = latents.contiguous()
    dst = torch.empty((b, ch // 4, hh, ww), dtype=latents.dtype, device=latents.device)
    with _on_device(latents.device):
        rc = _cabi.lib().mixgrpo_unpack_latents(src.data_ptr(), dst.data_ptr(), code, b, ch // 4, hh, ww, float(divisor),
                                                float(shift), _stream_ptr(latents.device))
    _cabi.check(rc, "unpack_latents")
    launch_count += 1
    return dst
