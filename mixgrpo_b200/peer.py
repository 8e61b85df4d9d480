"""Fused peer-memory exchange (one node, NVLink/NVSwitch): reward gather + advantages and the logging all-reduce as
ONE kernel each, with no NCCL call and no extra launch (``csrc/peer_kernels.cu``).

``TR`` = ``/root/reference/fastvideo/train_grpo_flux.py``.  The reference gathers every reward model's ``[local_B]``
vector with a list all-gather + ``torch.cat`` (TR:332-338, TR:417-425), computes advantages with ~10 eager kernels per
group per model (TR:439-501) and all-reduces four logging scalars per (sample, window step) (TR:586-600).  With a
:class:`PeerExchange` a rank does

    px = PeerExchange(group)                               # once: region per rank, CUDA-IPC handles exchanged
    adv, gathered = px.gather_advantages(rewards, 12, weights)       # one launch, replaces TR:417-501
    stats = px.allreduce_stats(stats_rows.sum(0))                    # one launch, replaces TR:586-600

torch.distributed is only used at construction (to ship 64-byte IPC handles and for the closing barrier); the steady
state is kernels pushing floats into peers' memory.  Every launch is CUDA-graph capturable.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.distributed as dist

from . import _cabi
from . import ops as _ops
from ._cabi import ADV_GLOBAL, ADV_GROUP_LOCAL, ADV_GROUP_SPLIT, PEER_HANDLE_BYTES, PEER_MAX_WORLD

_MODES = {"local": ADV_GROUP_LOCAL, "split": ADV_GROUP_SPLIT, "global": ADV_GLOBAL}


class PeerExchange:
    """One rank's endpoint of the peer-memory exchange.

    ``group``: a torch.distributed process group whose ranks all live on ONE node (default group when None).  Without
    an initialised process group the exchange degenerates to world size 1 (same kernels, no peers).
    ``cap_floats``: capacity of a rank's reward slot, ``>= n_models * local_B``."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, cap_floats: int = 4096,
                 device: Optional[torch.device] = None, *, _regions: Optional[Sequence[int]] = None,
                 _rank: Optional[int] = None):
        if os.environ.get("MIXGRPO_PEER_DISABLE") == "1" and _regions is None:
            # operator switch (same value on every rank): callers fall back to NCCL (grpo.gather_rewards / reduce_step_stats)
            raise RuntimeError("mixgrpo_b200: PeerExchange disabled by MIXGRPO_PEER_DISABLE=1")
        self._lib = _cabi.lib()
        self.cap = int(cap_floats)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if self.device.type != "cuda":
            raise RuntimeError("mixgrpo_b200: PeerExchange needs a CUDA device — this package has no CPU fallback")
        self._owned: Optional[int] = None
        self._opened: List[int] = []
        self._group = group
        self._distributed = False
        if _regions is not None:                      # several "ranks" inside one process (tests / single-GPU drivers)
            self.world, self.rank = len(_regions), int(_rank)
            self._ptrs = [int(p) for p in _regions]
        else:
            distributed = dist.is_available() and dist.is_initialized()
            self.world = dist.get_world_size(group) if distributed else 1
            self.rank = dist.get_rank(group) if distributed else 0
            if self.world > PEER_MAX_WORLD:
                raise ValueError(f"mixgrpo_b200: PeerExchange supports up to {PEER_MAX_WORLD} ranks of one node")
            handle = (C.c_ubyte * PEER_HANDLE_BYTES)()
            region = C.c_void_p()
            err: Optional[str] = None
            with torch.cuda.device(self.device):
                rc = self._lib.mixgrpo_peer_region_alloc(self.world, self.cap, C.byref(region), handle)
            if rc != 0:
                err = f"rank {self.rank}: peer_region_alloc failed with code {rc}"
                if self.world == 1:
                    raise RuntimeError("mixgrpo_b200: " + err)
            self._owned = region.value
            self._ptrs = [0] * self.world
            self._ptrs[self.rank] = region.value
            if self.world > 1:
                # every step below is collective and every rank reaches every step, whatever failed locally: a rank that
                # cannot allocate or map reports it, and ALL ranks raise together instead of deadlocking in a barrier
                handles: List[Optional[bytes]] = [None] * self.world
                dist.all_gather_object(handles, None if err else bytes(handle), group=group)
                if err is None and all(h is not None for h in handles):
                    with torch.cuda.device(self.device):
                        for q, hb in enumerate(handles):
                            if q == self.rank:
                                continue
                            buf = (C.c_ubyte * PEER_HANDLE_BYTES).from_buffer_copy(hb)
                            peer = C.c_void_p()
                            rc = self._lib.mixgrpo_peer_region_open(buf, C.byref(peer))
                            if rc != 0:
                                err = f"rank {self.rank}: peer_region_open(rank {q}) failed with code {rc} ({self._lib.mixgrpo_error_string(rc).decode()})"
                                break
                            self._opened.append(peer.value)
                            self._ptrs[q] = peer.value
                elif err is None:
                    err = f"rank {self.rank}: a peer could not allocate its region"
                errs: List[Optional[str]] = [None] * self.world
                dist.all_gather_object(errs, err, group=group)    # also the "everything is mapped everywhere" rendezvous
                if any(errs):
                    self._distributed = True
                    self.close()
                    raise RuntimeError("mixgrpo_b200: PeerExchange setup failed: " + "; ".join(e for e in errs if e))
                self._distributed = True
        self._regions_c = (C.c_void_p * self.world)(*self._ptrs)

    # ------------------------------------------------------------------ single-process construction
    @classmethod
    def local_ranks(cls, world: int, device: Optional[torch.device] = None, cap_floats: int = 4096) -> List["PeerExchange"]:
        """``world`` endpoints on ONE device in this process (regions are plain device allocations; drive each endpoint
        on its own CUDA stream).  Used by the single-GPU parity tests of the protocol."""
        device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        lib = _cabi.lib()
        ptrs = []
        with torch.cuda.device(device):
            for _ in range(world):
                region = C.c_void_p()
                _cabi.check(lib.mixgrpo_peer_region_alloc(world, cap_floats, C.byref(region), None), "peer_region_alloc")
                ptrs.append(region.value)
        eps = [cls(None, cap_floats, device, _regions=ptrs, _rank=r) for r in range(world)]
        for r, e in enumerate(eps):
            e._owned = ptrs[r]
        return eps

    # ------------------------------------------------------------------ the two exchanges
    def gather_advantages(self, rewards: Union[torch.Tensor, Dict[str, torch.Tensor]], num_generations: int,
                          reward_weights: Optional[Union[Dict[str, float], Sequence[float], torch.Tensor]] = None,
                          trimmed_ratio: float = 0.0, mode: str = "local",
                          want_gathered: bool = True) -> Tuple[torch.Tensor, Optional[Union[torch.Tensor, Dict[str, torch.Tensor]]]]:
        """All-gather the rewards AND compute this rank's advantages in one launch.

        ``rewards``: dict ``{model: [local_B]}`` (advantage_aggr, TR:417-422) or a tensor ``[local_B]`` / ``[n_models,
        local_B]`` (reward_aggr, TR:423-425).  ``mode``: ``"local"`` — groups are runs of ``num_generations`` of this
        rank's samples (the reference, TR:443-461); ``"split"`` — runs in the rank-major gathered order, a group may
        span ranks (SURVEY §8e extended mode); ``"global"`` — ``use_group=False`` statistics of the gathered vector
        (TR:495-499).  Returns ``(advantages [local_B], gathered)`` where ``gathered`` has the container type of
        ``rewards`` with entries of length ``world*local_B`` in ``torch.cat`` order (TR:338), or None."""
        from . import grpo as _grpo
        mat, names = _grpo.stack_rewards(rewards)
        _ops._require_cuda(mat, "rewards")
        mat = mat.contiguous()
        n_models, local_B = mat.shape
        if mode not in _MODES:
            raise ValueError("mode must be 'local', 'split' or 'global'")
        w_p = None
        if mode == "global":
            if names is not None:
                raise ValueError("multi_reward_mix 'advantage_aggr' is not supported when use_group is False.")   # TR:496
        elif torch.is_tensor(reward_weights):
            weights = reward_weights.to(device=mat.device, dtype=torch.float32).contiguous()
            w_p = weights.data_ptr()
        elif names is not None or (reward_weights is not None and n_models > 1):
            if reward_weights is None:
                raise ValueError("reward_weights is required for multi-reward (advantage_aggr) rewards")
            w = tuple(float(reward_weights[k]) for k in names) if isinstance(reward_weights, dict) else tuple(float(x) for x in reward_weights)
            weights = _grpo._weights_on(mat.device, w)
            w_p = weights.data_ptr()
        trim = 0
        if trimmed_ratio > 0 and mode != "global":                              # TR:451-454
            trim = min(int(num_generations * trimmed_ratio), num_generations - 1)
        adv = torch.empty((local_B,), dtype=torch.float32, device=mat.device)
        gathered = torch.empty((n_models, self.world * local_B), dtype=torch.float32, device=mat.device) if want_gathered else None
        with torch.cuda.device(mat.device):
            rc = self._lib.mixgrpo_peer_gather_advantages(self._regions_c, self.rank, self.world, self.cap, mat.data_ptr(), w_p,
                                                          n_models, local_B, int(num_generations), trim, _MODES[mode],
                                                          gathered.data_ptr() if want_gathered else None, adv.data_ptr(),
                                                          _ops._stream_ptr(mat.device))
        _cabi.check(rc, "peer_gather_advantages")
        _ops.launch_count += 1
        if gathered is None:
            return adv, None
        if names is not None:
            return adv, {k: gathered[i] for i, k in enumerate(names)}
        if isinstance(rewards, torch.Tensor) and rewards.dim() == 1:
            return adv, gathered.reshape(-1)
        return adv, gathered

    def allreduce_stats(self, values: torch.Tensor, average: bool = True) -> torch.Tensor:
        """In-place sum (in rank order — same bits on every rank) or average of ``values`` (contiguous fp32) over the ranks:
        the ONE reduction per ``train_one_step`` that replaces TR:586-600's four all-reduce + ``.item()`` pairs per (sample,
        window step).  One launch per 256 entries (the kernel's message size); the usual [4] or [window, B, 4] sums are one."""
        _ops._require_cuda(values, "values")
        if values.dtype != torch.float32 or not values.is_contiguous() or values.numel() < 1:
            raise ValueError("mixgrpo_b200: allreduce_stats takes a contiguous fp32 tensor with at least one entry")
        flat = values.view(-1)
        with torch.cuda.device(values.device):
            st = _ops._stream_ptr(values.device)
            for lo in range(0, flat.numel(), 256):
                cnt = min(256, flat.numel() - lo)
                rc = self._lib.mixgrpo_peer_allreduce(self._regions_c, self.rank, self.world, self.cap, flat.data_ptr() + 4 * lo, cnt,
                                                      1 if average else 0, st)
                _cabi.check(rc, "peer_allreduce")
                _ops.launch_count += 1
        return values

    # ------------------------------------------------------------------ introspection / teardown
    def status(self) -> Tuple[int, int, int]:
        """(gather calls completed, all-reduce calls completed, status) of this rank's region; synchronises."""
        a, b, s = C.c_int(), C.c_int(), C.c_int()
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.mixgrpo_peer_region_status(self._ptrs[self.rank], C.byref(a), C.byref(b), C.byref(s)), "peer_region_status")
        return a.value, b.value, s.value

    def close(self) -> None:
        """Unmap the peers' regions and free our own (collective when created from a process group: nobody frees a
        region a peer's kernel might still push into)."""
        if self._ptrs is None:
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for p in self._opened:
                self._lib.mixgrpo_peer_region_close(p)
            self._opened = []
            if self._distributed:
                dist.barrier(group=self._group)
            if self._owned is not None:
                self._lib.mixgrpo_peer_region_free(self._owned)
                self._owned = None
        self._ptrs = None
