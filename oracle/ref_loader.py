"""TEST INFRASTRUCTURE ONLY — loads the *unmodified* reference operators for pinning.

The reference (zqqqqz2000/MixGRPO) is pure Python.  Its sampler/log-prob operator
file ``fastvideo/utils/sampling_utils.py`` has one import that is absent from this
image (``diffusers.utils.torch_utils.randn_tensor``, sampling_utils.py:3).  We put a
stub module in ``sys.modules`` whose ``randn_tensor`` pops a caller-supplied noise
tensor from a queue, so the reference's rollout mode can be driven with *explicit*
noise.  Nothing is copied: the file is executed from where it lies.

Search order for the reference tree: ``$MIXGRPO_REF_ROOT``, ``/root/reference``,
``<repo>/baseline/_ref`` (a git-ignored drop used once for the on-GPU probe,
tools/probe_cuda_rounding.py).  ``load()`` returns ``None`` when no tree is found —
the GPU box has none, and nothing in ``-m gpu`` tests, smoke() or bench.py needs it.

Only tests/ and tools/ (fixture generators) import this module; the product package
``mixgrpo_b200`` never does.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from pathlib import Path
from typing import List, Optional

import torch

_REPO = Path(__file__).resolve().parent.parent
NOISE_QUEUE: List[torch.Tensor] = []   # explicit-noise queue consumed by the stubbed randn_tensor
_cached = {}


def _candidates():
    env = os.environ.get("MIXGRPO_REF_ROOT")
    if env:
        yield Path(env)
    yield Path("/root/reference")
    yield _REPO / "baseline" / "_ref"


def reference_root() -> Optional[Path]:
    for c in _candidates():
        if (c / "fastvideo" / "utils" / "sampling_utils.py").is_file():
            return c
    return None


def _stub_randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    if not NOISE_QUEUE:
        raise RuntimeError("oracle.ref_loader: reference asked for noise but NOISE_QUEUE is empty")
    t = NOISE_QUEUE.pop(0)
    assert tuple(t.shape) == tuple(shape), (t.shape, shape)
    return t.to(device=device, dtype=dtype)


def _exec_file(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, str(path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load():
    """Return the reference ``sampling_utils`` module (or None if no reference tree here)."""
    if "su" in _cached:
        return _cached["su"]
    root = reference_root()
    if root is None:
        return None
    if "diffusers.utils.torch_utils" not in sys.modules:
        d, du, dut = (types.ModuleType(n) for n in ("diffusers", "diffusers.utils", "diffusers.utils.torch_utils"))
        dut.randn_tensor = _stub_randn_tensor
        d.utils, du.torch_utils = du, dut
        sys.modules.update({"diffusers": d, "diffusers.utils": du, "diffusers.utils.torch_utils": dut})
    else:  # a real or earlier stub: force ours so noise is explicit
        sys.modules["diffusers.utils.torch_utils"].randn_tensor = _stub_randn_tensor
    su = _exec_file("_mixgrpo_ref_sampling_utils", root / "fastvideo" / "utils" / "sampling_utils.py")
    su.randn_tensor = _stub_randn_tensor
    _cached["su"] = su
    return su


def load_states():
    """Return the reference ``grpo_states`` module (numpy only) or None."""
    if "st" in _cached:
        return _cached["st"]
    root = reference_root()
    if root is None:
        return None
    _cached["st"] = _exec_file("_mixgrpo_ref_grpo_states", root / "fastvideo" / "utils" / "grpo_states.py")
    return _cached["st"]


def load_reward_utils():
    """Return the reference ``models/reward_model/utils.py`` (balance_pos_neg) or None."""
    if "ru" in _cached:
        return _cached["ru"]
    root = reference_root()
    if root is None:
        return None
    _cached["ru"] = _exec_file("_mixgrpo_ref_reward_utils", root / "fastvideo" / "models" / "reward_model" / "utils.py")
    return _cached["ru"]
