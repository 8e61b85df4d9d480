"""TEST INFRASTRUCTURE ONLY — executes the reference's *inline* GRPO arithmetic where it lies.

``fastvideo/train_grpo_flux.py`` cannot be imported in this image (accelerate / diffusers / reward-model
packages are absent), and the advantage (TR:439-501) and loss (TR:560-583) computations are statements
inside ``train_one_step`` rather than functions.  To pin ``oracle/grpo_oracle.py`` against the reference's
own code — not against our reading of it — this module parses the file with ``ast`` (no import, no
execution of anything else), cuts out exactly those statements, and executes them in a namespace we
supply.  Nothing is copied into the repo; when no reference tree is present the functions return None.
"""
from __future__ import annotations

import ast
import types
from typing import Dict, Optional

import torch

from . import ref_loader

_LOSS_NAMES = {"advantages", "ratio", "unclipped_loss", "clipped_loss", "clip_frac", "policy_loss", "kl_loss", "loss"}
_cache: Dict[str, object] = {}


def _train_one_step_ast():
    if "fn" in _cache:
        return _cache["fn"]
    root = ref_loader.reference_root()
    path = None if root is None else root / "fastvideo" / "train_grpo_flux.py"
    if path is None or not path.is_file():
        _cache["fn"] = None
        return None
    tree = ast.parse(path.read_text())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "train_one_step")
    _cache["fn"] = fn
    return fn


def _is_args_attr(node, attr: str) -> bool:
    return isinstance(node, ast.Attribute) and node.attr == attr and isinstance(node.value, ast.Name) and node.value.id == "args"


def reference_advantages(rewards, gathered_reward, *, use_group: bool, num_generations: int, trimmed_ratio: float,
                         multi_reward_mix: str, reward_weights: Optional[dict]) -> Optional[torch.Tensor]:
    """Run the reference's own ``if args.use_group: ... else: ...`` block (TR:440-501).  ``rewards`` is what
    ``samples["rewards"]`` holds there: a dict of tensors (advantage_aggr) or a tensor (reward_aggr)."""
    fn = _train_one_step_ast()
    if fn is None:
        return None
    block = next(n for n in fn.body if isinstance(n, ast.If) and _is_args_attr(n.test, "use_group") and n.orelse)
    code = compile(ast.Module(body=[block], type_ignores=[]), "<TR:440-501>", "exec")
    ns = {"torch": torch, "samples": {"rewards": rewards}, "gathered_reward": gathered_reward, "reward_weights": reward_weights,
          "args": types.SimpleNamespace(use_group=use_group, num_generations=num_generations, trimmed_ratio=trimmed_ratio,
                                        multi_reward_mix=multi_reward_mix)}
    exec(code, ns)
    return ns["samples"]["advantages"]


def reference_loss(new_log_probs: torch.Tensor, old_log_probs: torch.Tensor, advantages: torch.Tensor, *, clip_range: float,
                   adv_clip_max: float, kl_coeff: float, gradient_accumulation_steps: int, n_train_timesteps: int):
    """Run the reference's own loss statements (TR:560-583) for one (sample batch, step).  Returns
    (loss, policy_loss, kl_loss, clip_frac) or None."""
    fn = _train_one_step_ast()
    if fn is None:
        return None
    inner = None
    for node in ast.walk(fn):
        if isinstance(node, ast.For) and isinstance(node.target, ast.Name) and node.target.id == "_":
            inner = node
            break
    stmts = [s for s in inner.body if isinstance(s, ast.Assign) and len(s.targets) == 1 and isinstance(s.targets[0], ast.Name)
             and s.targets[0].id in _LOSS_NAMES]
    code = compile(ast.Module(body=stmts, type_ignores=[]), "<TR:560-583>", "exec")
    ns = {"torch": torch, "new_log_probs": new_log_probs, "clip_range": clip_range, "adv_clip_max": adv_clip_max, "_": 0,
          "sample": {"advantages": advantages, "log_probs": old_log_probs.reshape(-1, 1)},
          "train_timesteps": list(range(n_train_timesteps)),
          "args": types.SimpleNamespace(gradient_accumulation_steps=gradient_accumulation_steps, kl_coeff=kl_coeff)}
    exec(code, ns)
    return ns["loss"], ns["policy_loss"], ns["kl_loss"], ns["clip_frac"]
