"""Sliding-window scheduler: which denoising steps are stochastic (SDE) and trained this iteration.

Drop-in for ``/root/reference/fastvideo/utils/grpo_states.py:7-159`` (same class name, constructor
fields, methods and integer behaviour — checked step by step against the reference in
tests/test_grpo_states.py and the golden traces in tests/golden/grpo_states_traces.json).  Host integers
only; nothing here touches the GPU.

The window covers ``range(cur_timestep, min(cur_timestep + group_size, max_timesteps))``; it advances
every ``iters_per_group`` calls to ``update_iteration`` (a fixed, linearly decaying or exponentially
decaying budget depending on ``sample_strategy``), by ``prog_overlap_step`` when ``prog_overlap`` else by
``group_size``; past the end it either clips at ``max_timesteps`` or rolls back to the initial step.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

_ADVANCING = ("progressive", "decay", "exp_decay")


@dataclass
class GRPOTrainingStates:
    iters_per_group: int
    group_size: int
    max_timesteps: int
    cur_timestep: int = 0
    cur_iter_in_group: int = 0
    sample_strategy: str = "progressive"
    prog_overlap: bool = False
    prog_overlap_step: int = 1
    max_iters_per_group: Optional[int] = None
    min_iters_per_group: Optional[int] = None
    roll_back: bool = False
    exp_decay_thre_timestep: int = 13
    exp_decay_k: float = 0.1

    def __post_init__(self):
        self.init_timestep = self.cur_timestep                    # where roll_back_start() returns to
        if self.sample_strategy == "decay":                       # defaults of the decaying budget (grpo_states.py:48-52)
            hi, lo = self.max_iters_per_group, self.min_iters_per_group
            self.max_iters_per_group = self.iters_per_group if hi is None else hi
            self.min_iters_per_group = max(1, self.iters_per_group // 4) if lo is None else lo

    def set_params(self, params: dict):
        """Overwrite any attributes by name (the reference's resume hook)."""
        for name in params:
            setattr(self, name, params[name])

    # ---- per-window iteration budget -------------------------------------------------------
    def get_dynamic_iters_per_group(self) -> int:
        """Linear interpolation max→min over the schedule ("decay", grpo_states.py:55-67)."""
        if self.sample_strategy != "decay":
            return self.iters_per_group
        frac = self.cur_timestep / self.max_timesteps
        budget = int(self.max_iters_per_group * (1 - frac) + self.min_iters_per_group * frac)
        return max(self.min_iters_per_group, budget)

    def get_exp_decay_iters_per_group(self):
        """iters * exp(-k * relu(t - threshold)), rounded up ("exp_decay", grpo_states.py:69-83)."""
        if self.sample_strategy != "exp_decay":
            return self.iters_per_group
        excess = max(0, self.cur_timestep - self.exp_decay_thre_timestep)
        return np.ceil(self.iters_per_group * np.exp(-self.exp_decay_k * excess))

    def _budget(self):
        if self.sample_strategy == "decay":
            return self.get_dynamic_iters_per_group()
        if self.sample_strategy == "exp_decay":
            return self.get_exp_decay_iters_per_group()
        return self.iters_per_group

    # ---- state machine ----------------------------------------------------------------------
    def update_iteration(self, seed=None) -> None:
        """grpo_states.py:85-133."""
        if self.sample_strategy == "random":
            rng = np.random.default_rng(seed)
            self.cur_timestep = rng.integers(0, self.max_timesteps - self.group_size + 1)
            return
        if self.sample_strategy not in _ADVANCING:
            raise ValueError(f"Invalid sample strategy: {self.sample_strategy}")
        self.cur_iter_in_group += 1
        if self.cur_iter_in_group >= self._budget():
            self.cur_iter_in_group = 0
            self.cur_timestep += self.prog_overlap_step if self.prog_overlap else self.group_size
        if self.cur_timestep > self.max_timesteps:
            if self.roll_back:
                self.roll_back_start()
            else:
                self.cur_timestep = self.max_timesteps

    def roll_back_start(self) -> None:
        self.cur_timestep = self.init_timestep
        self.cur_iter_in_group = 0

    def get_current_timesteps(self) -> List[int]:
        """grpo_states.py:141-148."""
        return list(range(self.cur_timestep, min(self.cur_timestep + self.group_size, self.max_timesteps)))

    def is_training_complete(self) -> bool:
        """Only the two strategies that walk the schedule once ever finish (grpo_states.py:150-159)."""
        return self.sample_strategy in ("progressive", "decay") and self.cur_timestep >= self.max_timesteps
