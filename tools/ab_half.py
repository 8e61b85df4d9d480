#!/usr/bin/env python
"""Interleaved A/B of mixgrpo_set_tuning key 6 (0 never | 1 auto | 2 always) on whole bench steps, several rounds in shuffled order."""
import random
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from mixgrpo_b200 import _cabi  # noqa: E402

dev = torch.device("cuda:0")
lib = _cabi.lib()
random.seed(1)
names = sys.argv[1:] or ["mixgrpo"]
for name in names:
    res = {0: [], 1: [], 2: []}
    for rnd in range(4):
        order = [0, 1, 2]
        random.shuffle(order)
        for half in order:
            lib.mixgrpo_set_tuning(6, half)
            before = lib.mixgrpo_set_tuning(8, 0)
            line, _ = bench.time_scenario(name, dev, 0, 1, 100, None)
            res[half].append((line["ms_per_step"], lib.mixgrpo_set_tuning(8, 0) - before))
    lib.mixgrpo_set_tuning(6, 1)
    for half, v in res.items():
        print(name, "key6 =", half, " ms/step:", [a for a, _ in v], " 128-thread launches while capturing:", [b for _, b in v], flush=True)
