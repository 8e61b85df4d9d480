#!/usr/bin/env python
"""A/B on a B200 of the CTA-shape knobs (mixgrpo_set_tuning keys 6 and 7) with bench.py's own per-kernel method:
key 6 = deferred step launches as 128-thread half-tile CTAs (0 never | 1 auto | 2 always), key 7 = CTA size of the backward kernels."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from mixgrpo_b200 import _cabi  # noqa: E402

dev = torch.device("cuda:0")
peak = bench.load_peaks()[0]
lib = _cabi.lib()
out = {}
for B, S in ((12, 4096), (24, 4096)):
    for half in (0, 2):
        for thr in (256, 128):
            lib.mixgrpo_set_tuning(6, half)
            lib.mixgrpo_set_tuning(7, thr)
            r = bench.measure_kernels(dev, peak, B, S, full=True)
            key = f"B{B}_S{S}_half{half}_bwd{thr}"
            out[key] = {k: r[k]["us_per_launch"] for k in ("ode", "sde", "sde_x0", "bwd_x4 (one launch)", "bwd", "train_fwd_x4 (one launch)")}
            print(key, out[key], flush=True)
            torch.cuda.empty_cache()
# whole steps (one CUDA graph each, bench.py's `configs` method): the shapes as a step really meets them
for name in ("mixgrpo", "flash", "large_b24_1024sq", "large_b24_512sq_bf16", "large_b24_512sq_f32"):
    for half in (0, 1, 2):
        for thr in (256, 128):
            lib.mixgrpo_set_tuning(6, half)
            lib.mixgrpo_set_tuning(7, thr)
            line, _ = bench.time_scenario(name, dev, 0, 1, 50, None)
            key = f"step_{name}_half{half}_bwd{thr}"
            out[key] = line["ms_per_step"]
            print(key, out[key], flush=True)
            torch.cuda.empty_cache()
lib.mixgrpo_set_tuning(6, 1)
lib.mixgrpo_set_tuning(7, 256)
if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text(json.dumps(out, indent=1))
