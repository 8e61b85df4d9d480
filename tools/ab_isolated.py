#!/usr/bin/env python
"""The deferred Euler-ODE step WITHOUT programmatic dependent launch (mixgrpo_set_tuning key 1 = 0): every launch waits for the one
before to drain, as after a DiT forward (non-PDL kernels) in a real rollout.  256-thread vs 128-thread CTAs (key 6 = 0 | 2), cold
inputs (rotating buffer sets), bench.py's timing method."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from mixgrpo_b200 import _cabi, coefs, ops  # noqa: E402
from mixgrpo_b200._cabi import SRC_DETERMINISTIC, SRC_NOISE  # noqa: E402

dev = torch.device("cuda:0")
lib = _cabi.lib()
peak = bench.load_peaks()[0]
S, C, NL = 4096, 64, 25
sig = torch.linspace(1, 0, 26)
sig = (3.0 * sig) / (1 + 2.0 * sig)
k, _ = coefs.flow(sig, 9, 0.7, "ref_cuda", True)
for B in (12, 24):
    ns = 10 if B <= 12 else 8
    g = torch.Generator(device=dev).manual_seed(7)
    xs = [torch.randn(B, S, C, device=dev, generator=g) for _ in range(ns)]
    vs = [torch.randn(B, S, C, device=dev, generator=g).bfloat16() for _ in range(ns)]
    es = [torch.randn(B, S, C, device=dev, generator=g).bfloat16() for _ in range(ns)]
    outs = [torch.empty(B, S, C, device=dev) for _ in range(ns)]
    acc = ops.DeferredLogProbs(dev, NL, B, S * C)
    lp = torch.empty(NL, B, device=dev)
    s = torch.cuda.Stream(device=dev)
    e = B * S * C
    for name, bpe, kw in (("ode", 10, lambda i: dict(src=SRC_DETERMINISTIC)), ("sde", 12, lambda i: dict(src=SRC_NOISE, noise=es[i]))):
        for pdl in (0, 1):
            row = []
            for half in (0, 2, 0, 2):
                lib.mixgrpo_set_tuning(1, pdl)
                lib.mixgrpo_set_tuning(6, half)

                def run():
                    for j in range(NL):
                        i = j % ns
                        ops.fused_step(ops.FLOW, vs[i], xs[i], k, out_x_next=outs[i], want_x0=False, defer=acc.slot(j, k), round_like_torch=True, early=1, **kw(i))
                    acc.finalize(lp)
                row.append(round(bench._time_graph(run, NL, s), 3))
            print(f"B={B} {name} pdl={pdl}: 256-thread {row[0]} {row[2]} us | 128-thread {row[1]} {row[3]} us   "
                  f"({e * bpe / min(row[0], row[2]) / 1e3 / peak:.3f} | {e * bpe / min(row[1], row[3]) / 1e3 / peak:.3f} of peak)", flush=True)
    del xs, vs, es, outs
    torch.cuda.empty_cache()
lib.mixgrpo_set_tuning(1, 1)
lib.mixgrpo_set_tuning(6, 1)
