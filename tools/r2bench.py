#!/usr/bin/env python
"""Round-2 kernel timings on a B200: per-launch CUDA-event times of the sampler-step / policy kernels over rotating buffer
sets (> L2), inside CUDA graphs with programmatic dependent launch — the same method as bench.py's roofline leg.

  python tools/r2bench.py [--groups 12 24 36] [--json out.json]

Rows: ode / sde / train_fwd / bwd as single launches (early = 0 | 1 | 2), the window's 4 forwards / 4 backwards as 4
launches vs ONE batched launch (mixgrpo_policy_fwd_multi / _bwd_multi), and the in-kernel-noise SDE step."""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

S, C = 4096, 64
PEAK = 6533.5


def timed_graph(fn, n_launch, stream, reps=20):
    with torch.cuda.stream(stream):
        fn()
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            fn()
        for _ in range(3):
            g.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            g.replay()
        b.record(stream)
        b.synchronize()
    return a.elapsed_time(b) * 1e3 / (reps * n_launch)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--groups", type=int, nargs="+", default=[12, 24, 36])
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    from mixgrpo_b200 import coefs, ops
    from mixgrpo_b200._cabi import SRC_DETERMINISTIC, SRC_GIVEN, SRC_NOISE, SRC_PHILOX
    dev = torch.device("cuda:0")
    try:
        PEAKS = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
    except Exception:  # noqa: BLE001
        PEAKS = PEAK
    sig = torch.linspace(1, 0, 26)
    sig = (3.0 * sig) / (1 + 2.0 * sig)
    ks = [coefs.flow(sig, t, 0.7, "ref_cuda", True)[0] for t in range(4)]
    k = coefs.flow(sig, 9, 0.7, "ref_cuda", True)[0]
    out = {}
    stream = torch.cuda.Stream(device=dev)
    for B in args.groups:
        ns = 12 if B <= 12 else (8 if B <= 24 else 6)
        g = torch.Generator(device=dev).manual_seed(7)
        xs = [torch.randn(B, S, C, device=dev, generator=g) for _ in range(ns)]
        vs = [torch.randn(B, S, C, device=dev, generator=g).bfloat16() for _ in range(ns)]
        es = [torch.randn(B, S, C, device=dev, generator=g).bfloat16() for _ in range(ns)]
        outs = [torch.empty(B, S, C, device=dev) for _ in range(ns)]
        x0s = [torch.empty(B, S, C, device=dev) for _ in range(ns)]
        gvs = [torch.empty(B, S, C, device=dev, dtype=torch.bfloat16) for _ in range(ns)]
        lps = torch.empty(ns, B, device=dev)
        glp = torch.randn(B, device=dev)
        old = torch.randn(ns, B, device=dev) * 0.01 - 1
        adv = torch.randn(B, device=dev)
        rows = torch.zeros(ns, B, 4, device=dev)
        e = B * S * C
        res = {}

        def rec(name, us, bpe, launches_equiv=1):
            res[name] = {"us": round(us, 3), "GBps": round(e * bpe * launches_equiv / us / 1e3, 1), "frac": round(e * bpe * launches_equiv / us / 1e3 / PEAKS, 4)}

        for early in (0, 1, 2):
            def ode():
                for i in range(ns):
                    ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_DETERMINISTIC, out_x_next=outs[i], out_logp=lps[i], want_x0=False, round_like_torch=True, early=early)
            rec(f"ode_early{early}", timed_graph(ode, ns, stream), 10)

            def sde():
                for i in range(ns):
                    ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_NOISE, noise=es[i], out_x_next=outs[i], out_logp=lps[i], want_x0=False, round_like_torch=True, early=early)
            rec(f"sde_early{early}", timed_graph(sde, ns, stream), 12)

            def sde_x0():
                for i in range(ns):
                    ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_NOISE, noise=es[i], out_x_next=outs[i], out_logp=lps[i], want_x0=True, out_x0=x0s[i], round_like_torch=True, early=early)
            rec(f"sde_x0_early{early}", timed_graph(sde_x0, ns, stream), 16)

        # a dependent chain, the way the rollout runs: step i reads what step i-1 wrote (L2-resident x), v independent
        for early in (0, 1):
            def chain():
                for i in range(ns):
                    ops.fused_step(ops.FLOW, vs[i], outs[(i - 1) % ns], k, src=SRC_DETERMINISTIC, out_x_next=outs[i], out_logp=lps[i], want_x0=False, round_like_torch=True, early=early)
            rec(f"ode_chain_early{early}", timed_graph(chain, ns, stream), 10)

        def philox():
            for i in range(ns):
                ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_PHILOX, philox=(1234, 4 * i), out_x_next=outs[i], out_logp=lps[i], want_x0=False, round_like_torch=True)
        rec("sde_philox", timed_graph(philox, ns, stream), 10)

        def fwd1():
            for i in range(ns):
                ops.policy_forward(ops.FLOW, vs[i], xs[i], outs[(i + 1) % ns], k, old[i], adv, 1e-4, 5.0, 0.01, 12.0, stats_rows=rows[i], round_like_torch=True, out_logp=lps[i], accumulate=False)
        rec("fwd_single", timed_graph(fwd1, ns, stream), 10)

        def bwd1():
            for i in range(ns):
                ops.policy_backward(ops.FLOW, vs[i], xs[i], outs[(i + 1) % ns], lps[i], k, old[i], adv, 1e-4, 5.0, 0.01, 12.0, round_like_torch=True, early_loads=True)
        rec("bwd_single_early", timed_graph(bwd1, ns, stream), 12)

        # window: 4 items per launch, rotating over the ns sets (ns // 4 batched launches per graph)
        J = 4
        groups = [[(4 * q + j) % ns for j in range(J)] for q in range(max(1, ns // J))]
        for early in (False, True):
            def fwdm():
                for idx in groups:
                    ops.policy_forward_multi(ops.FLOW, [vs[i] for i in idx], [xs[i] for i in idx], [outs[(i + 1) % ns] for i in idx], ks, [old[i] for i in idx], adv,
                                             1e-4, 5.0, 0.01, 12.0, stats_rows=[rows[i] for i in idx], round_like_torch=True, out_logps=lps[idx[0]:idx[0] + J] if idx[0] + J <= ns else None,
                                             accumulate=False, early_loads=early)
            rec(f"fwd_multi4_early{int(early)}", timed_graph(fwdm, len(groups), stream), 10, J)
            gbuf = [[gvs[i] for i in idx] for idx in groups]

            def bwdm():
                for q, idx in enumerate(groups):
                    ops.policy_backward_multi(ops.FLOW, [vs[i] for i in idx], [xs[i] for i in idx], [outs[(i + 1) % ns] for i in idx], lps[:J], ks, [old[i] for i in idx], adv,
                                              1e-4, 5.0, 0.01, 12.0, round_like_torch=True, early_loads=early, out_grads=gbuf[q])
            rec(f"bwd_multi4_early{int(early)}", timed_graph(bwdm, len(groups), stream), 12, J)

        def pair():
            for q, idx in enumerate(groups):
                nl = ops.policy_forward_multi(ops.FLOW, [vs[i] for i in idx], [xs[i] for i in idx], [outs[(i + 1) % ns] for i in idx], ks, [old[i] for i in idx], adv,
                                              1e-4, 5.0, 0.01, 12.0, stats_rows=[rows[i] for i in idx], round_like_torch=True, out_logps=lps[:J], accumulate=False)
                ops.policy_backward_multi(ops.FLOW, [vs[i] for i in idx], [xs[i] for i in idx], [outs[(i + 1) % ns] for i in idx], nl, ks, [old[i] for i in idx], adv,
                                          1e-4, 5.0, 0.01, 12.0, round_like_torch=True, early_loads=True, out_grads=gbuf[q])
        rec("window_pair_multi4 (fwd+bwd, 22 B/elem x 4)", timed_graph(pair, len(groups), stream), 22, J)
        out[f"B{B}"] = res
        print(f"--- group {B}: E = {e} scalars/launch, peak {PEAKS} GB/s")
        for name, r in res.items():
            print(f"{name:46s} {r['us']:8.2f} us  {r['GBps']:8.1f} GB/s  {r['frac']:.3f}")
        del xs, vs, es, outs, x0s, gvs
        torch.cuda.empty_cache()
    if args.json:
        Path(args.json).write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
