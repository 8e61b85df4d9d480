"""Policy-update micro-benchmark (GPU box): the two-launch pair (mixgrpo_policy_fwd + mixgrpo_policy_bwd, 22 B/elem) against
the single-pass kernel (mixgrpo_policy_step, 12 B/elem) on the same data, as 10-update CUDA graphs over rotating buffer sets
(> L2), CUDA events.  Sweeps the single-pass kernel's CTAs/SM and cooperative-launch knobs and checks that four
concurrent single-pass launches on four streams complete (no co-residency deadlock).  One JSON line per variant.
Usage: python tools/pbench.py [--B 12] [--S 4096]"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mixgrpo_b200 import _cabi, coefs, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=12)
    ap.add_argument("--S", type=int, default=4096)
    ap.add_argument("--sets", type=int, default=10)
    ap.add_argument("--reps", type=int, default=30)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _cabi.lib()
    lib.mixgrpo_set_tuning(5, 3000)
    B, S, ns = a.B, a.S, a.sets
    E = B * S * 64
    g = torch.Generator(device=dev).manual_seed(0)
    xs = [torch.randn(B, S, 64, device=dev, generator=g) for _ in range(ns)]
    vs = [torch.randn(B, S, 64, device=dev, generator=g).bfloat16() for _ in range(ns)]
    xn = [x + 0.3 * torch.randn(B, S, 64, device=dev, generator=g) for x in xs]
    old = torch.full((B,), -1.0, device=dev)
    adv = torch.randn(B, device=dev, generator=g)
    rows = torch.zeros(B, 4, device=dev)
    sig = torch.linspace(1, 0, 26)
    sig = (3.0 * sig) / (1 + 2.0 * sig)
    k, _ = coefs.flow(sig, 9, 0.7, "ref_cuda", True)
    args = (1e-4, 5.0, 0.01, 12.0)

    def pair(i):
        lp = ops.policy_forward(ops.FLOW, vs[i], xs[i], xn[i], k, old, adv, *args, stats_rows=rows, round_like_torch=True, accumulate=False)
        return ops.policy_backward(ops.FLOW, vs[i], xs[i], xn[i], lp, k, old, adv, *args, round_like_torch=True, early_loads=True)

    def single(i):
        r = ops.policy_step(ops.FLOW, vs[i], xs[i], xn[i], k, old, adv, *args, stats_rows=rows, round_like_torch=True, accumulate=False)
        assert r is not None, "single-pass kernel refused the shape"
        return r[1]

    def graph_time(fn):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            out = fn(0)
            s.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                for i in range(ns):
                    out = fn(i)
            for _ in range(3):
                gr.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(a.reps):
                gr.replay()
            e1.record(s)
            e1.synchronize()
        ok = bool(torch.isfinite(out.float()).all().item())
        return e0.elapsed_time(e1) * 1e3 / (a.reps * ns), ok

    def report(name, us, ok, **kw):
        print(json.dumps({"name": name, "B": B, "S": S, "us_per_update": round(us, 3), "finite": ok,
                          "GBps_at_12B": round(E * 12 / us / 1e3, 1), "GBps_at_22B": round(E * 22 / us / 1e3, 1), **kw}), flush=True)

    us, ok = graph_time(pair)
    report("policy_fwd + policy_bwd (2 launches, 22 B/elem)", us, ok)
    for coop in (1, 0):
        lib.mixgrpo_set_tuning(4, coop)
        for r in (6, 5, 4, 3, 2):
            lib.mixgrpo_set_tuning(3, r)
            try:
                us, ok = graph_time(single)
                report("policy_step (1 launch, 12 B/elem)", us, ok, ctas_per_sm=r, cooperative=coop)
            except Exception as e:  # noqa: BLE001
                print(json.dumps({"name": "policy_step", "ctas_per_sm": r, "cooperative": coop, "error": f"{type(e).__name__}: {e}"[:300]}), flush=True)
                torch.cuda.synchronize()
    lib.mixgrpo_set_tuning(4, 1)
    lib.mixgrpo_set_tuning(3, 6)
    # four concurrent single-pass launches (bench.py runs the window's four updates on parallel graph branches)
    for coop in (1, 0):
        lib.mixgrpo_set_tuning(4, coop)
        streams = [torch.cuda.Stream() for _ in range(4)]
        outs = []
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for rep in range(20):
            for j, s in enumerate(streams):
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    outs.append(single((4 * rep + j) % ns))
            for s in streams:
                torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        ok = all(bool(torch.isfinite(o.float()).all().item()) for o in outs[-8:])
        print(json.dumps({"name": "4 concurrent policy_step launches on 4 streams x 20", "cooperative": coop, "finite": ok,
                          "us_per_update": round(e0.elapsed_time(e1) * 1e3 / 80, 2)}), flush=True)
        outs.clear()
    lib.mixgrpo_set_tuning(4, 1)


if __name__ == "__main__":
    main()
