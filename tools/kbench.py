"""Kernel micro-benchmark (GPU box): CUDA-event time per launch over rotating buffer sets (> L2), for
tuning.  Prints one JSON line per variant.  Usage: python tools/kbench.py [--B 12] [--S 4096] [--reps 200]"""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mixgrpo_b200 import _cabi, coefs, ops  # noqa: E402
from mixgrpo_b200._cabi import SRC_DETERMINISTIC, SRC_GIVEN, SRC_NOISE  # noqa: E402


def timeit(fn, nsets, reps, warm=20):
    for i in range(warm):
        fn(i % nsets)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        fn(i % nsets)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=12)
    ap.add_argument("--S", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=300)
    ap.add_argument("--sets", type=int, default=10)
    ap.add_argument("--profile", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, S, ns = a.B, a.S, a.sets
    E = B * S * 64
    sig = torch.linspace(1, 0, 26)
    sig = (3.0 * sig) / (1 + 2.0 * sig)
    lib = _cabi.lib()
    g = torch.Generator(device=dev).manual_seed(0)
    xs = [torch.randn(B, S, 64, device=dev, generator=g) for _ in range(ns)]
    vs = [torch.randn(B, S, 64, device=dev, generator=g).bfloat16() for _ in range(ns)]
    es = [torch.randn(B, S, 64, device=dev, generator=g).bfloat16() for _ in range(ns)]
    outs = [torch.empty(B, S, 64, device=dev) for _ in range(ns)]
    x0s = [torch.empty(B, S, 64, device=dev) for _ in range(ns)]
    gvs = [torch.empty(B, S, 64, device=dev, dtype=torch.bfloat16) for _ in range(ns)]
    lp = torch.empty(B, device=dev)
    glp = torch.randn(B, device=dev)
    ws = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
    k, _ = coefs.flow(sig, 9, 0.7, "ref_cuda", True)
    st = torch.cuda.current_stream().cuda_stream
    n = S * 64

    def flow(i, src, x0=True, flags=1):
        rc = lib.mixgrpo_flow_step(vs[i].data_ptr(), 1, xs[i].data_ptr(), n, es[i].data_ptr() if src == SRC_NOISE else None,
                                   outs[(i + 1) % ns].data_ptr() if src == SRC_GIVEN else None, n,
                                   outs[i].data_ptr() if src != SRC_GIVEN else None, n, x0s[i].data_ptr() if x0 else None, None,
                                   lp.data_ptr(), ws.data_ptr(), ws.numel(), B, n, C.byref(k), src, flags, st, None)
        assert rc == 0, rc

    def bwd(i):
        rc = lib.mixgrpo_logprob_bwd(0, vs[i].data_ptr(), 1, xs[i].data_ptr(), n, outs[i].data_ptr(), n, glp.data_ptr(),
                                     gvs[i].data_ptr(), B, n, C.byref(k), 1, st)
        assert rc == 0, rc

    def copy(i):
        outs[i].copy_(xs[i])

    def report(name, us, bytes_per_elem, **kw):
        gbs = E * bytes_per_elem / us / 1e3
        print(json.dumps({"name": name, "B": B, "S": S, "us": round(us, 3), "GBps": round(gbs, 1), "bytes_per_elem": bytes_per_elem,
                          "frac_6533": round(gbs / 6533.5, 4), **kw}), flush=True)

    if a.profile:
        # short run for ncu: a few launches of the headline kernel and the train-path kernels
        for i in range(6):
            flow(i % ns, SRC_NOISE)
        for i in range(3):
            flow(i % ns, SRC_GIVEN, x0=False)
            bwd(i % ns)
        torch.cuda.synchronize()
        print("profile run done")
        return

    def graph_time(fn):
        nonlocal st
        s = torch.cuda.Stream()
        st_old = st
        with torch.cuda.stream(s):
            st = s.cuda_stream
            fn(0)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                st = torch.cuda.current_stream().cuda_stream
                for i in range(ns):
                    fn(i)
            st = st_old
            us = timeit(lambda i: gr.replay(), 1, max(a.reps // ns, 10), warm=3) / ns
        return us

    report("torch_copy_f32", timeit(copy, ns, a.reps), 8)
    kw = {}
    report("flow_sde_rollout_x0", timeit(lambda i: flow(i, SRC_NOISE), ns, a.reps), 16, **kw)
    report("flow_sde_rollout_x0_graph", graph_time(lambda i: flow(i, SRC_NOISE)), 16, **kw)
    report("flow_sde_rollout_nox0_graph", graph_time(lambda i: flow(i, SRC_NOISE, x0=False)), 12, **kw)
    report("flow_ode_rollout_x0_graph", graph_time(lambda i: flow(i, SRC_DETERMINISTIC)), 14, **kw)
    report("flow_ode_rollout_nox0_graph", graph_time(lambda i: flow(i, SRC_DETERMINISTIC, x0=False)), 10, **kw)
    report("flow_train_fwd", timeit(lambda i: flow(i, SRC_GIVEN, x0=False), ns, a.reps), 10, **kw)
    report("flow_train_fwd_graph", graph_time(lambda i: flow(i, SRC_GIVEN, x0=False)), 10, **kw)
    report("flow_sde_rollout_x0_noround_graph", graph_time(lambda i: flow(i, SRC_NOISE, flags=0)), 16, **kw)
    report("logprob_bwd", timeit(bwd, ns, a.reps), 12)
    report("logprob_bwd_graph", graph_time(bwd), 12)


if __name__ == "__main__":
    main()
