#!/usr/bin/env python
"""Times the log-prob tail variants of tools/exp_tail.cu on a B200 the way bench.py times its roofline kernel: 25 launches
cycling through 10 buffer sets (every input cold) + 1 collecting launch inside a CUDA graph, programmatic dependent launch,
CUDA events around 20 replays.  Measurement only.

  python tools/exp_tail.py --build          # here (no GPU): nvcc -> build/libexp_tail.so (travels to the GPU box)
  python tools/exp_tail.py [--json out]     # on the GPU box"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
SO = ROOT / "build" / "libexp_tail.so"


def build():
    SO.parent.mkdir(exist_ok=True)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC",
                           str(ROOT / "tools" / "exp_tail.cu"), "-o", str(SO)])
    print(SO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--json", default=None)
    ap.add_argument("--groups", type=int, nargs="+", default=[12])
    ap.add_argument("--rounds", type=int, default=3)
    args = ap.parse_args()
    if args.build:
        return build()
    import torch
    from mixgrpo_b200 import coefs, ops
    from mixgrpo_b200._cabi import SRC_DETERMINISTIC
    sys.path.insert(0, str(ROOT))
    from bench import _time_graph, load_peaks
    lib = C.CDLL(str(SO))
    lib.exp_variant_name.restype = C.c_char_p
    lib.exp_ode.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    lib.exp_collect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    dev = torch.device("cuda:0")
    peak = load_peaks()[0]
    S, CH, NL = 4096, 64, 25
    sig = torch.linspace(1, 0, 26)
    sig = (3.0 * sig) / (1 + 2.0 * sig)
    k, _ = coefs.flow(sig, 9, 0.7, "ref_cuda", True)
    out = {}
    for B in args.groups:
        ns = 10 if B <= 12 else (8 if B <= 24 else 6)
        g = torch.Generator(device=dev).manual_seed(7)
        xs = [torch.randn(B, S, CH, device=dev, generator=g) for _ in range(ns)]
        vs = [torch.randn(B, S, CH, device=dev, generator=g).bfloat16() for _ in range(ns)]
        outs = [torch.empty(B, S, CH, device=dev) for _ in range(ns)]
        rec = torch.zeros(NL, B * 4, dtype=torch.int64, device=dev)
        lib.exp_sub_words.restype = C.c_int64
        lib.exp_sub_words.argtypes = [C.c_int64]
        sub = torch.zeros(NL, int(lib.exp_sub_words(B)), dtype=torch.int64, device=dev)
        means = torch.zeros(NL, B, dtype=torch.float64, device=dev)
        s = torch.cuda.Stream(device=dev)
        e = B * S * CH
        names, variants = {}, []
        v = 0
        while lib.exp_variant_name(v):
            names[v] = lib.exp_variant_name(v).decode()
            variants.append(v)
            v += 1

        def run(variant):
            def fn():
                st = torch.cuda.current_stream().cuda_stream
                for j in range(NL):
                    i = j % ns
                    rc = lib.exp_ode(variant, vs[i].data_ptr(), xs[i].data_ptr(), outs[i].data_ptr(), rec[j].data_ptr(), sub[j].data_ptr(), B, S * CH, C.addressof(k), st)
                    assert rc == 0, rc
                rc = lib.exp_collect(sub.data_ptr(), rec.data_ptr(), means.data_ptr(), NL * B, st)
                assert rc == 0, rc
            return fn

        # what ships, through the product library, for the cross-reference
        acc = ops.DeferredLogProbs(dev, NL, B, S * CH)
        lp25 = torch.empty(NL, B, device=dev)

        def product():
            for j in range(NL):
                i = j % ns
                ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_DETERMINISTIC, out_x_next=outs[i], want_x0=False, defer=acc.slot(j, k), round_like_torch=True, early=1)
            acc.finalize(lp25)

        res = {"product": []}
        ref_mean = None
        for r in range(args.rounds):
            res["product"].append(round(_time_graph(product, NL, s), 3))
            for var in variants:
                us = _time_graph(run(var), NL, s)
                res.setdefault(names[var], []).append(round(us, 3))
                torch.cuda.synchronize()
                if "no log-prob" not in names[var] and "PRODUCT" not in names[var]:
                    m = means[0].clone()
                    if ref_mean is None:
                        ref_mean = m
                    err = float((m - ref_mean).abs().max())        # absolute, on a log-prob of order 1
                    assert err < 2e-7, (names[var], err, m, ref_mean)
        torch.cuda.synchronize()
        want = (-lp25[0].double() - float(k.log_scale) - float(k.log_norm))
        print(f"--- group {B}: {e * 10 / 1e6:.1f} MB per launch, peak {peak} GB/s; mean(d^2/2s^2) variants vs product: "
              f"{float((ref_mean - want).abs().max()):.2e} abs")
        for name, us in res.items():
            best = min(us)
            print(f"{name:75s} {' '.join(f'{u:6.2f}' for u in us)} us   best {e * 10 / best / 1e3:7.1f} GB/s  {e * 10 / best / 1e3 / peak:.3f}")
        out[f"B{B}"] = res
        del xs, vs, outs
        torch.cuda.empty_cache()
    if args.json:
        Path(args.json).write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
