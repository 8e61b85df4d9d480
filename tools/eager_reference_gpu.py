"""GPU-side "before" number: the reference's own PyTorch-eager operator path on the B200 — its unmodified
`flow_grpo_step` when a reference tree is reachable (oracle/ref_loader.py: /root/reference or the git-ignored baseline/_ref drop),
else the oracle's restatement of the same op sequence — with CUDA tensors, CUDA-event timed, next to this package's fused
kernels on the same inputs.  Measurement tool only (it imports oracle/, so it is not part of the product or of bench.py).
Usage: python tools/eager_reference_gpu.py"""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mixgrpo_b200 import ops, sampling_utils as su  # noqa: E402
from oracle import ref_loader, sampling_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
ETA = 0.7
sig = O.sd3_time_shift(3.0, torch.linspace(1, 0, 26)).to(dev)
ref = ref_loader.load()


def timed(fn, n=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n, (time.perf_counter() - t0) * 1e6 / n


for B in (1, 12):
    g = torch.Generator(device=dev).manual_seed(B)
    x = torch.randn(B, 4096, 64, device=dev, generator=g)
    v = torch.randn(B, 4096, 64, device=dev, generator=g).bfloat16()
    eps = torch.randn(B, 4096, 64, device=dev, generator=g).bfloat16()
    xn = O.flow_step(v, x, ETA, sig, 9, None, eps, False)[0]
    rows = {}

    def ref_rollout():
        if ref is not None:
            ref_loader.NOISE_QUEUE.append(eps)
            return ref.flow_grpo_step(v, x, ETA, sig, 9, None)
        return O.flow_step(v, x, ETA, sig, 9, None, eps, False)

    def ref_train():
        vg = v.detach().requires_grad_(True)
        lp = (ref.flow_grpo_step(vg, x, ETA, sig, 9, xn) if ref is not None else O.flow_step(vg, x, ETA, sig, 9, xn))[2]
        lp.sum().backward()
        return vg.grad

    def our_train():
        vg = v.detach().requires_grad_(True)
        lp = su.flow_grpo_step(vg, x, ETA, sig, 9, xn, return_mean=False)[2]
        lp.sum().backward()
        return vg.grad

    def our_train_fused():
        # the policy update as the trainer issues it (rollout.policy_update): forward + loss and backward, no autograd at all
        k = our_train_fused.k
        nl = ops.policy_forward(ops.FLOW, v, x, xn, k, old, adv, 1e-4, 5.0, 0.01, 12.0, round_like_torch=True)
        return ops.policy_backward(ops.FLOW, v, x, xn, nl, k, old, adv, 1e-4, 5.0, 0.01, 12.0, round_like_torch=True, early_loads=True)

    from mixgrpo_b200 import coefs
    our_train_fused.k = coefs.flow(sig, 9, ETA, "ref_cuda", True)[0]
    old = torch.randn(B, device=dev) * 0.01 - 1
    adv = torch.randn(B, device=dev)
    def autograd_floor():
        # what torch itself charges for the same call pattern with a trivial built-in op in place of ours
        vg = v.detach().requires_grad_(True)
        lp = vg.mean(dim=(1, 2))
        lp.sum().backward()
        return vg.grad

    rows["reference eager: rollout SDE step + log-prob"] = timed(ref_rollout)
    rows["reference eager: log-prob forward + autograd backward"] = timed(ref_train)
    rows["torch autograd floor: v.detach().requires_grad_(); v.mean((1,2)).sum().backward() (no mixgrpo code)"] = timed(autograd_floor)
    compiled = ops.binding()
    for loader in (["compiled binding"] if compiled is not None else []) + ["ctypes loader"]:
        ops._binding_mod = compiled if loader == "compiled binding" else None
        rows[f"mixgrpo_b200 [{loader}]: rollout SDE step + log-prob (drop-in flow_grpo_step, 1 launch)"] = timed(lambda: su.flow_grpo_step(v, x, ETA, sig, 9, None, noise=eps, return_mean=False))
        rows[f"mixgrpo_b200 [{loader}]: drop-in flow_grpo_step, full 5-tuple (mean written too)"] = timed(lambda: su.flow_grpo_step(v, x, ETA, sig, 9, None, noise=eps))
        rows[f"mixgrpo_b200 [{loader}]: log-prob forward + closed-form backward through autograd (lp.sum().backward(), 2 launches + torch's sum/backward)"] = timed(our_train)
        rows[f"mixgrpo_b200 [{loader}]: fused policy forward + backward (ops.policy_forward / policy_backward, 2 launches, no autograd)"] = timed(our_train_fused)
    ops._binding_mod = compiled
    for k, (dev_us, wall_us) in rows.items():
        print(json.dumps({"B": B, "what": k, "us_per_call_cuda_events": round(dev_us, 1), "us_per_call_wall": round(wall_us, 1),
                          "reference_impl": "unmodified reference file" if ref is not None else "oracle restatement"}), flush=True)
