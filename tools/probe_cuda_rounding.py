"""One-off GPU probe (SURVEY.md §7 stage 0): how does the REFERENCE round on CUDA tensors?

Runs the unmodified reference operators (loaded by oracle/ref_loader.py from a git-ignored
``baseline/_ref`` drop of the reference tree — it does not exist on the GPU box otherwise) on CUDA
tensors with bf16 model output and explicit noise, and compares them bit-for-bit with
  (a) the oracle restatement executed on the same CUDA tensors,
  (b) the CUDA kernels in rounding="ref_cuda" and rounding="ref_cpu".
Writes gpurun_out/probe_cuda_rounding.json and gpurun_out/golden_cuda_reference.npz (reference-on-CUDA
outputs for small seeded inputs; committed under tests/golden/ as the CUDA golden vectors).
Not a test; nothing in tests/, smoke() or bench.py depends on the reference tree.
"""
import json
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_loader, sampling_oracle as O  # noqa: E402


def main():
    su_ref = ref_loader.load()
    assert su_ref is not None, "reference tree not found (baseline/_ref)"
    from mixgrpo_b200 import sampling_utils as su
    dev = torch.device("cuda:0")
    sig_cpu = O.sd3_time_shift(3.0, torch.linspace(1, 0, 26))
    sig = sig_cpu.to(dev)
    report, golden = {}, {}
    g = torch.Generator().manual_seed(0)
    B, S = 2, 2      # small on purpose: these become committed fixtures

    def mk(dtype):
        x = torch.randn(B, S, 64, generator=g)
        v = torch.randn(B, S, 64, generator=g).to(dtype)
        e = torch.randn(B, S, 64, generator=g).to(dtype)
        xn = torch.randn(B, S, 64, generator=g)
        return x, v, e, xn

    def eq(a, b):
        a, b = a.detach().float().cpu(), b.detach().float().cpu()
        if torch.equal(torch.nan_to_num(a), torch.nan_to_num(b)):
            return 0.0
        return float(((a - b).norm() / b.norm().clamp_min(1e-30)))

    for dtype, dn in ((torch.bfloat16, "bf16"), (torch.float32, "f32")):
        for idx in (0, 1, 3, 12, 23, 24):
            for det in (False, True):
                x, v, e, xn = mk(dtype)
                key = f"flow_{dn}_i{idx}_det{int(det)}"
                ref_loader.NOISE_QUEUE[:] = [e.to(dev)]
                r = su_ref.flow_grpo_step(v.to(dev), x.to(dev), 0.7, sig, idx, None, determistic=det)
                o = O.flow_step(v.to(dev), x.to(dev), 0.7, sig, idx, None, e.to(dev), det)
                k_cuda = su.flow_grpo_step(v.to(dev), x.to(dev), 0.7, sig, idx, None, determistic=det, noise=e.to(dev), rounding="ref_cuda")
                k_cpu = su.flow_grpo_step(v.to(dev), x.to(dev), 0.7, sig, idx, None, determistic=det, noise=e.to(dev), rounding="ref_cpu")
                k_f32 = su.flow_grpo_step(v.to(dev), x.to(dev), 0.7, sig, idx, None, determistic=det, noise=e.to(dev), rounding="fp32")
                rc = su_ref.flow_grpo_step(v, x, 0.7, sig_cpu, idx, None, determistic=det) if ref_loader.NOISE_QUEUE.append(e) is None else None
                report[key] = {
                    "oracle_on_cuda_vs_ref": [eq(a, b) for a, b in zip(o[:4], r[:4])],
                    "kernel_ref_cuda_vs_ref": [eq(a, b) for a, b in zip(k_cuda[:4], r[:4])],
                    "kernel_ref_cpu_vs_ref": [eq(a, b) for a, b in zip(k_cpu[:4], r[:4])],
                    "kernel_fp32_vs_ref": [eq(a, b) for a, b in zip(k_f32[:4], r[:4])],
                    "ref_cpu_vs_ref_cuda": [eq(a, b) for a, b in zip(rc[:4], r[:4])],
                }
                for nm, t in zip(("x", "v", "eps", "prev", "x0", "logp", "mean"), (x, v.float(), e.float(), *r[:4])):
                    golden[f"{key}/{nm}"] = t.detach().float().cpu().numpy()
                # train path + autograd on CUDA
                if not det:
                    vg = v.to(dev).requires_grad_(True)
                    rr = su_ref.flow_grpo_step(vg, x.to(dev), 0.7, sig, idx, xn.to(dev))
                    rr[2].sum().backward()
                    vk = v.to(dev).requires_grad_(True)
                    kk = su.flow_grpo_step(vk, x.to(dev), 0.7, sig, idx, xn.to(dev), rounding="ref_cuda")
                    kk[2].sum().backward()
                    report[key]["train_logp_kernel_vs_ref"] = eq(kk[2], rr[2])
                    report[key]["train_grad_kernel_vs_ref"] = eq(vk.grad, vg.grad)
                    report[key]["train_grad_dtype"] = str(vg.grad.dtype)
                    golden[f"{key}/xn_train"] = xn.numpy()
                    golden[f"{key}/train_logp"] = rr[2].detach().cpu().numpy()
                    golden[f"{key}/train_grad"] = vg.grad.float().cpu().numpy()
        # dance
        for idx in (0, 3, 12, 24):
            for sde in (True, False):
                x, v, e, xn = mk(dtype)
                key = f"dance_{dn}_i{idx}_sde{int(sde)}"
                r = su_ref.dance_grpo_step(v.to(dev), x.to(dev), 0.7, sig, idx, xn.to(dev), True, sde)
                o = O.dance_step(v.to(dev), x.to(dev), 0.7, sig, idx, xn.to(dev), None, True, sde)
                kc = su.dance_grpo_step(v.to(dev), x.to(dev), 0.7, sig, idx, xn.to(dev), True, sde, rounding="ref_cuda")
                report[key] = {"oracle_on_cuda_vs_ref": [eq(a, b) for a, b in zip(o, r)],
                               "kernel_ref_cuda_vs_ref": [eq(a, b) for a, b in zip(kc, r)]}
                for nm, t in zip(("x", "v", "xn", "x0", "logp"), (x, v.float(), xn, r[1], r[2])):
                    golden[f"{key}/{nm}"] = t.detach().float().cpu().numpy()
        # dpm (Flash configuration: dpmsolver++ order 2 midpoint/heun), ODE and SDE
        for stype in ("midpoint", "heun"):
            for sde in (False, True):
                args = types.SimpleNamespace(dpm_algorithm_type="dpmsolver++", dpm_solver_type=stype, dpm_solver_order=2)
                rs, ks = su_ref.DPMState(order=2), su.DPMState(order=2)
                worst = [0.0, 0.0, 0.0]
                for idx in range(25):
                    x, v, e, _ = mk(dtype)
                    e = e.float()
                    r = su_ref.dpm_step(args, v.to(dev), x.to(dev), idx, sig[:-1], sig, dpm_state=rs, variance_noise=e.to(dev), sde_solver=sde)
                    kk = su.dpm_step(args, v.to(dev), x.to(dev), idx, sig[:-1], sig, dpm_state=ks, variance_noise=e.to(dev), sde_solver=sde, rounding="ref_cuda")
                    ks.model_outputs = list(rs.model_outputs)     # same history on both sides
                    errs = [eq(a, b) for a, b in zip(kk, r)]
                    if not sde:
                        errs[2] = 0.0
                    worst = [max(a, b) for a, b in zip(worst, errs)]
                    if idx in (1, 12, 23):
                        key = f"dpm_{dn}_{stype}_sde{int(sde)}_i{idx}"
                        for nm, t in zip(("x", "v", "eps", "m1", "prev", "x0", "logp"),
                                         (x, v.float(), e, rs.model_outputs[0], *r)):
                            golden[f"{key}/{nm}"] = t.detach().float().cpu().numpy()
                report[f"dpm_{dn}_{stype}_sde{int(sde)}"] = {"kernel_ref_cuda_vs_ref_worst": worst}
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "probe_cuda_rounding.json").write_text(json.dumps(report, indent=1))
    np.savez_compressed(out / "golden_cuda_reference.npz", **golden)
    bad = {k: v for k, v in report.items() if any(x != 0.0 for x in v.get("kernel_ref_cuda_vs_ref", [0.0])[:2] + v.get("kernel_ref_cuda_vs_ref", [0, 0, 0, 0])[3:4])}
    print(json.dumps({"n_cases": len(report), "n_not_bit_exact_ref_cuda": len(bad), "examples": dict(list(bad.items())[:3])}, indent=1))


if __name__ == "__main__":
    main()
