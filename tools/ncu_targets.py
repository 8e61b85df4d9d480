#!/usr/bin/env python
"""A short run of exactly the kernel instantiations the timed bench step launches, for `ncu --set full` (GPU box).

    python tools/ncu_targets.py                                  # must exit 0 on its own first
    ncu --set full --clock-control none --import-source on -o gpurun_out/r02_targets python tools/ncu_targets.py
    python tools/ncu_full_summary.py gpurun_out/r02_targets.ncu-rep --kernel-regex '<regex>' --json profiles/<name>.json ...

Launch order (3 launches each, rotating buffer sets so no input is hot in L2 beyond what ncu's cache control leaves); the
rollout's step launches are issued as the rollout issues them — log-prob sums accumulated (MIXGRPO_FLAG_DEFER_LOGP), ONE finalize:
  0. first step: bf16 latent in, all_latents[:, 0] and [:, 1] out   mg::step_kernel<flow,bf16,SRC_NOISE,OUT=0,EXT=2,HALF>   (1 launch per step)
  1. Euler-ODE sampler step + log-prob        mg::step_kernel<flow,bf16,SRC_DETERMINISTIC,OUT=0,HALF>   (21 of a step's 29 launches; 128-thread CTAs)
  2. SDE sampler step + log-prob, no x0       mg::step_kernel<flow,bf16,SRC_NOISE,OUT=0,HALF>           (3 launches)
  2b. mg::logp_finalize_kernel                                                                   (1 launch)
  3. window forward, 4 items in one launch    mg::policy_fwd_multi_kernel<flow,bf16>               (1 launch)
  4. window backward, 4 items in one launch   mg::policy_bwd_multi_kernel<flow,bf16>               (1 launch)
  5. stored-transition log-prob (SRC_GIVEN)   mg::step_kernel<flow,bf16,SRC_GIVEN,OUT=0>           (drop-in / per-step path)
  6. log-prob backward                        mg::logprob_bwd_kernel<flow,bf16>                    (drop-in / per-step path)
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mixgrpo_b200 import coefs, ops  # noqa: E402
from mixgrpo_b200._cabi import SRC_DETERMINISTIC, SRC_GIVEN, SRC_NOISE  # noqa: E402

B, S, C, NS, J = 12, 4096, 64, 8, 4


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    xs = [torch.randn(B, S, C, device=dev, generator=g) for _ in range(NS)]
    vs = [torch.randn(B, S, C, device=dev, generator=g).bfloat16() for _ in range(NS)]
    es = [torch.randn(B, S, C, device=dev, generator=g).bfloat16() for _ in range(NS)]
    outs = [torch.empty(B, S, C, device=dev) for _ in range(NS)]
    gvs = [torch.empty(B, S, C, device=dev, dtype=torch.bfloat16) for _ in range(NS)]
    lps = torch.empty(NS, B, device=dev)
    old = torch.randn(NS, B, device=dev) * 0.01 - 1
    adv = torch.randn(B, device=dev)
    glp = torch.randn(B, device=dev)
    rows = torch.zeros(NS, B, 4, device=dev)
    sig = torch.linspace(1, 0, 26)
    sig = (3.0 * sig) / (1 + 2.0 * sig)
    k, _ = coefs.flow(sig, 9, 0.7, "ref_cuda", True)
    ks = [coefs.flow(sig, t, 0.7, "ref_cuda", True)[0] for t in range(J)]
    acc = ops.DeferredLogProbs(dev, 7, B, S * C)
    z0 = torch.randn(B, S, C, device=dev, generator=g).bfloat16()
    seed_slot = torch.empty(B, S, C, device=dev)
    ops.fused_step(ops.FLOW, vs[7], z0, k, src=SRC_NOISE, noise=es[7], out_x_next=outs[7], want_x0=False, round_like_torch=True, early=1, defer=acc.slot(6, k),
                   seed_out=seed_slot)
    for i in range(3):
        ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_DETERMINISTIC, out_x_next=outs[i], want_x0=False, round_like_torch=True, early=1, defer=acc.slot(i, k))
    for i in range(3, 6):
        ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_NOISE, noise=es[i], out_x_next=outs[i], want_x0=False, round_like_torch=True, early=1, defer=acc.slot(i, k))
    acc.finalize(torch.empty(7, B, device=dev))
    for q in range(2):
        idx = [q * J + j for j in range(J)]
        ops.policy_forward_multi(ops.FLOW, [vs[i] for i in idx], [xs[i] for i in idx], [outs[(i + 1) % NS] for i in idx], ks, [old[i] for i in idx], adv,
                                 1e-4, 5.0, 0.01, 12.0, stats_rows=[rows[i] for i in idx], round_like_torch=True, out_logps=lps[q * J:(q + 1) * J], accumulate=False,
                                 early_loads=True)
        ops.policy_backward_multi(ops.FLOW, [vs[i] for i in idx], [xs[i] for i in idx], [outs[(i + 1) % NS] for i in idx], lps[q * J:(q + 1) * J], ks,
                                  [old[i] for i in idx], adv, 1e-4, 5.0, 0.01, 12.0, round_like_torch=True, early_loads=True, out_grads=[gvs[i] for i in idx])
    for i in range(3):
        ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_GIVEN, x_next=outs[(i + 1) % NS], out_logp=lps[i], want_x0=False, round_like_torch=True)
        ops.logprob_backward(ops.FLOW, vs[i], xs[i], outs[(i + 1) % NS], glp, k, True, out=gvs[i])
    torch.cuda.synchronize()
    print("ncu_targets: done,", ops.launch_count, "launches")


if __name__ == "__main__":
    main()
