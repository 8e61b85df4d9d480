#!/usr/bin/env python
"""bench.py — MixGRPO rollout + policy-update hot path on B200 (driver contract: one JSON line).

Headline workload (BASELINE.json configs[1]): FLUX.1-dev-shape 1024^2 packed latents (12, 4096, 64), group size 12,
25 sampling steps, SDE window 4, bf16 model output and noise, fp32 latents.  One bench "step" = one GRPO iteration's
hot path for the prompt group(s) this rank owns, on synthetic random-init tensors:

  rollout        25 fused sampler steps (21 Euler-ODE + 4 SDE with log-prob), each reading a distinct pre-generated
                 model output v_i and writing all_latents[:, i+1] in place                      (1 + 25 launches)
  exchange +     the [3 models x 12] rewards gathered across ranks AND the group-relative advantages in ONE peer-memory
  advantages     kernel (N > 1; N = 1: the advantage kernel alone)                              (1 launch)
  policy update  the window's 4 (batch, step) updates as TWO launches: log-prob + clipped-ratio loss forward for all
                 four, loss-grad + log-prob backward for all four -> grad wrt each model output (2 launches)
  logging        ONE all-reduce of the loss/policy/kl/clip_frac sums (N > 1 only: one peer-memory kernel)

metric   = sampler-step latent GB/s = algorithmic bytes of all sampler / log-prob kernels in the step (SURVEY.md §8d
           per-element figures) / step time, summed over ranks ("weak" scaling: each rank owns its own prompt groups;
           no data-path collective).  rollout_steps_per_s is reported beside it.
roofline = the kernel that DOMINATES the timed step: the Euler-ODE instantiation of mg::step_kernel (21 of 29 launches),
           timed alone by CUDA events over graph replays on rotating buffer sets larger than L2, exactly as the step
           launches it; `roofline.step_weighted` is the same figure for the whole set of streaming launches of a step.
configs  = BASELINE configs[3] (MixGRPO-Flash: DPM-Solver++ order-2 ODE tail on the compressed schedule) and configs[4]
           (group 24 at 1024^2 bf16 and at 512^2 fp32-vs-bf16) as whole-step lines at the same N.
e2e      = the same step through the public Python API with HOST (pinned) model outputs and rewards, H2D + D2H copies
           inside the timed region (the SDE noise is drawn on the device, as the reference does), next to the measured
           pinned-copy ceiling of the box at the same N.
--impl reference: the reference's own CPU implementation on the host cores — the unmodified sampling_utils.py executed
           from baseline/_ref when the drop is there (oracle/ref_loader.py), else the oracle restatement pinned to it.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

C = 64                           # channels of a packed FLUX latent token
N_STEPS, WINDOW, N_MODELS = 25, 4, 3
ETA, SHIFT = 0.7, 3.0
CLIP, ADV_CLIP, KL, GA = 1e-4, 5.0, 0.01, 3
WEIGHTS = (1.0, 0.5, 2.0)

# name -> (group size B, packed tokens S, model-output dtype, MixGRPO-Flash?)      BASELINE.json configs[...]
SCENARIOS = {
    "mixgrpo": (12, 4096, torch.bfloat16, False),               # configs[1] / configs[2]: the headline
    "flash": (12, 4096, torch.bfloat16, True),                  # configs[3]: DPM-Solver++ order 2 midpoint, post, ratio 0.4
    "large_b24_1024sq": (24, 4096, torch.bfloat16, False),      # configs[4]
    "large_b24_512sq_bf16": (24, 1024, torch.bfloat16, False),  # configs[4]
    "large_b24_512sq_f32": (24, 1024, torch.float32, False),    # configs[4], fp32 model output and noise
    "mixgrpo_ode_logp_off": (12, 4096, torch.bfloat16, False),  # configs[1] with SamplerConfig.ode_log_probs=False (dead columns skipped)
}
WORKLOAD_TEXT = {
    "mixgrpo": "FLUX.1-dev-shape 1024^2 packed latents (12,4096,64), group 12, 25 steps, SDE window 4 (BASELINE configs[1]); one prompt group per GPU",
    "flash": "MixGRPO-Flash (BASELINE configs[3]): (12,4096,64), SDE window 4 then DPM-Solver++ order-2 midpoint ODE on the compressed schedule (ratio 0.4 -> 11 steps)",
    "large_b24_1024sq": "BASELINE configs[4]: group 24 at 1024^2 (24,4096,64), bf16 model output, 25 steps, SDE window 4",
    "large_b24_512sq_bf16": "BASELINE configs[4]: group 24 at 512^2 (24,1024,64), bf16 model output, 25 steps, SDE window 4",
    "large_b24_512sq_f32": "BASELINE configs[4]: group 24 at 512^2 (24,1024,64), fp32 model output and noise, 25 steps, SDE window 4",
    "mixgrpo_ode_logp_off": "configs[1] with SamplerConfig.ode_log_probs=False: the 21 deterministic steps skip the log-prob reduction (columns the reference computes but train_one_step never reads, TR:536-553); same bytes, trajectory and window log-probs bit-identical",
}


def sampler_config(flash: bool):
    from mixgrpo_b200 import rollout as R
    kw = dict(sampling_steps=N_STEPS, eta=ETA, shift=SHIFT, pdl_early_v=True)   # model outputs are precomputed: not written by the previous launch
    if flash:   # scripts/finetune/finetune_flux_grpo_MixGRPO_Flash.sh:75-79
        kw.update(dpm_algorithm_type="dpmsolver++", dpm_apply_strategy="post", dpm_post_compress_ratio=0.4, dpm_solver_order=2,
                  dpm_solver_type="midpoint")
    return R.SamplerConfig(**kw)


def step_plan(flash: bool, window):
    """The launches of one step as (kind, count) in SURVEY §8d's vocabulary — mirrors rollout.rollout's choices."""
    det = [i not in window for i in range(N_STEPS)]
    if not flash:
        return [("ode", sum(det)), ("sde", len(window)), ("train_fwd", len(window)), ("bwd", len(window))], N_STEPS
    last = max(window)
    n_post = int(max((N_STEPS - last) * 0.4, 1))                                       # SU:44 on a 26-entry schedule
    n_steps = last + n_post                                                             # schedule = sigmas[:last+1] + n_post tail entries
    plan = []
    n_ode_pre = sum(1 for i in range(last + 1) if det[i])
    if n_ode_pre:
        plan.append(("ode_x0", n_ode_pre))                                              # the DPM history needs every x0 (SU:116-117)
    plan.append(("sde_x0", len(window)))
    plan.append(("dpm2_ode_x0", n_steps - (last + 1) - 1))                              # order 2 from the first tail step (history fed by the window)
    plan.append(("dpm1_ode_x0", 1))                                                     # lower_order_final (SU:308)
    plan += [("train_fwd", len(window)), ("bwd", len(window))]
    return plan, n_steps


def bytes_per_elem(kind: str, f32: bool) -> int:
    """SURVEY §8d / BASELINE.md §2: algorithmic bytes per latent scalar of one launch."""
    vb = 4 if f32 else 2                    # model output; flow noise has the same dtype (SU:193)
    return {"ode": vb + 4 + 4, "ode_x0": vb + 4 + 8, "sde": vb + 4 + vb + 4, "sde_x0": vb + 4 + vb + 8, "train_fwd": vb + 8, "bwd": vb + 8 + vb,
            "dpm2_ode_x0": vb + 4 + 4 + 8, "dpm1_ode_x0": vb + 4 + 8, "sde_x0_philox": vb + 4 + 8}[kind]


def algorithmic_bytes_per_step(name: str = "mixgrpo") -> int:
    B, S, dt, flash = SCENARIOS[name]
    plan, _ = step_plan(flash, list(range(WINDOW)))
    return B * S * C * sum(bytes_per_elem(k, dt == torch.float32) * c for k, c in plan)


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback"


def load_ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel from the committed ncu --set full capture."""
    for nm in ("r02_ode_kernel_ncu.json",):
        try:
            d = json.loads((ROOT / "profiles" / nm).read_text())
            l2w = [p["lts__t_sectors_srcunit_tex_op_write.sum"][0] * 32 for p in d.get("per_launch", []) if "lts__t_sectors_srcunit_tex_op_write.sum" in p]
            note = (f"{d.get('note', '')}; per launch: DRAM read {d['dram_bytes_read']} B (= the algorithmic reads, 18.87 MB), DRAM write {d['dram_bytes_write']} B "
                    f"(the 12.58 MB of output is still in the 126 MB L2 when the launch ends), L2-side writes {int(sum(l2w) / max(len(l2w), 1))} B (= the algorithmic writes)")
            return d["dram_bytes_read"] + d["dram_bytes_write"], note
        except Exception:  # noqa: BLE001
            continue
    return None, "no ncu capture of this instantiation committed"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 50 ms from just before the timed region to the end of the
    roofline loops (same kernels, same conditions; the K-step region itself can be shorter than one sample)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:  # noqa: BLE001
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:  # noqa: BLE001
                continue
            for nm, val in zip(names, r[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": "timed region + roofline loops"}


# ------------------------------------------------------------------------------------------ native arm
class Workload:
    """Device-resident synthetic tensors of one scenario's shape (random-init; there is no dataset)."""

    def __init__(self, dev, rank: int, name: str = "mixgrpo", v32=None):
        from mixgrpo_b200 import rollout as R
        self.name = name
        self.B, self.S, self.dtype, self.flash = SCENARIOS[name]
        B, S = self.B, self.S
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        self.dev = dev
        self.window = list(range(WINDOW))
        self.plan, self.n_steps = step_plan(self.flash, self.window)
        self.cfg = sampler_config(self.flash)
        if name == "mixgrpo_ode_logp_off":
            self.cfg.ode_log_probs = False
        self.sig = R.sigma_schedule(N_STEPS, SHIFT)                        # host schedule: no per-step sync
        self.z0 = torch.randn(B, S, C, device=dev, generator=g).bfloat16()
        if v32 is not None:                                                # same underlying values in another dtype (fp32-vs-bf16 check)
            self.v = [t.to(self.dtype) for t in v32["v"]]
            self.eps = [t.to(self.dtype) for t in v32["eps"]]
            self.z0 = v32["z0"]
        else:
            self.v = [torch.randn(B, S, C, device=dev, generator=g).to(self.dtype) for _ in range(self.n_steps)]
            self.eps = [torch.randn(B, S, C, device=dev, generator=g).to(self.dtype) for _ in range(WINDOW)]
        self.rewards = torch.randn(N_MODELS, B, device=dev, generator=g)
        self.weights = torch.tensor(WEIGHTS, device=dev)
        self.stats_rows = torch.zeros(WINDOW, B, 4, device=dev)   # per (window step, sample): loss, policy, kl, clip_frac
        self.prev_rows = torch.zeros(WINDOW, B, 4, device=dev)    # the other stats buffer (reduced one step late, off the critical path)
        self.side = [torch.cuda.Stream(device=dev) for _ in range(WINDOW)]
        world = dist.get_world_size() if dist.is_initialized() else 1
        self.gbuf = torch.empty(world * N_MODELS, B, device=dev)      # all-gathered rewards, rank-major rows
        self.px = None                                                # mixgrpo_b200.peer.PeerExchange (--collectives peer, N > 1)
        self.gathered = None
        self.policy = "window"                                        # window (2 launches) | pair (4 x 2 on parallel branches) | single
        self.px_stream = torch.cuda.Stream(device=dev)
        self.bytes_per_step = B * S * C * sum(bytes_per_elem(k, self.dtype == torch.float32) * c for k, c in self.plan)


def native_step(w: Workload, v_list=None, eps=None, rewards=None, group=None, collectives=True, comm=None):
    """One GRPO iteration's hot path through the public API (mixgrpo_b200.rollout / .grpo / .peer).
    Returns (stats_rows [W,B,4], all_log_probs [B,N], [grad_v per window step], advantages [B])."""
    from mixgrpo_b200 import grpo, rollout as R
    B, window = w.B, w.window
    v_list = v_list if v_list is not None else w.v
    cur0 = torch.cuda.current_stream(w.dev)
    if comm is not None:
        # --collectives graph: the two NCCL collectives on a side branch, concurrent with the rollout (this step's reward
        # all-gather, the PREVIOUS step's stats all-reduce — logging-only quantities, TR:427-437 / TR:586-600)
        comm.wait_stream(cur0)
        with torch.cuda.stream(comm):
            dist.all_gather_into_tensor(w.gbuf, w.rewards)
            dist.all_reduce(w.stats_rows, op=dist.ReduceOp.AVG)
    rew = rewards if rewards is not None else w.rewards
    adv = None
    if w.px is not None:
        # The path's single exchange AND the advantages as ONE kernel over NVLink peer memory, plus the PREVIOUS step's
        # logging sums as one all-reduce kernel (csrc/peer_kernels.cu) — on a side branch concurrent with the rollout,
        # joined before the policy updates need the advantages.  No NCCL anywhere in the step.
        px_place = os.environ.get("MIXGRPO_BENCH_PX_PLACE", "side")            # tuning: side branch | head of the main chain
        if px_place == "side":
            w.px_stream.wait_stream(cur0)
        with torch.cuda.stream(w.px_stream if px_place == "side" else cur0):
            from mixgrpo_b200 import ops
            adv, w.gathered = w.px.gather_advantages(rew, B, w.weights)
            # the previous step's stats rows are reduced one step late, off the critical path: snapshot them (one tiny launch of
            # ours) before this step's policy update overwrites them — ONE graph serves every step (two alternating graphs with
            # their own buffers cost 6.7 us per step: measured at N = 1, profiles/r02_scaling.md)
            ops.cast_rows(w.stats_rows.view(1, -1), w.prev_rows.view(1, -1))
            w.px.allreduce_stats(w.prev_rows.view(-1))
    det = R.window_mask(N_STEPS, window)
    nz = [None] * N_STEPS
    for j, i in enumerate(window):
        nz[i] = (eps if eps is not None else w.eps)[j]
    _, _, traj, logps, sig_used = R.rollout(lambda lat, s, i: v_list[i], w.z0, w.sig, det, w.cfg, noises=nz)
    if w.px is not None:
        if os.environ.get("MIXGRPO_BENCH_PX_PLACE", "side") == "side":
            cur0.wait_stream(w.px_stream)
    else:
        if collectives and dist.is_initialized() and dist.get_world_size() > 1:
            gathered = grpo.gather_rewards(rew, group)                      # the path's single exchange
            _ = gathered                                                    # feeds logging only in parity mode (TR:427-437)
        adv = grpo.compute_group_advantages(rew, B, w.weights)
    if comm is not None:
        cur0.wait_stream(comm)
    kw = dict(clip_range=CLIP, adv_clip_max=ADV_CLIP, kl_coeff=KL, gradient_accumulation_steps=GA)
    if w.policy == "window":
        # the window's (batch, step) updates are independent given their model outputs (TR:536-585 walks them one by one):
        # ONE forward launch and ONE backward launch for all four; every (step, sample) owns its stats row
        _, _, grads = R.policy_update_window([v_list[t] for t in window], traj, window, logps, adv, sig_used, w.cfg, stats_rows=w.stats_rows,
                                             accumulate=False, early_loads=True, **kw)
    else:
        cur = torch.cuda.current_stream(w.dev)
        grads = [None] * len(window)
        for j, t in enumerate(window):                                      # one stream per window step: their kernels overlap
            s = w.side[j]
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                _, _, grads[j] = R.policy_update(v_list[t], traj[:, t], traj[:, t + 1], logps[:, t], adv, sig_used, t, w.cfg,
                                                 num_train_timesteps=len(window), stats_rows=w.stats_rows[j], accumulate=False,
                                                 single_pass=w.policy == "single", **kw)
        for j in range(len(window)):
            cur.wait_stream(w.side[j])
    if w.px is None and collectives:
        grpo.reduce_step_stats(w.stats_rows, group)
    return w.stats_rows, logps, grads, adv


def capture_step(w: Workload, comm=None):
    """One whole step as a CUDA graph.  With ``comm`` (N > 1, --collectives graph) the two NCCL collectives are captured too,
    on a side branch concurrent with the rollout; returns (graph, outputs, collectives_in_graph, launches per step)."""
    from mixgrpo_b200 import ops
    s = torch.cuda.Stream(device=w.dev)
    s.wait_stream(torch.cuda.current_stream(w.dev))
    with torch.cuda.stream(s):
        before = ops.launch_count
        native_step(w, collectives=False, comm=comm)   # warm-up on the capture stream (allocations, workspaces, coef tables, NCCL)
        launches = ops.launch_count - before           # kernels of OUR library launched by one step (counted, not assumed)
    s.synchronize()
    in_graph = comm is not None
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            out = native_step(w, collectives=False, comm=comm)
    except Exception as e:  # noqa: BLE001  (NCCL capture unavailable: keep the collectives eager on a side stream)
        if comm is None:
            raise
        print(f"[bench] NCCL graph capture failed ({type(e).__name__}: {e}); collectives stay eager", file=sys.stderr)
        torch.cuda.synchronize(w.dev)
        in_graph = False
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            out = native_step(w, collectives=False)
    torch.cuda.current_stream(w.dev).wait_stream(s)
    return g, out, in_graph, launches


def _time_graph(fn, n_launch, stream, reps=20):
    """us per launch of ``fn``'s n_launch launches: captured once, replayed `reps` times between two CUDA events on `stream`."""
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
    del g
    return a.elapsed_time(b) * 1e3 / (reps * n_launch)


def measure_kernels(dev, peak_gbs, B=12, S=4096, full=True):
    """CUDA-event time per launch of every streaming kernel a step uses, each replayed from a CUDA graph over rotating buffer
    sets (every input and output stream rotates; the set is > 126 MB L2) and launched exactly as the step launches it
    (programmatic dependent launch; model output / noise loads ahead of the dependency wait)."""
    from mixgrpo_b200 import _cabi, coefs, ops
    from mixgrpo_b200._cabi import SRC_DETERMINISTIC, SRC_GIVEN, SRC_NOISE, SRC_PHILOX
    ns = 10 if B <= 12 else (8 if B <= 24 else 6)
    g = torch.Generator(device=dev).manual_seed(7)
    xs = [torch.randn(B, S, C, device=dev, generator=g) for _ in range(ns)]
    vs = [torch.randn(B, S, C, device=dev, generator=g).bfloat16() for _ in range(ns)]
    es = [torch.randn(B, S, C, device=dev, generator=g).bfloat16() for _ in range(ns)]
    outs = [torch.empty(B, S, C, device=dev) for _ in range(ns)]
    x0s = [torch.empty(B, S, C, device=dev) for _ in range(ns)]
    gvs = [torch.empty(B, S, C, device=dev, dtype=torch.bfloat16) for _ in range(ns)]
    lps = torch.empty(ns, B, device=dev)
    old = torch.randn(ns, B, device=dev) * 0.01 - 1
    adv = torch.randn(B, device=dev)
    rows = torch.zeros(ns, B, 4, device=dev)
    glp = torch.randn(B, device=dev)
    sig = torch.linspace(1, 0, N_STEPS + 1)
    sig = (SHIFT * sig) / (1 + (SHIFT - 1) * sig)
    k, _ = coefs.flow(sig, 9, ETA, "ref_cuda", True)
    ks = [coefs.flow(sig, t, ETA, "ref_cuda", True)[0] for t in range(WINDOW)]
    e = B * S * C
    res = {}
    s = torch.cuda.Stream(device=dev)

    def rec(name, us, kind, mult=1):
        bpe = bytes_per_elem(kind, False)
        res[name] = {"us_per_launch": round(us, 3), "bytes_per_elem": bpe * mult, "GBps": round(e * bpe * mult / us / 1e3, 1),
                     "frac": round(e * bpe * mult / us / 1e3 / peak_gbs, 4)}

    def loop(fn):
        return lambda: [fn(i) for i in range(ns)]

    common = dict(round_like_torch=True, early=1)
    # as the rollout launches them: a step only accumulates its log-prob sums (MIXGRPO_FLAG_DEFER_LOGP) and ONE finalize launch
    # per rollout writes every step's log-probs — here, like a rollout, 25 step launches (cycling through the buffer sets: a set
    # comes round again only after > 126 MB of other traffic) + 1 finalize inside the timed graph, divided by 25
    nl = N_STEPS
    acc = ops.DeferredLogProbs(dev, nl, B, S * C)
    lp25 = torch.empty(nl, B, device=dev)
    def deferred(fn):
        def run():
            for j in range(nl):
                fn(j % ns, j)
            acc.finalize(lp25)
        return run
    rec("ode", _time_graph(deferred(lambda i, j: ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_DETERMINISTIC, out_x_next=outs[i], want_x0=False, defer=acc.slot(j, k), **common)), nl, s), "ode")
    rec("sde", _time_graph(deferred(lambda i, j: ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_NOISE, noise=es[i], out_x_next=outs[i], want_x0=False, defer=acc.slot(j, k), **common)), nl, s), "sde")
    rec("sde_x0", _time_graph(deferred(lambda i, j: ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_NOISE, noise=es[i], out_x_next=outs[i], want_x0=True, out_x0=x0s[i], defer=acc.slot(j, k), **common)), nl, s), "sde_x0")
    if not full:
        return res
    # the window's policy update as the step launches it: 4 forwards in ONE launch, 4 backwards in ONE launch
    J = WINDOW
    nq = ns // J
    def fwd_multi(q):
        idx = [q * J + j for j in range(J)]
        ops.policy_forward_multi(ops.FLOW, [vs[i] for i in idx], [xs[i] for i in idx], [outs[(i + 1) % ns] for i in idx], ks, [old[i] for i in idx], adv,
                                 CLIP, ADV_CLIP, KL, float(GA * J), stats_rows=[rows[i] for i in idx], round_like_torch=True, out_logps=lps[q * J:(q + 1) * J],
                                 accumulate=False, early_loads=True)
    def bwd_multi(q):
        idx = [q * J + j for j in range(J)]
        ops.policy_backward_multi(ops.FLOW, [vs[i] for i in idx], [xs[i] for i in idx], [outs[(i + 1) % ns] for i in idx], lps[q * J:(q + 1) * J], ks,
                                  [old[i] for i in idx], adv, CLIP, ADV_CLIP, KL, float(GA * J), round_like_torch=True, early_loads=True,
                                  out_grads=[gvs[i] for i in idx])
    rec("train_fwd_x4 (one launch)", _time_graph(lambda: [fwd_multi(q) for q in range(nq)], nq, s), "train_fwd", J)
    rec("bwd_x4 (one launch)", _time_graph(lambda: [bwd_multi(q) for q in range(nq)], nq, s), "bwd", J)
    # the drop-in operators return their log-prob at once: same kernels with the returning atomic (flow_grpo_step & co.)
    rec("ode_immediate", _time_graph(loop(lambda i: ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_DETERMINISTIC, out_x_next=outs[i], out_logp=lps[i], want_x0=False, **common)), ns, s), "ode")
    rec("ode_no_logp", _time_graph(loop(lambda i: ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_DETERMINISTIC, out_x_next=outs[i], want_logp=False, want_x0=False, **common)), ns, s), "ode")
    rec("sde_x0_immediate", _time_graph(loop(lambda i: ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_NOISE, noise=es[i], out_x_next=outs[i], out_logp=lps[i], want_x0=True, out_x0=x0s[i], **common)), ns, s), "sde_x0")
    rec("train_fwd", _time_graph(loop(lambda i: ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_GIVEN, x_next=outs[(i + 1) % ns], out_logp=lps[i], want_x0=False, round_like_torch=True)), ns, s), "train_fwd")
    rec("bwd", _time_graph(loop(lambda i: ops.logprob_backward(ops.FLOW, vs[i], xs[i], outs[(i + 1) % ns], glp, k, True, out=gvs[i])), ns, s), "bwd")
    rec("sde_x0_philox", _time_graph(loop(lambda i: ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_PHILOX, philox=(1234, 4 * i), out_x_next=outs[i], out_logp=lps[i], want_x0=True, out_x0=x0s[i], round_like_torch=True)), ns, s), "sde_x0_philox")
    # Flash tail: DPM-Solver++ order-2 midpoint ODE step, previous x0 as the extra stream (configs[3])
    kdpm, _ = coefs.dpm(sig, 12, 2, "dpmsolver++", "midpoint", "ref_cuda", True)
    rec("dpm2_ode_x0", _time_graph(loop(lambda i: ops.fused_step(ops.DPM, vs[i], xs[i], kdpm, src=SRC_DETERMINISTIC, m1=x0s[(i + 1) % ns], order=2, out_x_next=outs[i], out_logp=lps[i],
                                                                  want_x0=True, out_x0=x0s[i], **common)), ns, s), "dpm2_ode_x0")
    # fp32 model output and noise (configs[4]'s fp32 leg)
    del gvs, es
    v32s = [torch.randn(B, S, C, device=dev, generator=g) for _ in range(ns)]
    e32s = [torch.randn(B, S, C, device=dev, generator=g) for _ in range(ns)]
    k32, _ = coefs.flow(sig, 9, ETA, "ref_cuda", False)
    us = _time_graph(loop(lambda i: ops.fused_step(ops.FLOW, v32s[i], xs[i], k32, src=SRC_NOISE, noise=e32s[i], out_x_next=outs[i], out_logp=lps[i], want_x0=True, out_x0=x0s[i], early=1)), ns, s)
    res["sde_x0_f32"] = {"us_per_launch": round(us, 3), "bytes_per_elem": 20, "GBps": round(e * 20 / us / 1e3, 1), "frac": round(e * 20 / us / 1e3 / peak_gbs, 4)}
    # A/B: the headline kernel without programmatic dependent launch (each launch waits for the previous one to drain)
    _cabi.lib().mixgrpo_set_tuning(1, 0)
    rec("ode_no_pdl", _time_graph(loop(lambda i: ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_DETERMINISTIC, out_x_next=outs[i], out_logp=lps[i], want_x0=False, round_like_torch=True)), ns, s), "ode")
    _cabi.lib().mixgrpo_set_tuning(1, 1)
    return res


def measure_roofline(dev, peak_gbs, peak_kind):
    res = measure_kernels(dev, peak_gbs, 12, 4096, full=True)
    e = 12 * 4096 * C
    top = res["ode"]
    # every streaming launch of one headline step, each at its isolated (cold inputs) time
    plan = [("ode", N_STEPS - WINDOW, 1), ("sde", WINDOW, 1), ("train_fwd_x4 (one launch)", 1, 1), ("bwd_x4 (one launch)", 1, 1)]
    tot_b = sum(res[k]["bytes_per_elem"] * e * c for k, c, _ in plan)
    tot_us = sum(res[k]["us_per_launch"] * c for k, c, _ in plan)
    roof = {"bound": "hbm",
            "kernel": "mg::step_kernel<flow, bf16, SRC_DETERMINISTIC, OUT=0, HALF> — the Euler-ODE sampler step + log-prob (x, v -> x_next, log_prob: 10 B/elem), 21 of a step's 29 launches and the largest share of its device time; launched as the rollout launches it (128-thread CTAs of one half-tile, log-prob sums accumulated as integers, one finalize launch per 25 step launches, its time included)",
            "achieved": top["GBps"], "peak": peak_gbs, "peak_kind": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback",
            "unit": "GB/s", "frac": round(top["GBps"] / peak_gbs, 4), "traffic": None, "us_per_launch": top["us_per_launch"],
            "peak_note": "the measured peak is a COPY (1 byte written per byte read); a sampler step reads 1.5-2x what it writes and, chained by programmatic dependent launch, requests its model output while the previous launch drains — so read-heavy rows under `kernels` / `kernels_by_group_size` can sit at or slightly above 1.0 of this denominator (HBM3e nominal ~7.7 TB/s)",
            "algorithmic_bytes_per_launch": e * bytes_per_elem("ode", False),
            "how": "CUDA events around 20 replays of a CUDA graph of 25 launches + 1 finalize cycling through 10 buffer sets (10 x (6.3 + 12.6 + 12.6) MB > L2: every input is cold), launched as the step launches it (PDL: model-output loads ahead of the dependency wait, the latents — which inside a rollout chain the previous launch wrote — behind it; deferred log-prob finalization)",
            "step_weighted": {"achieved": round(tot_b / tot_us / 1e3, 1), "frac": round(tot_b / tot_us / 1e3 / peak_gbs, 4), "bytes": tot_b, "us": round(tot_us, 2),
                              "what": "all 27 streaming launches of a headline step (21 ode, 4 sde, window forward x4, window backward x4), each at its isolated cold-input time"}}
    return roof, res


def by_group_curve(dev, peak_gbs):
    """Bytes per launch is what sets the roofline fraction of these kernels: the same instantiations when a rank owns 2 or 3 prompt
    groups per launch (train_batch_size > 1, TR:737-749) — the B = 24 / 36 rows."""
    out = {}
    for Bn in (24, 36):
        r = measure_kernels(dev, peak_gbs, Bn, 4096, full=False)
        out[str(Bn)] = {k: {"us_per_launch": v["us_per_launch"], "frac": v["frac"]} for k, v in r.items()}
        torch.cuda.empty_cache()
    return out


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(x: float, dev) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def per_rank(x: float, dev):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [x]
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    out = torch.empty(dist.get_world_size(), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, t)
    return [float(v) for v in out.tolist()]


def h2d_ceiling(dev, nbytes: int, reps: int = 5):
    """What THIS box gives a plain pinned-host -> device cudaMemcpyAsync of `nbytes`, with every rank of the job copying at
    the same time — the ceiling of the e2e leg, whose step time is its upload time.  The pinned block is filled by a D2H
    copy first (lines a CPU just wrote are snooped out of its caches and read 15-25 % slower: measured 45 vs 55 GB/s here).
    Returns GB/s: the better of the best single copy and a run of `reps` back-to-back copies; ranks start together."""
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    st = torch.cuda.Stream(device=dev)
    best = 0.0
    with torch.cuda.stream(st):
        h.copy_(d, non_blocking=True)
        d.copy_(h, non_blocking=True)
        st.synchronize()
        for _ in range(reps):
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            d.copy_(h, non_blocking=True)
            b.record(st)
            b.synchronize()
            best = max(best, nbytes / (a.elapsed_time(b) * 1e-3) / 1e9)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
        b.record(st)
        b.synchronize()
        best = max(best, reps * nbytes / (a.elapsed_time(b) * 1e-3) / 1e9)
    del h, d
    return best


def e2e_run(w: Workload, steps: int, warmup: int, chunks: int, ref_out):
    """Same step with HOST inputs: every step copies its 25 model outputs and the rewards from pinned host memory, runs
    through the public API (eager launches), and reads the stats rows + log-probs back.  The SDE noise is drawn on the
    device, as in the reference (randn_tensor(..., device=model_output.device), SU:189-194) — it is not an input.

    The host inputs ARE the graph leg's synthetic tensors (copied to pinned memory once, outside the timed region), so one
    untimed checked step — same noise as the graph leg — must reproduce the graph leg's log-probs and stats bit for bit.

    The copies are pipelined the way a streaming caller would: two device input sets; while step k computes and its results
    travel back, step k+1's inputs are already on the wire (copy stream; `chunks` cudaMemcpyAsync calls per step, one event
    each, so sampler step i only waits for ITS chunk).  The first timed step's upload is NOT prefetched and the last one
    prefetches nothing, so the timed region contains exactly `steps` uploads and `steps` read-backs."""
    B, S, n = w.B, w.S, w.n_steps
    hv = torch.empty(n, B, S, C, dtype=w.dtype).pin_memory()              # ONE pinned block: 157 MB
    for i in range(n):
        hv[i].copy_(w.v[i])
    hr = torch.empty(N_MODELS, B).pin_memory()
    hr.copy_(w.rewards)                                                    # the graph leg's rewards: seeded, identical on every run
    torch.cuda.synchronize(w.dev)
    sets = [{"dv": torch.empty(n, B, S, C, dtype=w.dtype, device=w.dev), "dr": torch.empty(N_MODELS, B, device=w.dev), "evs": None} for _ in range(2)]
    h_stats = torch.empty(WINDOW, B, 4).pin_memory()
    h_lp = torch.empty(B, n).pin_memory()
    h2d = hv.numel() * hv.element_size() + hr.numel() * 4
    d2h = h_stats.numel() * 4 + h_lp.numel() * 4
    copy_stream = torch.cuda.Stream(device=w.dev)
    main = torch.cuda.current_stream(w.dev)
    gen = torch.Generator(device=w.dev).manual_seed(99)
    bounds = [(c * n) // chunks for c in range(chunks + 1)]
    chunk_of = [max(c for c in range(chunks) if bounds[c] <= i) for i in range(n)]

    def upload(slot):
        # the set was last read by the step before the previous one, which has completed (every step ends with a sync)
        st, evs = sets[slot], []
        with torch.cuda.stream(copy_stream):
            st["dr"].copy_(hr, non_blocking=True)
            for c in range(chunks):
                st["dv"][bounds[c]:bounds[c + 1]].copy_(hv[bounds[c]:bounds[c + 1]], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                evs.append(ev)
        st["evs"] = evs

    def one(k, prefetch_next, eps=None):
        slot = k % 2
        if sets[slot]["evs"] is None:
            upload(slot)
        st = sets[slot]
        evs, dv = st["evs"], st["dv"]

        class Lazy(list):
            def __getitem__(self, i):
                main.wait_event(evs[chunk_of[i]])
                return dv[i]
        main.wait_event(evs[0])
        if eps is None:
            eps = [torch.randn(B, S, C, device=w.dev, dtype=w.dtype, generator=gen) for _ in range(WINDOW)]
        stats, logps, _, _ = native_step(w, v_list=Lazy(), eps=eps, rewards=st["dr"])
        if prefetch_next:
            upload(1 - slot)
        h_stats.copy_(stats, non_blocking=True)
        h_lp.copy_(logps, non_blocking=True)
        main.synchronize()
        st["evs"] = None
        return float(h_stats.sum(dim=(0, 1))[0])

    # checked step: host inputs + the graph leg's noise -> the graph leg's outputs, bitwise.  At N > 1 the graph leg's stats
    # rows were rank-averaged in place by the logging reduction, so this step's rows go through the same reduction first.
    one(0, prefetch_next=False, eps=w.eps)
    got_stats = h_stats.clone()
    if dist.is_initialized() and dist.get_world_size() > 1:
        red = w.stats_rows.clone()
        if w.px is not None:
            w.px.allreduce_stats(red)
        else:
            dist.all_reduce(red, op=dist.ReduceOp.AVG)
        got_stats = red.cpu()
    check = {"e2e_logp_equal": bool(torch.equal(h_lp, ref_out["logps"])), "e2e_stats_equal": bool(torch.equal(got_stats, ref_out["stats"]))}
    for k in range(warmup):
        one(k, prefetch_next=k + 1 < warmup)
    torch.cuda.synchronize(w.dev)
    barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        loss = one(k, prefetch_next=k + 1 < steps)
    torch.cuda.synchronize(w.dev)
    dt = time.perf_counter() - t0
    del sets
    return dt / steps, h2d, d2h, loss, check


def exchange_selfcheck(w: Workload, dev, rank: int, world: int):
    """N > 1: the path's one exchange (TR:332-338, 417-425) and its logging reduction (TR:586-600) computed three ways on
    every rank's own rewards — the fused peer-memory kernels, NCCL + the advantage kernel, the advantage kernel alone — and
    compared; every rank's verdict is AND-ed.  A mismatch fails the bench."""
    from mixgrpo_b200 import grpo
    out = {}
    rew = w.rewards                                                  # differs per rank (seed 1234 + rank)
    adv_local = grpo.compute_group_advantages(rew, w.B, w.weights)
    gathered_nccl = grpo.gather_rewards(rew)                         # [n_models, world*B] in torch.cat order (TR:338)
    ok_gather, ok_adv_err, ok_split_err, ok_ar = True, 0.0, 0.0, True
    ar_err = 0.0
    if w.px is not None:
        adv_peer, gathered_peer = w.px.gather_advantages(rew, w.B, w.weights)                   # mode "local" = the reference's groups
        ok_gather = bool(torch.equal(gathered_peer, gathered_nccl))
        ok_adv_err = float((adv_peer - adv_local).abs().max().item())
        # extended mode: groups of 2B samples in gathered order span two ranks — peer kernel vs NCCL gather + group kernel
        if world % 2 == 0:
            adv_split_peer, _ = w.px.gather_advantages(rew, 2 * w.B, w.weights, mode="split")
            adv_split_nccl = grpo.compute_group_advantages_split(rew, 2 * w.B, w.weights)
            ok_split_err = float((adv_split_peer - adv_split_nccl).abs().max().item())
        # logging reduction: rank-order sum / world (bitwise), and against NCCL's AVG
        x = torch.randn(WINDOW * w.B * 4, device=dev, generator=torch.Generator(device=dev).manual_seed(77 + rank))
        allx = torch.empty(world, x.numel(), device=dev)
        dist.all_gather_into_tensor(allx, x)
        acc = allx[0].clone()
        for q in range(1, world):
            acc = acc + allx[q]
        acc = acc / world
        mine = w.px.allreduce_stats(x.clone())
        ok_ar = bool(torch.equal(mine, acc))
        nc = x.clone()
        dist.all_reduce(nc, op=dist.ReduceOp.AVG)
        ar_err = float((mine - nc).abs().max().item())
        torch.cuda.synchronize(dev)
    flags = torch.tensor([1.0 if ok_gather else 0.0, 1.0 if ok_ar else 0.0, -ok_adv_err, -ok_split_err, -ar_err], dtype=torch.float64, device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    f = flags.tolist()
    out = {"exchange": "peer" if w.px is not None else "nccl", "gathered_equal": bool(f[0] == 1.0), "allreduce_equal": bool(f[1] == 1.0),
           "peer_adv_max_abs_err": -f[2], "split_adv_max_abs_err": -f[3], "allreduce_vs_nccl_avg_max_abs_err": -f[4],
           "what": "every rank: peer gather == NCCL all_gather (bitwise); peer advantages vs the local advantage kernel; groups spanning two ranks vs NCCL gather + group kernel; "
                   "peer all-reduce == rank-order sum / world (bitwise) and vs NCCL AVG; MIN over ranks"}
    return out


def time_scenario(name, dev, rank, world, steps, px, v32=None, keep=False):
    """Whole-step line of one BASELINE config: capture, 3 warm-ups, `steps` replays between CUDA events, max over ranks."""
    w = Workload(dev, rank, name, v32=v32)
    w.px = px
    g, (stats, logps, _, _), _, launches = capture_step(w)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize(dev)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        g.replay()
    b.record()
    torch.cuda.synchronize(dev)
    ms = max_over_ranks(a.elapsed_time(b) / steps, dev)
    line = {"workload": WORKLOAD_TEXT[name], "ms_per_step": round(ms, 4), "value": round(w.bytes_per_step * world / (ms * 1e-3) / 1e9, 1), "unit": "GB/s",
            "algorithmic_bytes_per_step": w.bytes_per_step, "sampler_steps": w.n_steps, "launches_per_step": launches,
            "rollout_steps_per_s": round(w.B * w.n_steps * world / (ms * 1e-3), 1), "steps": steps,
            "loss": float(stats.sum(dim=(0, 1))[0].item()), "logp_mean_window": float(logps[:, w.window].mean().item())}
    res = (line, logps[:, w.window].clone() if keep else None)
    del g
    return res


def run_native(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: a CUDA device is required (the product has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")        # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    import mixgrpo_b200
    from mixgrpo_b200 import ops
    from mixgrpo_b200.grpo_states import GRPOTrainingStates
    mixgrpo_b200.load_library()
    peak, peak_kind = load_peaks()
    w = Workload(dev, rank, args.config)
    w.policy = args.policy
    B, S = w.B, w.S
    states = GRPOTrainingStates(iters_per_group=25, group_size=WINDOW, max_timesteps=N_STEPS - 2, prog_overlap=True, prog_overlap_step=1)
    assert list(states.get_current_timesteps()) == w.window

    main_stream = torch.cuda.current_stream(dev)
    comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    peer_mode = world > 1 and args.collectives == "peer"
    px_check = None
    if world == 1 and os.environ.get("MIXGRPO_BENCH_PEER_N1") == "1":          # tuning: the exchange kernels with no peer
        from mixgrpo_b200.peer import PeerExchange
        w.px = PeerExchange()
    if world > 1:
        from mixgrpo_b200.peer import PeerExchange
        mixgrpo_b200._cabi.lib().mixgrpo_set_tuning(2, 60000)      # a lost peer fails the bench after 60 s instead of 10 min
        try:
            px_check = PeerExchange()
            failed = 0
        except Exception as e:  # noqa: BLE001  (CUDA IPC unavailable on this box: every rank falls back together)
            print(f"[bench] peer exchange unavailable ({type(e).__name__}: {e}); using NCCL on a side stream", file=sys.stderr)
            failed = 1
        t = torch.tensor([failed], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if int(t.item()):
            if px_check is not None:
                px_check.close()
            px_check, peer_mode = None, False
    selfcheck = None
    if world > 1:
        w.px = px_check
        selfcheck = exchange_selfcheck(w, dev, rank, world)
        w.px = px_check if peer_mode else None
    graph, (stats, logps, _, adv_graph), coll_in_graph, launches = capture_step(w, comm_stream if args.collectives == "graph" else None)
    coll_in_graph = coll_in_graph or peer_mode
    # --collectives eager only: two graphs with their own stats rows, used alternately, so step k+1 never has to wait for step
    # k's NCCL all-reduce to finish reading its rows (peer mode snapshots the rows inside its single graph instead)
    graphs, rows = [graph], [w.stats_rows]
    if (world > 1 and not coll_in_graph and args.collectives != "none") or os.environ.get("MIXGRPO_BENCH_TWO_GRAPHS") == "1":
        w.stats_rows, w.prev_rows = w.prev_rows, w.stats_rows
        g2, _, _, _ = capture_step(w, None)
        graphs.append(g2)
        rows.append(w.stats_rows)
    done = [torch.cuda.Event() for _ in graphs]
    counter = [0]

    def step():
        k = counter[0] % len(graphs)
        counter[0] += 1
        if world > 1 and not coll_in_graph and args.collectives != "none":
            main_stream.wait_event(done[k])          # the collectives that read rows[k] two steps ago
        graphs[k].replay()
        if world > 1 and not coll_in_graph and args.collectives != "none":
            comm_stream.wait_stream(main_stream)
            with torch.cuda.stream(comm_stream):
                dist.all_gather_into_tensor(w.gbuf, w.rewards)
                dist.all_reduce(rows[k], op=dist.ReduceOp.AVG)
                done[k].record(comm_stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(dev)
    barrier()
    torch.cuda.synchronize(dev)
    with ClockSampler(local) as clk:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.profiler.start()              # `ncu --profile-from-start off` captures exactly the timed region
        a.record()
        for _ in range(args.steps):
            step()
        if world > 1:
            if peer_mode:
                w.px.allreduce_stats(rows[(counter[0] - 1) % len(graphs)].view(-1))   # the last step's sums (earlier ones were reduced one step late)
            elif coll_in_graph:
                dist.all_reduce(w.stats_rows, op=dist.ReduceOp.AVG)
            elif args.collectives != "none":
                main_stream.wait_stream(comm_stream)  # the last step's collectives end inside the timed region
        b.record()
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        barrier()
        ms = a.elapsed_time(b)
        if args.profile_only:
            if rank == 0:
                emit(json.dumps({"profile_only": True, "steps": args.steps, "ms_per_step": ms / args.steps}))
            if world > 1:
                torch.cuda.synchronize(dev)
                dist.barrier()
                os._exit(0)
            return
        roof, kernels = measure_roofline(dev, peak, peak_kind) if (rank == 0 and not args.skip_e2e) else (None, None)
        if roof is not None:
            roof["traffic"], roof["traffic_note"] = load_ncu_traffic()
    rank_ms = per_rank(ms / args.steps, dev)
    ms_per_step = max(rank_ms)
    total_bytes = w.bytes_per_step * world
    value = total_bytes / (ms_per_step * 1e-3) / 1e9
    # the graph leg's outputs (every replay recomputes the same values from the same inputs)
    ref_out = {"logps": logps.detach().cpu(), "stats": rows[(counter[0] - 1) % len(graphs)].detach().cpu()}
    loss_host = float(ref_out["stats"].sum(dim=(0, 1))[0].item())
    check = {"loss": loss_host, "logp_mean": float(logps[:, w.window[0]].mean().item())}
    if world > 1:
        # the advantages the timed step used (they came over NVLink in peer mode) against the local advantage kernel
        from mixgrpo_b200 import grpo
        w_px, w.px = w.px, None
        adv_local = grpo.compute_group_advantages(w.rewards, B, w.weights)
        w.px = w_px
        err = torch.tensor([float((adv_graph - adv_local).abs().max().item())], dtype=torch.float64, device=dev)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        selfcheck["step_adv_max_abs_err"] = float(err.item())
        check.update(selfcheck)

    if args.skip_e2e:
        if rank == 0:
            emit(json.dumps({"tuning_only": True, "n_gpus": world, "ms_per_step": ms_per_step, "ms_per_step_per_rank": rank_ms, "value": value,
                              "collectives": args.collectives, "in_graph": coll_in_graph, "policy": args.policy, "check": check}))
        if world > 1:
            torch.cuda.synchronize(dev); dist.barrier(); os._exit(0)
        return
    curve = by_group_curve(dev, peak) if (rank == 0 and args.config == "mixgrpo") else None
    barrier()
    e2e_s, h2d, d2h, e2e_loss, e2e_check = e2e_run(w, max(3, min(args.steps, 20)), 3, args.e2e_chunks, ref_out)
    e2e_rank = per_rank(e2e_s, dev)
    e2e_s = max(e2e_rank)
    e2e_value = total_bytes / e2e_s / 1e9
    ceil_rank = per_rank(h2d_ceiling(dev, h2d), dev)
    ok = torch.tensor([1.0 if e2e_check["e2e_logp_equal"] else 0.0, 1.0 if e2e_check["e2e_stats_equal"] else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    check.update({"e2e_loss": e2e_loss, "e2e_logp_equal": bool(ok[0].item() == 1.0), "e2e_stats_equal": bool(ok[1].item() == 1.0),
                  "e2e_what": "one untimed e2e step on the graph leg's inputs and noise reproduces its log-probs and (rank-averaged) stats rows bitwise, on every rank"})

    # BASELINE configs[3] / configs[4] as whole-step lines at the same N (each: own graph, 3 warm-ups, CUDA events, max over ranks)
    configs = None
    if not args.no_configs and args.config == "mixgrpo":
        del graphs, graph
        torch.cuda.empty_cache()
        configs, ksteps = {}, max(5, min(args.steps, 50))
        for nm in ("flash", "large_b24_1024sq", "mixgrpo_ode_logp_off"):
            configs[nm], _ = time_scenario(nm, dev, rank, world, ksteps, w.px)
            torch.cuda.empty_cache()
        # fp32 vs bf16 at (24,1024,64): the SAME model outputs / noise, once in fp32 and once rounded to bf16
        g = torch.Generator(device=dev).manual_seed(4321 + rank)
        Bn, Sn = SCENARIOS["large_b24_512sq_f32"][:2]
        v32 = {"v": [torch.randn(Bn, Sn, C, device=dev, generator=g) for _ in range(N_STEPS)],
               "eps": [torch.randn(Bn, Sn, C, device=dev, generator=g) for _ in range(WINDOW)],
               "z0": torch.randn(Bn, Sn, C, device=dev, generator=g).bfloat16()}
        configs["large_b24_512sq_f32"], lp32 = time_scenario("large_b24_512sq_f32", dev, rank, world, ksteps, w.px, v32=v32, keep=True)
        configs["large_b24_512sq_bf16"], lp16 = time_scenario("large_b24_512sq_bf16", dev, rank, world, ksteps, w.px, v32=v32, keep=True)
        delta = ((lp16 - lp32).abs() / lp32.abs()).max()
        configs["fp32_vs_bf16_logprob"] = {"max_rel_delta": float(delta.item()), "tolerance": 1e-2, "ok": bool(delta.item() < 1e-2),
                                           "what": "window log-probs of the rollout at (24,1024,64) from the same model outputs and noise in fp32 vs rounded to bf16 (torch's bf16 promotion reproduced); "
                                                   "the delta is the input rounding itself — the reference shows the same one — not a kernel error (each dtype is within 1e-4 of the reference in tests/)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_step(None, 1, budget_s=12.0)
    failed = world > 1 and not (check["gathered_equal"] and check["allreduce_equal"] and check["peer_adv_max_abs_err"] == 0.0
                                and check["split_adv_max_abs_err"] <= 1e-6 and check["step_adv_max_abs_err"] == 0.0)
    failed = failed or not (check["e2e_logp_equal"] and check["e2e_stats_equal"])
    if rank == 0:
        coll = ("none (N=1)" if world == 1 else
                "fused peer-memory kernels inside the step graph, no NCCL: reward gather + advantages (1 launch; 64-bit {call,value} words pushed into the peers' memory over NVLink), "
                "[4x12x4] stats all-reduce of the previous step (snapshot + all-reduce: 2 launches), all on a side branch of the graph" if peer_mode else
                "1 all_gather_into_tensor [3x12 f32] + 1 all_reduce [4x12x4 f32] per step, " +
                ("captured in the step graph on a side branch" if coll_in_graph else "eager on a side stream, double-buffered stats rows"))
        line = {
            "metric": "sampler-step latent GB/s", "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": base_config(args.config, world),
            "ms_per_step_per_rank": [round(x, 4) for x in rank_ms],
            "collectives": coll, "launch": "one CUDA graph per step", "policy_update": args.policy,
            "io_dtypes": "model_output/noise bf16 in, latents/trajectory/log-prob fp32, grad bf16; arithmetic fp32",
            "l2": "per step 157 MB of model outputs + 25 MB noise + 327 MB trajectory slots + 25 MB grads stream through (> 126 MB L2); inside the 25-step chain x_{i} (12.6 MB) was "
                  "written by the previous launch and is re-read from L2, which is why the whole-step value can exceed the isolated per-kernel fractions under `kernels` — a real rollout has a DiT forward between sampler steps, "
                  "which is what the isolated (cold-input) figures in `roofline` / `kernels` describe",
            "rollout_steps_per_s": round(B * w.n_steps * world / (ms_per_step * 1e-3), 1),
            "algorithmic_bytes_per_step": w.bytes_per_step,
            "e2e": {"value": round(e2e_value, 2), "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e2e_s * 1e3, 3), "ms_per_step_per_rank": [round(x * 1e3, 3) for x in e2e_rank],
                    "h2d_gbs_per_gpu": round(h2d / e2e_s / 1e9, 2), "h2d_ceiling_gbs": round(min(ceil_rank), 2),
                    "h2d_ceiling_gbs_per_rank": [round(x, 2) for x in ceil_rank], "frac_of_ceiling": round((h2d / e2e_s / 1e9) / min(ceil_rank), 4),
                    "ceiling_how": "tools/h2d_ceiling.py method: plain pinned cudaMemcpyAsync of h2d_bytes_per_step per rank, all ranks at once; better of the best single copy of 5 and 5 back-to-back",
                    "upload": f"{args.e2e_chunks} cudaMemcpyAsync of one pinned block per step, double-buffered across steps",
                    "api": "mixgrpo_b200.rollout.rollout + grpo.compute_group_advantages (peer.PeerExchange.gather_advantages at N > 1) + rollout.policy_update_window, eager launches"},
            "gpu_launches": launches * args.steps + (1 if peer_mode else 0),
            "launches_per_step": launches,
            "clocks": clk.summary(), "roofline": roof, "kernels": kernels, "kernels_by_group_size": curve, "configs": configs, "cpu_baseline": cpu,
            "check": check,
            "library": mixgrpo_b200.library_path(),
        }
        emit(json.dumps(line))
    if world > 1:
        # A live CUDA graph that holds captured NCCL kernels makes destroy_process_group() block at teardown (seen on
        # this image: the line was printed, then the ranks hung).  Drain, rendezvous once more, and leave without it.
        torch.cuda.synchronize(dev)
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(3 if failed else 0)
    if failed:
        raise SystemExit(3)


def base_config(name: str, world: int):
    """The `config` object — identical in both arms."""
    B, S, dt, flash = SCENARIOS[name]
    return {"workload": WORKLOAD_TEXT[name], "group_size": B, "tokens": S, "channels": C, "sampling_steps": N_STEPS, "sde_window": WINDOW,
            "reward_models": N_MODELS, "parallelism": f"dp{world} by prompt group"}


# ------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_reference_step(steps, warmup: int, budget_s: float = 0.0, name: str = "mixgrpo"):
    """The reference's own CPU path for the same step on the host cores.  When the reference drop is present
    (baseline/_ref, shipped with the snapshot; oracle/ref_loader.py executes the UNMODIFIED sampling_utils.py and
    oracle/ref_extract.py the unmodified inline advantage / loss statements of train_grpo_flux.py) that is what runs
    (kind "reference"); otherwise the oracle restatement pinned bit-exact to it (kind "port").  Returns cpu_baseline."""
    from oracle import grpo_oracle as GO
    from oracle import ref_extract, ref_loader
    from oracle import sampling_oracle as O
    B, S, dt, flash = SCENARIOS[name]
    g = torch.Generator().manual_seed(1234)
    sig = O.sd3_time_shift(SHIFT, torch.linspace(1, 0, N_STEPS + 1))
    z0 = torch.randn(B, S, C, generator=g).bfloat16()
    v = [torch.randn(B, S, C, generator=g).bfloat16() for _ in range(N_STEPS)]
    window = list(range(WINDOW))
    eps = {i: torch.randn(B, S, C, generator=g).bfloat16() for i in window}
    rewards = {f"m{j}": torch.randn(B, generator=g) for j in range(N_MODELS)}
    weights = {"m0": WEIGHTS[0], "m1": WEIGHTS[1], "m2": WEIGHTS[2]}
    det = [i not in window for i in range(N_STEPS)]
    ref = ref_loader.load()
    have_tr = ref is not None and ref_extract._train_one_step_ast() is not None
    torch.set_num_threads(os.cpu_count() or 1)          # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)

    def ref_flow(vv, x, i, x_next, noise, determistic):
        if ref is None:
            return O.flow_step(vv, x, ETA, sig, i, x_next, noise, determistic)
        if x_next is None:
            ref_loader.NOISE_QUEUE.append(noise)          # the reference draws through randn_tensor: hand it this step's noise
        return ref.flow_grpo_step(vv, x, ETA, sig, i, x_next, determistic=determistic)

    def advantages():
        if have_tr:
            return ref_extract.reference_advantages(rewards, None, use_group=True, num_generations=B, trimmed_ratio=0.0,
                                                    multi_reward_mix="advantage_aggr", reward_weights=weights)
        return GO.group_advantages(rewards, B, weights)

    def loss_of(lp, old, adv):
        if have_tr:
            return ref_extract.reference_loss(lp, old, adv, clip_range=CLIP, adv_clip_max=ADV_CLIP, kl_coeff=KL,
                                              gradient_accumulation_steps=GA, n_train_timesteps=len(window))[0]
        return GO.grpo_loss(lp, old, adv, CLIP, ADV_CLIP, KL, GA, len(window))[0]

    def one(batched=True, n_samples=B, draw_noise=False):
        """batched: the whole group as one batch with explicit noise (generous to the CPU: the reference loops batch-1 rollouts,
        TR:213,231, and draws a bf16 randn on every step, SU:188-195 — that is `batched=False, draw_noise=True`)."""
        parts = [slice(0, n_samples)] if batched else [slice(b, b + 1) for b in range(n_samples)]
        trajs, lps = [], []
        with torch.no_grad():
            for sl in parts:
                x = z0[sl].to(torch.float32)
                tr, lp = [x], []
                for i in range(N_STEPS):
                    nz = torch.randn(x.shape, dtype=torch.bfloat16) if draw_noise else eps.get(i, z0)[sl]
                    x, _, l, _, _ = ref_flow(v[i][sl], x, i, None, nz, det[i])
                    tr.append(x)
                    lp.append(l)
                trajs.append(torch.stack(tr, dim=1))
                lps.append(torch.stack(lp, dim=1))
        traj, logps = torch.cat(trajs), torch.cat(lps)
        adv = advantages()
        tot = 0.0
        for t in window:
            for sl in parts:
                vt = v[t][sl].clone().requires_grad_(True)
                lp = ref_flow(vt, traj[sl, t], t, traj[sl, t + 1], None, False)[2]
                idx = range(sl.start, sl.stop)
                loss = sum(loss_of(lp[j:j + 1], logps[i:i + 1, t], adv[i:i + 1]) for j, i in enumerate(idx))   # TR:536-585: one sample at a time
                loss.backward()
                tot += float(loss.detach())
        return tot

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    if steps is None:                                   # bounded sample: whole steps until ~budget_s of CPU work
        steps = 0
        while steps < 200 and (steps == 0 or time.perf_counter() - t0 < budget_s):
            loss = one()
            steps += 1
    else:
        for _ in range(steps):
            loss = one()
    dt_s = (time.perf_counter() - t0) / steps
    # the reference's real control flow on a bounded sample: 2 of the 12 samples, batch-1 rollouts, noise drawn per step
    ns = 2
    t1 = time.perf_counter()
    one(batched=False, n_samples=ns, draw_noise=True)
    per_sample_s = (time.perf_counter() - t1) * (B / ns)
    nbytes = algorithmic_bytes_per_step(name)
    return {"value": round(nbytes / dt_s / 1e9, 4), "unit": "GB/s", "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(),
            "kind": "reference" if ref is not None else "port",
            "source": (f"unmodified fastvideo/utils/sampling_utils.py executed from {ref_loader.reference_root()} (flow_grpo_step, explicit noise through its randn_tensor)"
                       + (" + train_grpo_flux.py:440-501 / 560-583 statements cut out with ast and executed" if have_tr else " + oracle/grpo_oracle.py for the inline advantage/loss code")
                       if ref is not None else "oracle/ = torch-CPU restatement, bit-exact vs the reference's own functions (tests/test_oracle_pin.py)"),
            "sample": f"{steps} full step(s) of the same workload (25 sampler steps at (12,4096,64), the group as ONE batch, explicit noise + advantages + 4 window updates with autograd), {warmup} warm-up",
            "s_per_step": round(dt_s, 3), "loss": loss,
            "reference_control_flow": {"value": round(nbytes / per_sample_s / 1e9, 4), "unit": "GB/s", "s_per_step": round(per_sample_s, 3),
                                       "sample": f"{ns} of the {B} samples as batch-1 rollouts with a bf16 randn drawn on every step (TR:213,231; SU:188-195) and per-sample window updates, scaled x{B // ns}"}}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1; the arm uses all host cores
    cpu = cpu_reference_step(args.steps, args.warmup, name=args.config)
    B = SCENARIOS[args.config][0]
    line = {"impl": "reference", "metric": "sampler-step latent GB/s", "value": cpu["value"], "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(cpu["s_per_step"] * 1e3, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": base_config(args.config, world), "device": "host CPU cores (reference PyTorch path)",
            "rollout_steps_per_s": round(B * N_STEPS / cpu["s_per_step"], 2), "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(json.dumps(line))


_json_out = None


def reserve_stdout():
    """stdout carries exactly ONE JSON line: keep a private handle to the real stdout for it and point file descriptor 1 at
    stderr, so nothing a library prints there (NCCL's version banner goes to stdout via printf whatever NCCL_DEBUG_FILE says)
    can land next to it."""
    global _json_out
    if _json_out is None:
        sys.stdout.flush()
        _json_out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(text: str):
    out = _json_out if _json_out is not None else sys.stdout
    out.write(text + "\n")
    out.flush()


def main():
    reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="mixgrpo", choices=["mixgrpo", "flash", "large", "large_b24_1024sq", "large_b24_512sq_bf16", "large_b24_512sq_f32", "mixgrpo_ode_logp_off"],
                    help="which BASELINE config is the line's headline workload (default configs[1]; the default line also carries the others under `configs`)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[3] / configs[4] lines")
    ap.add_argument("--collectives", default="peer", choices=["peer", "graph", "eager", "none"],
                    help="N>1 ('none' = tuning only: no exchange at all, to separate the exchange's cost from the multi-process environment's): 'peer' = fused peer-memory kernels in the step graph (no NCCL); 'graph' = the two NCCL collectives captured in the step graph; 'eager' = NCCL on a side stream")
    ap.add_argument("--policy", default="window", choices=["window", "pair", "single"],
                    help="window update as 2 launches for all 4 steps (default), as 4 x (forward, backward) on parallel branches, or the single-pass kernel per step")
    ap.add_argument("--e2e-chunks", type=int, default=1, help="cudaMemcpyAsync calls per e2e step upload (1 = one 157 MB copy; 25 = one per model output)")
    ap.add_argument("--skip-e2e", action="store_true", help="(tuning only) skip the e2e, roofline and configs legs")
    ap.add_argument("--profile-only", action="store_true", help="setup + warm-up + K timed steps between cudaProfilerStart/Stop, then exit (for ncu)")
    args = ap.parse_args()
    if args.config == "large":
        args.config = "large_b24_1024sq"
    if args.steps is None:
        args.steps = 5 if args.impl == "reference" else 200
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_native(args)


if __name__ == "__main__":
    main()
