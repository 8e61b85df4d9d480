#!/usr/bin/env python
"""bench.py — MixGRPO rollout + policy-update hot path on B200 (driver contract: one JSON line).

Workload (BASELINE.json configs[1]): FLUX.1-dev-shape 1024^2 packed latents (12, 4096, 64), group size 12,
25 sampling steps, SDE window 4, bf16 model output and noise, fp32 latents.  One bench "step" = one GRPO
iteration's hot path for the prompt group(s) this rank owns, on synthetic random-init tensors:

  rollout        25 fused sampler steps (21 Euler-ODE + 4 SDE with log-prob), each reading a distinct
                 pre-generated model output v_i and writing all_latents[:, i+1] in place
  exchange       ONE all-gather of the [3 models x 12] rewards                      (N > 1 only; default: fused with the
  advantages     group-relative, 3 reward models, weighted                         (1 launch)   advantages into one peer-memory kernel)
  policy update  for each of the 4 window steps: fused log-prob + clipped-ratio loss forward, fused loss-grad +
                 log-prob backward -> grad wrt model output                        (2 launches each)
  logging        ONE all_reduce(AVG) of the loss/policy/kl/clip_frac sums          (N > 1 only; default: one peer-memory kernel)

metric  = sampler-step latent GB/s = algorithmic bytes of all sampler/log-prob kernels in the step
          (SURVEY.md §8d per-element figures) / step time, summed over ranks ("weak" scaling: each rank
          owns its own prompt groups; no data-path collective).  rollout_steps_per_s is reported beside it.
roofline= the fused SDE step + log-prob kernel with the reference's full output signature
          (prev_sample, pred_x0, log_prob: 16 B/elem), timed by CUDA events over graph replays on
          rotating buffer sets larger than L2.
e2e     = the same step through the public Python API with HOST (pinned) model outputs and rewards, H2D + D2H copies
          inside the timed region (the SDE noise is drawn on the device, as the reference does).
--impl reference: the reference's algorithm on the host cores (oracle/: torch-CPU restatement pinned
          bit-exact to the reference — the reference itself is pure PyTorch, so this IS its CPU path).
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

B, S, C = 12, 4096, 64          # group size, packed tokens (1024^2), channels
N_STEPS, WINDOW, N_MODELS = 25, 4, 3
ETA, SHIFT = 0.7, 3.0
CLIP, ADV_CLIP, KL, GA = 1e-4, 5.0, 0.01, 3
BYTES = {"ode": 10, "sde": 12, "sde_x0": 16, "train_fwd": 10, "bwd": 12, "sde_x0_philox": 14,   # SURVEY §8d, bf16 v/noise
         "sde_x0_f32": 20, "dpm2_ode_x0": 18}                                                      # fp32 v/noise (configs[4]); Flash DPM-Solver++ order-2 ODE (configs[3])


def algorithmic_bytes_per_step() -> int:
    e = B * S * C
    return e * ((N_STEPS - WINDOW) * BYTES["ode"] + WINDOW * BYTES["sde"] + WINDOW * (BYTES["train_fwd"] + BYTES["bwd"]))


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
    p = ROOT / "profiles" / "step_kernel_ncu.json"
    try:
        d = json.loads(p.read_text())
        return d["dram_bytes_read"] + d["dram_bytes_write"], d.get("note", "")
    except Exception:  # noqa: BLE001
        return None, "no ncu capture committed"


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
    """Device-resident synthetic tensors of the configs[1] shape (random-init; there is no dataset)."""

    def __init__(self, dev, rank: int):
        from mixgrpo_b200 import rollout as R
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        self.dev = dev
        self.cfg = R.SamplerConfig(sampling_steps=N_STEPS, eta=ETA, shift=SHIFT)
        self.sig = R.sigma_schedule(N_STEPS, SHIFT)                        # host schedule: no per-step sync
        self.z0 = torch.randn(B, S, C, device=dev, generator=g).bfloat16()
        self.v = [torch.randn(B, S, C, device=dev, generator=g).bfloat16() for _ in range(N_STEPS)]   # 25 x 6.3 MB
        self.eps = [torch.randn(B, S, C, device=dev, generator=g).bfloat16() for _ in range(WINDOW)]
        self.rewards = torch.randn(N_MODELS, B, device=dev, generator=g)
        self.weights = torch.tensor([1.0, 0.5, 2.0], device=dev)
        self.stats_rows = torch.zeros(WINDOW, B, 4, device=dev)   # per (window step, sample): loss, policy, kl, clip_frac
        self.side = [torch.cuda.Stream(device=dev) for _ in range(WINDOW)]
        world = dist.get_world_size() if dist.is_initialized() else 1
        self.gbuf = torch.empty(world * N_MODELS, B, device=dev)      # all-gathered rewards, rank-major rows
        self.px = None                                                # mixgrpo_b200.peer.PeerExchange (--collectives peer, N > 1)
        self.gathered = None
        self.single_pass = False                                      # --policy single: mixgrpo_policy_step (one launch, 12 B/elem)
        self.prev_rows = torch.zeros(WINDOW, B, 4, device=dev)       # the other stats buffer (reduced one step late, off the critical path)
        self.px_stream = torch.cuda.Stream(device=dev)

    def noises(self, window):
        nz = [None] * N_STEPS
        for j, i in enumerate(window):
            nz[i] = self.eps[j]
        return nz


def native_step(w: Workload, window, v_list=None, eps=None, rewards=None, group=None, collectives=True, parallel=True, comm=None):
    """One GRPO iteration's hot path through the public API (mixgrpo_b200.rollout / .grpo)."""
    from mixgrpo_b200 import grpo, rollout as R
    v_list = v_list if v_list is not None else w.v
    cur0 = torch.cuda.current_stream(w.dev)
    if comm is not None:
        # pipelined collectives on a side branch, concurrent with the rollout: this step's reward all-gather and the
        # PREVIOUS step's stats all-reduce (logging-only quantities, TR:427-437 / TR:586-600); joined before the policy
        # updates overwrite the stats rows
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
        w.px_stream.wait_stream(cur0)
        with torch.cuda.stream(w.px_stream):
            adv, w.gathered = w.px.gather_advantages(rew, B, w.weights)
            w.px.allreduce_stats(w.prev_rows.view(-1))
    det = R.window_mask(N_STEPS, window)
    nz = [None] * N_STEPS
    for j, i in enumerate(window):
        nz[i] = (eps if eps is not None else w.eps)[j]
    _, _, traj, logps, _ = R.rollout(lambda lat, s, i: v_list[i], w.z0, w.sig, det, w.cfg, noises=nz)
    if w.px is not None:
        cur0.wait_stream(w.px_stream)
    else:
        if collectives and dist.is_initialized() and dist.get_world_size() > 1:
            gathered = grpo.gather_rewards(rew, group)                      # the path's single exchange
            _ = gathered                                                    # feeds logging only in parity mode (TR:427-437)
        adv = grpo.compute_group_advantages(rew, B, w.weights)
    if comm is not None:
        cur0.wait_stream(comm)
    # the window's policy updates are independent of one another (TR:536-585 loops over them): one stream each, so
    # their kernels overlap; every (step, sample) owns its stats row -> no zeroing, no race, summed when logged
    cur = torch.cuda.current_stream(w.dev)
    grads = [None] * len(window)
    for j, t in enumerate(window):
        s = w.side[j] if parallel else cur
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            _, _, grads[j] = R.policy_update(v_list[t], traj[:, t], traj[:, t + 1], logps[:, t], adv, w.sig, t, w.cfg, clip_range=CLIP,
                                             adv_clip_max=ADV_CLIP, kl_coeff=KL, gradient_accumulation_steps=GA,
                                             num_train_timesteps=len(window), stats_rows=w.stats_rows[j], accumulate=False,
                                             single_pass=w.single_pass)
    if parallel:
        for j in range(len(window)):
            cur.wait_stream(w.side[j])
    if w.px is None and collectives:
        grpo.reduce_step_stats(w.stats_rows, group)
    return w.stats_rows, logps, grads


LAUNCHES_PER_STEP = 1 + N_STEPS + 1 + 2 * WINDOW  # all ours: trajectory seed + 25 sampler + 1 advantage + 4 x (policy fwd, policy bwd)


def capture_step(w: Workload, window, comm=None):
    """One whole step as a CUDA graph.  With ``comm`` (N > 1) the two NCCL collectives are captured too, on a side branch
    concurrent with the rollout, so a step costs the host ONE graph launch; returns (graph, outputs, collectives_in_graph)."""
    s = torch.cuda.Stream(device=w.dev)
    s.wait_stream(torch.cuda.current_stream(w.dev))
    with torch.cuda.stream(s):
        native_step(w, window, collectives=False, comm=comm)   # warm-up on the capture stream (allocations, workspaces, coef tables, NCCL)
    s.synchronize()
    in_graph = comm is not None
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            out = native_step(w, window, collectives=False, comm=comm)
    except Exception as e:  # noqa: BLE001  (NCCL capture unavailable: keep the collectives eager on a side stream)
        if comm is None:
            raise
        print(f"[bench] NCCL graph capture failed ({type(e).__name__}: {e}); collectives stay eager", file=sys.stderr)
        torch.cuda.synchronize(w.dev)
        in_graph = False
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            out = native_step(w, window, collectives=False)
    torch.cuda.current_stream(w.dev).wait_stream(s)
    return g, out, in_graph


def measure_roofline(dev, peak_gbs, peak_kind):
    """CUDA-event time per launch of the fused SDE step + log-prob kernel (16 B/elem signature) and its siblings,
    replayed from a CUDA graph over 10 rotating buffer sets (10 x 50 MB > 126 MB L2)."""
    from mixgrpo_b200 import coefs, ops
    from mixgrpo_b200._cabi import SRC_DETERMINISTIC, SRC_GIVEN, SRC_NOISE, SRC_PHILOX
    ns = 10
    g = torch.Generator(device=dev).manual_seed(7)
    xs = [torch.randn(B, S, C, device=dev, generator=g) for _ in range(ns)]
    vs = [torch.randn(B, S, C, device=dev, generator=g).bfloat16() for _ in range(ns)]
    es = [torch.randn(B, S, C, device=dev, generator=g).bfloat16() for _ in range(ns)]
    outs = [torch.empty(B, S, C, device=dev) for _ in range(ns)]
    x0s = [torch.empty(B, S, C, device=dev) for _ in range(ns)]      # every output stream rotates too (nothing stays hot in L2)
    gvs = [torch.empty(B, S, C, device=dev, dtype=torch.bfloat16) for _ in range(ns)]
    lps = torch.empty(ns, B, device=dev)
    glp = torch.randn(B, device=dev)
    sig = torch.linspace(1, 0, N_STEPS + 1)
    sig = (SHIFT * sig) / (1 + (SHIFT - 1) * sig)
    k, _ = coefs.flow(sig, 9, ETA, "ref_cuda", True)
    k32, _ = coefs.flow(sig, 9, ETA, "ref_cuda", False)
    kdpm, _ = coefs.dpm(sig, 12, 2, "dpmsolver++", "midpoint", "ref_cuda", True)
    v32s = [torch.randn(B, S, C, device=dev, generator=g) for _ in range(ns)]
    e32s = [torch.randn(B, S, C, device=dev, generator=g) for _ in range(ns)]
    e = B * S * C

    def run(kind, i):
        if kind == "sde_x0":
            ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_NOISE, noise=es[i], out_x_next=outs[i], out_logp=lps[i], want_x0=True, out_x0=x0s[i], round_like_torch=True)
        elif kind == "sde":
            ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_NOISE, noise=es[i], out_x_next=outs[i], out_logp=lps[i], want_x0=False, round_like_torch=True)
        elif kind == "ode":
            ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_DETERMINISTIC, out_x_next=outs[i], out_logp=lps[i], want_x0=False, round_like_torch=True)
        elif kind == "train_fwd":
            ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_GIVEN, x_next=outs[(i + 1) % ns], out_logp=lps[i], want_x0=False, round_like_torch=True)
        elif kind == "bwd":
            ops.logprob_backward(ops.FLOW, vs[i], xs[i], outs[(i + 1) % ns], glp, k, True, out=gvs[i])
        elif kind == "sde_x0_f32":         # fp32 model output and noise (configs[4]'s fp32 leg): 20 B/elem
            ops.fused_step(ops.FLOW, v32s[i], xs[i], k32, src=SRC_NOISE, noise=e32s[i], out_x_next=outs[i], out_logp=lps[i], want_x0=True, out_x0=x0s[i])
        elif kind == "dpm2_ode_x0":        # MixGRPO-Flash tail: DPM-Solver++ order-2 midpoint ODE step, previous x0 as the extra stream
            ops.fused_step(ops.DPM, vs[i], xs[i], kdpm, src=SRC_DETERMINISTIC, m1=x0s[(i + 1) % ns], order=2, out_x_next=outs[i], out_logp=lps[i], want_x0=True,
                           out_x0=x0s[i], round_like_torch=True)
        elif kind == "sde_x0_philox":      # noise drawn in the kernel: no noise tensor is read (and none was generated)
            ops.fused_step(ops.FLOW, vs[i], xs[i], k, src=SRC_PHILOX, philox=(1234, 4 * i), out_x_next=outs[i], out_logp=lps[i], want_x0=True, out_x0=x0s[i], round_like_torch=True)

    res = {}
    s = torch.cuda.Stream(device=dev)
    for kind in ("sde_x0", "sde", "ode", "train_fwd", "bwd", "sde_x0_philox", "sde_x0_f32", "dpm2_ode_x0"):
        with torch.cuda.stream(s):
            for i in range(ns):
                run(kind, i)
            s.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                for i in range(ns):
                    run(kind, i)
            for _ in range(3):
                gr.replay()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            a.record(s)
            for _ in range(reps):
                gr.replay()
            b.record(s)
            b.synchronize()
            us = a.elapsed_time(b) * 1e3 / (reps * ns)
        res[kind] = {"us_per_launch": round(us, 3), "bytes_per_elem": BYTES[kind], "GBps": round(e * BYTES[kind] / us / 1e3, 1)}
        del gr
    # A/B: the headline kernel without programmatic dependent launch (each launch waits for the previous one to drain)
    from mixgrpo_b200 import _cabi
    _cabi.lib().mixgrpo_set_tuning(1, 0)
    with torch.cuda.stream(s):
        for i in range(ns):
            run("sde_x0", i)
        s.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for i in range(ns):
                run("sde_x0", i)
        for _ in range(3):
            gr.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(20):
            gr.replay()
        b.record(s)
        b.synchronize()
        res["sde_x0_no_pdl"] = {"us_per_launch": round(a.elapsed_time(b) * 1e3 / (20 * ns), 3), "bytes_per_elem": 16,
                                "GBps": round(e * 16 / (a.elapsed_time(b) * 1e3 / (20 * ns)) / 1e3, 1)}
    _cabi.lib().mixgrpo_set_tuning(1, 1)
    del gr
    top = res["sde_x0"]
    roof = {"bound": "hbm", "kernel": "mg::step_kernel<flow, bf16, SRC_NOISE> (fused SDE step + log-prob -> prev_sample, pred_x0, log_prob)",
            "achieved": top["GBps"], "peak": peak_gbs, "peak_kind": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback",
            "unit": "GB/s", "frac": round(top["GBps"] / peak_gbs, 4), "traffic": None, "us_per_launch": top["us_per_launch"],
            "algorithmic_bytes_per_launch": e * 16, "how": "CUDA events around 20 replays of a 10-launch CUDA graph over rotating buffer sets (500 MB > L2)"}
    return roof, res


def e2e_run(w: Workload, window, steps: int, warmup: int):
    """Same step with HOST inputs: every step copies its 25 model outputs and the rewards from pinned host memory, runs
    through the public API (eager launches), and reads the stats rows + log-probs back.  The SDE noise is drawn on the
    device, as in the reference (randn_tensor(..., device=model_output.device), SU:189-194) — it is not an input.

    The copies are pipelined the way a streaming caller would: two device input sets; while step k computes and its
    results travel back, step k+1's inputs are already on the wire (copy stream, one event per tensor so sampler step i
    only waits for ITS model output).  The first timed step's upload is NOT prefetched and the last one prefetches
    nothing, so the timed region contains exactly `steps` uploads and `steps` read-backs."""
    hv = [torch.empty(B, S, C, dtype=torch.bfloat16).pin_memory() for _ in range(N_STEPS)]
    hr = torch.randn(N_MODELS, B).pin_memory()
    for t in hv:
        t.normal_()
    sets = [{"dv": [torch.empty_like(t, device=w.dev) for t in hv], "dr": torch.empty(N_MODELS, B, device=w.dev), "evs": None}
            for _ in range(2)]
    h_stats = torch.empty(WINDOW, B, 4).pin_memory()
    h_lp = torch.empty(B, N_STEPS).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in hv) + hr.numel() * 4
    d2h = h_stats.numel() * 4 + h_lp.numel() * 4
    copy_stream = torch.cuda.Stream(device=w.dev)
    main = torch.cuda.current_stream(w.dev)
    gen = torch.Generator(device=w.dev).manual_seed(99)

    def upload(slot):
        # the set was last read by the step before the previous one, which has completed (every step ends with a sync)
        st, evs = sets[slot], []
        with torch.cuda.stream(copy_stream):
            st["dr"].copy_(hr, non_blocking=True)
            for i in range(N_STEPS):
                st["dv"][i].copy_(hv[i], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                evs.append(ev)
        st["evs"] = evs

    def one(k, prefetch_next):
        slot = k % 2
        if sets[slot]["evs"] is None:
            upload(slot)
        st = sets[slot]
        evs, dv = st["evs"], st["dv"]

        class Lazy(list):
            def __getitem__(self, i):
                main.wait_event(evs[i])
                return dv[i]
        main.wait_event(evs[0])
        eps = [torch.randn(B, S, C, device=w.dev, dtype=torch.bfloat16, generator=gen) for _ in range(WINDOW)]
        stats, logps, _ = native_step(w, window, v_list=Lazy(), eps=eps, rewards=st["dr"])
        if prefetch_next:
            upload(1 - slot)
        h_stats.copy_(stats, non_blocking=True)
        h_lp.copy_(logps, non_blocking=True)
        main.synchronize()
        st["evs"] = None
        return float(h_stats.sum(dim=(0, 1))[0])

    for k in range(warmup):
        one(k, prefetch_next=k + 1 < warmup)
    torch.cuda.synchronize(w.dev)
    barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        loss = one(k, prefetch_next=k + 1 < steps)
    torch.cuda.synchronize(w.dev)
    dt = time.perf_counter() - t0
    return dt / steps, h2d, d2h, loss


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(x: float, dev) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


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
    w = Workload(dev, rank)
    w.single_pass = args.policy == "single"
    states = GRPOTrainingStates(iters_per_group=25, group_size=WINDOW, max_timesteps=N_STEPS - 2, prog_overlap=True, prog_overlap_step=1)
    window = states.get_current_timesteps()

    # The path's two tiny collectives (reward all-gather, stats all-reduce) feed logging only in the reference's
    # group mode (TR:427-437, TR:586-600).  They are captured into the step's graph on a side branch that runs
    # concurrently with the rollout (this step's rewards, the previous step's stats), so the host issues one graph
    # launch per step at any N; if NCCL capture is unavailable they run eagerly on a side stream instead.
    main_stream = torch.cuda.current_stream(dev)
    comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    peer_mode = world > 1 and args.collectives == "peer"
    if peer_mode:
        from mixgrpo_b200.peer import PeerExchange
        mixgrpo_b200._cabi.lib().mixgrpo_set_tuning(2, 60000)      # a lost peer fails the bench after 60 s instead of 10 min
        try:
            w.px = PeerExchange()
            failed = 0
        except Exception as e:  # noqa: BLE001  (CUDA IPC unavailable on this box: every rank falls back together)
            print(f"[bench] peer exchange unavailable ({type(e).__name__}: {e}); using NCCL on a side stream", file=sys.stderr)
            failed = 1
        t = torch.tensor([failed], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if int(t.item()):
            if w.px is not None:
                w.px.close()
            w.px, peer_mode = None, False
    graph, (stats, logps, _), coll_in_graph = capture_step(w, window, comm_stream if args.collectives == "graph" else None)
    coll_in_graph = coll_in_graph or peer_mode
    # eager mode: two graphs with their own stats rows, used alternately, so step k+1 never has to wait for step k's
    # all-reduce to finish reading its rows — the collectives overlap the next step completely
    graphs, rows = [graph], [w.stats_rows]
    if world > 1 and (peer_mode or not coll_in_graph):
        w.stats_rows, w.prev_rows = w.prev_rows, w.stats_rows
        g2, _, _ = capture_step(w, window, None)
        graphs.append(g2)
        rows.append(w.stats_rows)
    done = [torch.cuda.Event() for _ in graphs]
    counter = [0]

    def step():
        k = counter[0] % len(graphs)
        counter[0] += 1
        if world > 1 and not coll_in_graph:
            main_stream.wait_event(done[k])          # the collectives that read rows[k] two steps ago
        graphs[k].replay()
        if world > 1 and not coll_in_graph:
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
        # keep the region long enough for nvidia-smi to sample it: repeat the K steps inside the region is NOT allowed,
        # so K is what it is; clocks are additionally sampled over the roofline loop below
        torch.cuda.profiler.start()              # `ncu --profile-from-start off` captures exactly the timed region
        a.record()
        for _ in range(args.steps):
            step()
        if world > 1:
            if peer_mode:
                w.px.allreduce_stats(rows[(counter[0] - 1) % len(graphs)].view(-1))   # the last step's sums (earlier ones were reduced one step late)
            elif coll_in_graph:
                dist.all_reduce(w.stats_rows, op=dist.ReduceOp.AVG)   # the last step's stats (earlier ones were reduced one step late)
            else:
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
    ms_per_step = max_over_ranks(ms / args.steps, dev)
    total_bytes = algorithmic_bytes_per_step() * world
    value = total_bytes / (ms_per_step * 1e-3) / 1e9
    loss_host = float(stats.sum(dim=(0, 1))[0].item())

    if args.skip_e2e:
        if rank == 0:
            emit(json.dumps({"tuning_only": True, "n_gpus": world, "ms_per_step": ms_per_step, "value": value, "collectives": args.collectives,
                              "in_graph": coll_in_graph, "policy": args.policy}))
        if world > 1:
            torch.cuda.synchronize(dev); dist.barrier(); os._exit(0)
        return
    e2e_s, h2d, d2h, e2e_loss = e2e_run(w, window, max(3, min(args.steps, 20)), 3)
    e2e_s = max_over_ranks(e2e_s, dev)
    e2e_value = total_bytes / e2e_s / 1e9

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_step(None, 1, budget_s=12.0)
    if rank == 0:
        line = {
            "metric": "sampler-step latent GB/s", "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "FLUX.1-dev-shape 1024^2 packed latents (12,4096,64), group 12, 25 steps, SDE window 4 (BASELINE configs[1]); "
                                   "one prompt group per GPU", "io_dtypes": "model_output/noise bf16 in, latents/trajectory/log-prob fp32, grad bf16; arithmetic fp32", "group_size": B, "tokens": S, "channels": C, "sampling_steps": N_STEPS,
                       "sde_window": WINDOW, "reward_models": N_MODELS, "parallelism": f"dp{world} by prompt group", "collectives": ("none (N=1)" if world == 1 else "fused peer-memory kernels inside the step graph, no NCCL: reward gather + advantages (1 launch; 64-bit {call,value} words pushed into the peers' memory over NVLink), [4x12x4] stats all-reduce of the previous step (1 launch), both on a side branch of the graph" if peer_mode else "1 all_gather_into_tensor [3x12 f32] + 1 all_reduce [4x12x4 f32] per step, " + ("captured in the step graph on a side branch" if coll_in_graph else "eager on a side stream, double-buffered stats rows (in-graph NCCL measured 3.4x slower at N=8)")),
                       "l2": "inputs larger than L2: per step 157 MB model outputs + 25 MB noise + 327 MB trajectory + 25 MB grads", "launch": "CUDA graph per step"},
            "rollout_steps_per_s": round(B * N_STEPS * world / (ms_per_step * 1e-3), 1),
            "algorithmic_bytes_per_step": algorithmic_bytes_per_step(),
            "e2e": {"value": round(e2e_value, 2), "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e2e_s * 1e3, 3), "api": "mixgrpo_b200.rollout.rollout + grpo.compute_group_advantages (peer.PeerExchange.gather_advantages at N > 1) + rollout.policy_update, eager launches; uploads double-buffered so step k+1's inputs travel while step k computes"},
            "gpu_launches": (LAUNCHES_PER_STEP + (1 if peer_mode else 0)) * args.steps,
            "clocks": clk.summary(), "roofline": roof, "kernels": kernels, "cpu_baseline": cpu,
            "check": {"loss": loss_host, "e2e_loss": e2e_loss, "logp_mean": float(logps[:, window[0]].mean().item())},
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
        os._exit(0)


# ------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_reference_step(steps, warmup: int, budget_s: float = 0.0):
    """The reference's algorithm for the same step on the host cores (oracle = torch-CPU restatement pinned
    bit-exact to the reference's own functions).  Returns the cpu_baseline object."""
    from oracle import grpo_oracle as GO
    from oracle import sampling_oracle as O
    g = torch.Generator().manual_seed(1234)
    sig = O.sd3_time_shift(SHIFT, torch.linspace(1, 0, N_STEPS + 1))
    z0 = torch.randn(B, S, C, generator=g).bfloat16()
    v = [torch.randn(B, S, C, generator=g).bfloat16() for _ in range(N_STEPS)]
    window = list(range(WINDOW))
    eps = {i: torch.randn(B, S, C, generator=g).bfloat16() for i in window}
    rewards = {f"m{j}": torch.randn(B, generator=g) for j in range(N_MODELS)}
    weights = {"m0": 1.0, "m1": 0.5, "m2": 2.0}
    det = [i not in window for i in range(N_STEPS)]

    def one():
        with torch.no_grad():
            _, _, traj, logps = O.rollout(lambda z, s, i: v[i], z0, sig, det, [eps.get(i, z0) for i in range(N_STEPS)], eta=ETA, shift=SHIFT)
        adv = GO.group_advantages(rewards, B, weights)
        tot = 0.0
        for t in window:
            vt = v[t].clone().requires_grad_(True)
            lp = O.flow_step(vt, traj[:, t], ETA, sig, t, traj[:, t + 1])[2]
            loss = sum(GO.grpo_loss(lp[i:i + 1], logps[i:i + 1, t], adv[i:i + 1], CLIP, ADV_CLIP, KL, GA, len(window))[0] for i in range(B))
            loss.backward()
            tot += float(loss.detach())
        return tot

    torch.set_num_threads(os.cpu_count() or 1)          # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
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
    dt = (time.perf_counter() - t0) / steps
    gbs = algorithmic_bytes_per_step() / dt / 1e9
    return {"value": round(gbs, 4), "unit": "GB/s", "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(), "kind": "port",
            "sample": f"{steps} full step(s) of the same workload (25 sampler steps at (12,4096,64) + advantages + 4 window updates with autograd), "
                      f"{warmup} warm-up", "s_per_step": round(dt, 3), "loss": loss,
            "note": "oracle/ = torch-CPU restatement, bit-exact vs the reference's own functions (tests/test_oracle_pin.py); the reference is pure PyTorch"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1; the arm uses all host cores
    cpu = cpu_reference_step(args.steps, min(args.warmup, 1))
    line = {"impl": "reference", "metric": "sampler-step latent GB/s", "value": cpu["value"], "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": round(cpu["s_per_step"] * 1e3, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "FLUX.1-dev-shape 1024^2 packed latents (12,4096,64), group 12, 25 steps, SDE window 4 (BASELINE configs[1])",
                       "device": "host CPU cores (reference PyTorch path)"},
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--collectives", default="peer", choices=["peer", "graph", "eager"],
                    help="N>1: 'peer' = fused peer-memory kernels in the step graph (no NCCL); 'graph' = the two NCCL collectives captured in the step graph; 'eager' = NCCL on a side stream")
    ap.add_argument("--policy", default="pair", choices=["pair", "single"],
                    help="policy update as two launches (log-prob+loss forward, backward: 22 B/elem) or the single-pass kernel (12 B/elem)")
    ap.add_argument("--skip-e2e", action="store_true", help="(tuning only) skip the e2e and roofline legs")
    ap.add_argument("--profile-only", action="store_true", help="setup + warm-up + K timed steps between cudaProfilerStart/Stop, then exit (for ncu)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 5 if args.impl == "reference" else 200
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_native(args)


if __name__ == "__main__":
    main()
