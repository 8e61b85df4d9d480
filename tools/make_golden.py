"""Generates tests/golden/*.npz / *.json by RUNNING THE REFERENCE ITSELF (CPU) on small seeded inputs.

    python tools/make_golden.py          # needs /root/reference (or $MIXGRPO_REF_ROOT)

The reference cannot travel to the GPU box, so its outputs are committed as fixtures:
  reference_ops_cpu.npz        flow_grpo_step / dance_grpo_step / dpm_step (rollout, train path, autograd grads)
  reference_rollout_cpu.npz    run_sample_step: MixGRPO (flow), DanceGRPO, dpm "all", MixGRPO-Flash "post" midpoint/heun
  reference_grpo_cpu.npz       the inline advantage (TR:440-501) and loss (TR:560-583) statements, executed via ast
  grpo_states_traces.json      GRPOTrainingStates window sequences
  (cuda_reference_b200.npz is produced on the B200 box by tools/probe_cuda_rounding.py.)
bf16 tensors are stored widened to fp32 (exact).
"""
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_extract, ref_loader  # noqa: E402

OUT = ROOT / "tests" / "golden"
ETA, SHIFT, N = 0.7, 3.0, 25
Bq, Sq = 2, 2


def np32(t):
    return t.detach().to(torch.float32).cpu().numpy()


class StandIn(torch.nn.Module):
    """Deterministic stand-in for the FLUX transformer with the exact kwargs of SU:68-82 (returns a 1-tuple, bf16)."""

    def forward(self, hidden_states, encoder_hidden_states, timestep, guidance, txt_ids, pooled_projections, img_ids,
                joint_attention_kwargs, return_dict):
        z = hidden_states.float()
        t = timestep.float().view(-1, 1, 1)
        return (torch.tanh(1.3 * z.roll(1, dims=-1) + t).mul(0.9).add(0.05 * z).to(torch.bfloat16),)


def rollout_args(**kw):
    base = dict(dpm_algorithm_type="null", dpm_apply_strategy="post", dpm_post_compress_ratio=0.4, dpm_solver_order=2,
                dpm_solver_type="midpoint", sample_strategy="progressive", shift=SHIFT, flow_grpo_sampling=True, eta=ETA,
                drop_last_sample=False)
    base.update(kw)
    return types.SimpleNamespace(**base)


ROLLOUT_CASES = {
    "mixgrpo_w4_8": (rollout_args(), [8, 9, 10, 11]),
    "mixgrpo_w0_drop": (rollout_args(drop_last_sample=True), [0, 1, 2, 3]),
    "dance_w5": (rollout_args(flow_grpo_sampling=False), [5, 6, 7, 8]),
    "dpm_all_o2": (rollout_args(dpm_algorithm_type="dpmsolver++", dpm_apply_strategy="all"), [3, 4, 5, 6]),
    "flash_mid_04": (rollout_args(dpm_algorithm_type="dpmsolver++", dpm_apply_strategy="post", dpm_post_compress_ratio=0.4), [2, 3, 4, 5]),
    "flash_heun_02": (rollout_args(dpm_algorithm_type="dpmsolver++", dpm_apply_strategy="post", dpm_post_compress_ratio=0.2,
                                   dpm_solver_type="heun"), [0, 1, 2, 3]),
    "flash_dance_04": (rollout_args(dpm_algorithm_type="dpmsolver++", dpm_apply_strategy="post", flow_grpo_sampling=False), [4, 5, 6, 7]),
}


def main():
    su = ref_loader.load()
    st = ref_loader.load_states()
    assert su is not None and st is not None, "reference tree not found"
    OUT.mkdir(parents=True, exist_ok=True)
    sig = su.sd3_time_shift(SHIFT, torch.linspace(1, 0, N + 1))
    g = torch.Generator().manual_seed(2026)

    def mk(dtype):
        x = torch.randn(Bq, Sq, 64, generator=g)
        v = torch.randn(Bq, Sq, 64, generator=g).to(dtype)
        e = torch.randn(Bq, Sq, 64, generator=g).to(dtype)
        xn = torch.randn(Bq, Sq, 64, generator=g)
        return x, v, e, xn

    ops = {"sigmas": np32(sig)}
    for dtype, dn in ((torch.bfloat16, "bf16"), (torch.float32, "f32")):
        for idx in (0, 1, 3, 12, 23, 24):
            for det in (False, True):
                x, v, e, xn = mk(dtype)
                key = f"flow/{dn}/i{idx}/det{int(det)}"
                ref_loader.NOISE_QUEUE[:] = [e]
                r = su.flow_grpo_step(v, x, ETA, sig, idx, None, determistic=det)
                for nm, t in zip(("x", "v", "eps", "prev", "x0", "logp", "mean", "scale"), (x, v, e, *r)):
                    ops[f"{key}/{nm}"] = np32(t)
                if not det:
                    vg = v.clone().requires_grad_(True)
                    rr = su.flow_grpo_step(vg, x, ETA, sig, idx, xn)
                    (rr[2] * torch.tensor([0.7, -1.3])).sum().backward()
                    ops[f"{key}/xn_train"], ops[f"{key}/train_logp"], ops[f"{key}/train_grad"] = np32(xn), np32(rr[2]), np32(vg.grad)
        for idx in (0, 3, 12, 24):
            for sde in (True, False):
                x, v, e, xn = mk(dtype)
                key = f"dance/{dn}/i{idx}/sde{int(sde)}"
                torch.manual_seed(77 + idx)
                nz = torch.randn_like(x)
                torch.manual_seed(77 + idx)
                r = su.dance_grpo_step(v, x, ETA, sig, idx, None, True, sde)
                vg = v.clone().requires_grad_(True)
                rt = su.dance_grpo_step(vg, x, ETA, sig, idx, xn, True, sde)
                rt[2].sum().backward()
                for nm, t in zip(("x", "v", "noise", "xn_train", "prev", "x0", "logp", "train_logp", "train_grad"),
                                 (x, v, nz, xn, r[0], r[1], r[2], rt[2], vg.grad)):
                    ops[f"{key}/{nm}"] = np32(t)
        for algo in ("dpmsolver++", "dpmsolver"):
            for stype in ("midpoint", "heun"):
                for order in (1, 2, 3):
                    if algo == "dpmsolver" and order == 3:
                        continue
                    for sde in (False, True):
                        args = types.SimpleNamespace(dpm_algorithm_type=algo, dpm_solver_type=stype, dpm_solver_order=order)
                        state = su.DPMState(order=order)
                        for idx in range(N):
                            x, v, e, _ = mk(dtype)
                            e = e.float()
                            hist = [None if m is None else m.clone() for m in state.model_outputs]
                            r = su.dpm_step(args, v, x, idx, sig[:-1], sig, dpm_state=state, variance_noise=e, sde_solver=sde)
                            if idx in (0, 1, 2, 13, 24):
                                key = f"dpm/{dn}/{algo}/{stype}/o{order}/sde{int(sde)}/i{idx}"
                                for nm, t in zip(("x", "v", "eps", "prev", "x0", "logp"), (x, v, e, *r)):
                                    ops[f"{key}/{nm}"] = np32(t)
                                for j, m in enumerate(hist):
                                    if m is not None:
                                        ops[f"{key}/hist{j}"] = np32(m)
    np.savez_compressed(OUT / "reference_ops_cpu.npz", **ops)

    # ---- run_sample_step
    roll = {"sigmas": np32(sig)}
    model = StandIn()
    Br, Sr = 2, 4
    for name, (args, window) in ROLLOUT_CASES.items():
        det = [True] * N
        for i in window:
            det[i] = False
        z = torch.randn(Br, Sr, 64, generator=g).to(torch.bfloat16)
        flow_like = args.flow_grpo_sampling
        n_max = N
        noises = []
        for i in range(n_max):
            if args.dpm_algorithm_type != "null" and args.dpm_apply_strategy == "all":
                noises.append(torch.randn(Br, Sr, 64, generator=g))
            else:
                noises.append(torch.randn(Br, Sr, 64, generator=g).to(torch.bfloat16 if flow_like else torch.float32))
        enc = torch.zeros(Br, 4, 8)
        pooled = torch.zeros(Br, 8)
        text_ids = torch.zeros(Br, 3)
        image_ids = torch.zeros(Sr, 3)
        if flow_like or (args.dpm_algorithm_type != "null" and args.dpm_apply_strategy == "all"):
            # flow_grpo_step draws through randn_tensor at every step it runs; dpm "all" only on SDE steps
            if args.dpm_algorithm_type != "null" and args.dpm_apply_strategy == "all":
                ref_loader.NOISE_QUEUE[:] = [noises[i] for i in range(N) if not det[i]]
            else:
                ref_loader.NOISE_QUEUE[:] = list(noises)
            out = su.run_sample_step(args, z, range(N), sig, model, enc, pooled, text_ids, image_ids, True, det)
        else:
            # dance_grpo_step uses torch.randn_like on SDE steps only: replay the global generator
            torch.manual_seed(4242)
            sde_noise = {}
            for i in range(N):
                if not det[i]:
                    sde_noise[i] = torch.randn(Br, Sr, 64)
            torch.manual_seed(4242)
            out = su.run_sample_step(args, z, range(N), sig, model, enc, pooled, text_ids, image_ids, True, det)
            noises = [sde_noise.get(i, torch.zeros(Br, Sr, 64)) for i in range(N)]
        ref_loader.NOISE_QUEUE[:] = []
        roll[f"{name}/z"] = np32(z)
        roll[f"{name}/window"] = np.array(window)
        roll[f"{name}/noises"] = np.stack([np32(t) for t in noises])
        for nm, t in zip(("z_out", "latents", "all_latents", "all_log_probs"), out):
            roll[f"{name}/{nm}"] = np32(t)
    np.savez_compressed(OUT / "reference_rollout_cpu.npz", **roll)

    # ---- inline GRPO arithmetic, executed from the reference source via ast
    gr = {}
    r3 = {"hps": torch.randn(24, generator=g), "pick": torch.randn(24, generator=g) * 0.02 + 0.3, "ir": torch.randn(24, generator=g)}
    r3["ir"][12:] = 0.25
    w3 = {"hps": 1.0, "pick": 0.5, "ir": 2.0}
    for k, t in r3.items():
        gr[f"adv/rewards/{k}"] = np32(t)
    for ratio in (0.0, 0.2, 0.5):
        a = ref_extract.reference_advantages({k: t.clone() for k, t in r3.items()}, None, use_group=True, num_generations=12,
                                             trimmed_ratio=ratio, multi_reward_mix="advantage_aggr", reward_weights=w3)
        gr[f"adv/advantage_aggr/trim{ratio}"] = np32(a)
        a = ref_extract.reference_advantages(r3["hps"].clone(), None, use_group=True, num_generations=12, trimmed_ratio=ratio,
                                             multi_reward_mix="reward_aggr", reward_weights=None)
        gr[f"adv/reward_aggr/trim{ratio}"] = np32(a)
    gathered = torch.randn(48, generator=g)
    gr["adv/gathered"] = np32(gathered)
    gr["adv/nogroup"] = np32(ref_extract.reference_advantages(gathered[12:24].clone(), gathered, use_group=False, num_generations=12,
                                                              trimmed_ratio=0.0, multi_reward_mix="reward_aggr", reward_weights=None))
    for Bn in (1, 12):
        old = -1.0 + 0.1 * torch.randn(Bn, generator=g)
        new = old + 3e-4 * torch.randn(Bn, generator=g)
        adv = torch.randn(Bn, generator=g) * 3
        if Bn > 2:
            adv[0] = 9.0
            new[1] = old[1]
        for kl in (0.0, 0.01):
            nc = new.clone().requires_grad_(True)
            out = ref_extract.reference_loss(nc, old, adv, clip_range=1e-4, adv_clip_max=5.0, kl_coeff=kl,
                                             gradient_accumulation_steps=3, n_train_timesteps=4)
            out[0].backward()
            key = f"loss/B{Bn}/kl{kl}"
            for nm, t in zip(("old", "new", "adv", "loss", "policy", "kl", "clip_frac", "grad"), (old, new, adv, *out, nc.grad)):
                gr[f"{key}/{nm}"] = np32(t)
    np.savez_compressed(OUT / "reference_grpo_cpu.npz", **gr)

    # ---- window scheduler traces
    cfgs = [
        dict(iters_per_group=25, group_size=4, max_timesteps=23, prog_overlap=True, prog_overlap_step=1),
        dict(iters_per_group=20, group_size=4, max_timesteps=23, prog_overlap=True, prog_overlap_step=0),
        dict(iters_per_group=3, group_size=4, max_timesteps=23),
        dict(iters_per_group=2, group_size=5, max_timesteps=14, roll_back=True),
        dict(iters_per_group=8, group_size=4, max_timesteps=23, sample_strategy="decay", prog_overlap=True, prog_overlap_step=2),
        dict(iters_per_group=8, group_size=4, max_timesteps=23, sample_strategy="decay", max_iters_per_group=10, min_iters_per_group=3, roll_back=True),
        dict(iters_per_group=5, group_size=4, max_timesteps=23, sample_strategy="exp_decay", prog_overlap=True, prog_overlap_step=1),
        dict(iters_per_group=4, group_size=4, max_timesteps=23, sample_strategy="random"),
        dict(iters_per_group=4, group_size=4, max_timesteps=23, cur_timestep=6, roll_back=True, prog_overlap=True, prog_overlap_step=3),
    ]
    traces = []
    for cfg in cfgs:
        s = st.GRPOTrainingStates(**cfg)
        seq = []
        for it in range(160):
            seq.append({"t": [int(x) for x in s.get_current_timesteps()], "done": bool(s.is_training_complete())})
            s.update_iteration(seed=1000 + it) if cfg.get("sample_strategy") == "random" else s.update_iteration()
        traces.append({"config": cfg, "trace": seq})
    (OUT / "grpo_states_traces.json").write_text(json.dumps(traces, separators=(",", ":")))
    for f in sorted(OUT.iterdir()):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
