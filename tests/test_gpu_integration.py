"""-m gpu: BASELINE.json configs[0] as an end-to-end parity case — a tiny random-init 2-block FLUX-shaped transformer,
256^2 latents (S = 256 packed tokens), group size 4: mixed ODE/SDE rollout + log-probs + advantages + clipped-ratio
loss + backward INTO THE MODEL PARAMETERS.

Both sides run on the same B200 with the same model and inputs: the "reference" side is the reference's own
PyTorch op sequence (oracle functions on CUDA tensors, pinned bit-exact to the reference incl. on CUDA —
tests/golden/probe_cuda_rounding_b200.json) driven the reference's way (per-sample loop, autograd through step + loss);
the other side is this package (drop-in operators, and the fused batched path).  Identical model outputs on both
sides make the comparison tight: trajectories bit-exact, log-probs 1e-5, parameter gradients 1e-4.
"""
import types

import pytest
import torch

from oracle import grpo_oracle as GO
from oracle import sampling_oracle as O
from tiny_flux import TinyFluxTransformer

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
N, WINDOW, G = 10, [3, 4, 5, 6], 4


def _args(**kw):
    base = dict(w=256, h=256, sampling_steps=N, shift=3.0, eta=0.7, flow_grpo_sampling=True, dpm_algorithm_type="null",
                dpm_apply_strategy="post", dpm_post_compress_ratio=0.4, dpm_solver_order=2, dpm_solver_type="midpoint",
                sample_strategy="progressive", drop_last_sample=False, training_strategy="part", init_same_noise=False,
                clip_range=1e-4, adv_clip_max=5.0, kl_coeff=0.01, gradient_accumulation_steps=2, num_generations=G)
    base.update(kw)
    return types.SimpleNamespace(**base)


def _setup(seed=0):
    torch.manual_seed(seed)
    model = TinyFluxTransformer().to(DEV)
    g = torch.Generator(device=DEV).manual_seed(seed)
    enc = torch.randn(G, 6, 32, device=DEV, generator=g)
    pooled = torch.randn(G, 16, device=DEV, generator=g)
    text_ids = torch.zeros(G, 3, device=DEV)
    lat0 = torch.randn(G, 16, 32, 32, device=DEV, generator=g).bfloat16()
    noises = [torch.randn(G, 256, 64, device=DEV, generator=g).bfloat16() for _ in range(N)]
    rewards = {"hps": torch.randn(G, device=DEV, generator=g), "pick": torch.randn(G, device=DEV, generator=g)}
    return model, enc, pooled, text_ids, lat0, noises, rewards


def _ref_forward(model, z, enc, pooled, text_ids, image_ids, tval):
    ts = torch.full([z.shape[0]], tval, device=DEV, dtype=torch.long)
    with torch.autocast("cuda", torch.bfloat16):
        return model(hidden_states=z, encoder_hidden_states=enc, timestep=ts / 1000, guidance=torch.tensor([3.5], device=DEV, dtype=torch.bfloat16),
                     txt_ids=text_ids[:1].repeat(enc.shape[1], 1), pooled_projections=pooled, img_ids=image_ids, joint_attention_kwargs=None,
                     return_dict=False)[0]


@pytest.mark.parametrize("flow", [True, False])
def test_config0_full_iteration_matches_reference_path(flow):
    from mixgrpo_b200 import grpo, rollout as R, trainer
    args = _args(flow_grpo_sampling=flow)
    model, enc, pooled, text_ids, lat0, noises, rewards = _setup(1 if flow else 2)
    weights = {"hps": 1.0, "pick": 0.5}
    if not flow:
        noises = [n.float() for n in noises]
    det = R.window_mask(N, WINDOW)

    # ---------------- rollout: this package (batched, in-place trajectory) -------------------------------------
    model.eval()
    rew, all_lat, all_lp, sig, image_ids = trainer.sample_reference_model(
        args, DEV, model, enc, pooled, text_ids, lambda lat: rewards, WINDOW, input_latents=lat0, noises=noises)
    assert all_lat.shape == (G, N + 1, 256, 64) and all_lp.shape == (G, N) and all_lat.dtype == torch.float32

    # ---------------- rollout: the reference's op sequence on the same device ---------------------------------
    z0 = GO.pack(lat0, G, 16, 32, 32)
    assert torch.equal(z0, trainer._ops.pack_latents(lat0, G, 16, 32, 32))
    with torch.no_grad():
        ref = O.rollout(lambda z, s, i: _ref_forward(model, z, enc, pooled, text_ids, image_ids, int(s * 1000)), z0, sig, det, noises,
                        eta=args.eta, shift=args.shift, flow_grpo_sampling=flow)
    assert torch.equal(all_lat, ref[2]), (all_lat - ref[2]).abs().max()
    sde = [i for i in range(N) if not det[i]]
    assert torch.allclose(all_lp[:, sde], ref[3][:, sde], rtol=1e-5, atol=0)

    # ---------------- advantages -------------------------------------------------------------------------------
    adv = grpo.compute_group_advantages(rew, G, weights)
    ref_adv = GO.group_advantages({k: v.cpu() for k, v in rewards.items()}, G, weights)
    assert torch.allclose(adv.cpu(), ref_adv, atol=1e-6)

    samples = R.make_samples(all_lat, all_lp, sig, N)
    assert torch.equal(samples["latents"], ref[2][:, :-1][:, :-1]) and samples["timesteps"].shape == (G, N - 1)
    train_ts = [t for t in WINDOW if t < N - 1]
    T = len(train_ts)

    # ---------------- policy update, reference way: per sample, autograd through step + loss ------------------
    def ref_pass():
        model.zero_grad(set_to_none=True)
        model.train()
        tot = torch.zeros(4)
        per_sample = torch.zeros(G, 4)
        for i in range(G):
            for t in train_ts:
                lat = samples["latents"][i:i + 1, t]
                pred = _ref_forward(model, lat, enc[i:i + 1], pooled[i:i + 1], text_ids, image_ids, int(samples["timesteps"][i, t]))
                if flow:
                    lp = O.flow_step(pred, lat, args.eta, sig, t, samples["next_latents"][i:i + 1, t])[2]
                else:
                    lp = O.dance_step(pred, lat, args.eta, sig, t, samples["next_latents"][i:i + 1, t], None, True, True)[2]
                out = GO.grpo_loss(lp, samples["log_probs"][i:i + 1, t], adv[i:i + 1], args.clip_range, args.adv_clip_max, args.kl_coeff,
                                   args.gradient_accumulation_steps, T)
                out[0].backward()
                tot += torch.stack([o.detach() for o in out]).cpu()
                per_sample[i] += torch.stack([o.detach() for o in out]).cpu()
        return tot, per_sample, {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}   # the last block's text branch is dead

    ref_tot, ref_rows, ref_grads = ref_pass()

    # ---------------- (a) drop-in operators inside the reference's loop ----------------------------------------
    model.zero_grad(set_to_none=True)
    tot = torch.zeros(4)
    for i in range(G):
        for t in train_ts:
            lp = trainer.grpo_one_step(args, samples["latents"][i:i + 1, t], samples["next_latents"][i:i + 1, t], enc[i:i + 1], pooled[i:i + 1],
                                       text_ids, image_ids, model, samples["timesteps"][i:i + 1, t], t, sig)
            out = grpo.grpo_loss(lp, samples["log_probs"][i:i + 1, t], adv[i:i + 1], args.clip_range, args.adv_clip_max, args.kl_coeff,
                                 args.gradient_accumulation_steps, T)
            out[0].backward()
            tot += torch.stack([o.detach() for o in out]).cpu()
    assert torch.allclose(tot, ref_tot, rtol=1e-4, atol=1e-6), (tot, ref_tot)
    assert len(ref_grads) > 30
    for n, g_ref in ref_grads.items():
        err = (dict(model.named_parameters())[n].grad - g_ref).norm() / g_ref.norm().clamp_min(1e-20)
        assert err < 1e-4, (n, err.item())

    # ---------------- (b) fused path: train_window (two launches per (micro-batch, step), no loss kernel, no autograd
    # graph for the operator).  micro_batch=1 sees the same model outputs as the reference loop -> tight; micro_batch=2
    # changes the bf16 autocast GEMM shapes, so some model outputs move by a bf16 ulp -> loose.
    for mb, tol in ((1, 1e-4), (2, 1e-2)):
        model.zero_grad(set_to_none=True)
        calls = []
        rows = trainer.train_window(args, model, samples, adv, sig, train_ts, enc, pooled, text_ids, image_ids, micro_batch=mb,
                                    on_accumulated=calls.append)
        assert calls == [1, 3]                                 # gradient_accumulation_steps = 2 over 4 samples (TR:605-609)
        assert torch.allclose(rows.cpu(), ref_rows, rtol=2e-4, atol=1e-6), (rows, ref_rows)   # per sample: the group sum cancels to ~0
        params = dict(model.named_parameters())
        for n, g_ref in ref_grads.items():
            err = (params[n].grad - g_ref).norm() / g_ref.norm().clamp_min(1e-20)
            assert err < tol, (mb, n, err.item())


def test_config0_flash_rollout_with_tiny_transformer():
    """MixGRPO-Flash (dpmsolver++ order 2 after the window, compressed schedule) through the drop-in run_sample_step."""
    from mixgrpo_b200 import rollout as R, trainer
    args = _args(dpm_algorithm_type="dpmsolver++", dpm_apply_strategy="post", dpm_post_compress_ratio=0.5)
    model, enc, pooled, text_ids, lat0, noises, rewards = _setup(3)
    window = [1, 2, 3]
    det = R.window_mask(N, window)
    model.eval()
    _, all_lat, all_lp, sig, image_ids = trainer.sample_reference_model(args, DEV, model, enc, pooled, text_ids, lambda lat: rewards, window,
                                                                       input_latents=lat0, noises=noises)
    z0 = GO.pack(lat0, G, 16, 32, 32)
    with torch.no_grad():
        ref = O.rollout(lambda z, s, i: _ref_forward(model, z, enc, pooled, text_ids, image_ids, int(s * 1000)), z0, sig, det, noises,
                        eta=args.eta, shift=args.shift, flow_grpo_sampling=True, dpm_algorithm_type="dpmsolver++", dpm_apply_strategy="post",
                        dpm_post_compress_ratio=0.5, dpm_solver_order=2, dpm_solver_type="midpoint")
    n_entries = 4 + int(max((N + 1 - 1 - 3) * 0.5, 1))       # sigma[:last_sde+1] + compressed tail (SU:44-54) = 7 -> 6 steps
    assert all_lat.shape == ref[2].shape == (G, n_entries, 256, 64)
    assert torch.equal(all_lat[:, :5], ref[2][:, :5])                                     # Euler + SDE part: bit-exact
    assert ((all_lat - ref[2]).norm() / ref[2].norm()) < 1e-5                             # DPM tail: host vs device exp/log
    assert torch.allclose(all_lp[:, window], ref[3][:, window], rtol=1e-5, atol=0)


def test_config0_train_one_step_composition():
    """trainer.train_one_step (the hot path of TR:341-640 as one call) equals its pieces composed by hand (the test above
    pins those to the reference path), with NCCL-free single-process statistics and with a PeerExchange endpoint."""
    import random
    from mixgrpo_b200 import grpo, rollout as R, trainer
    from mixgrpo_b200.peer import PeerExchange
    args = _args(use_group=True, multi_reward_mix="advantage_aggr", advantage_rerange_strategy="null", trimmed_ratio=0.0)
    model, enc, pooled, text_ids, lat0, noises, rewards = _setup(1)
    weights = {"hps": 1.0, "pick": 0.5}
    enc1, pooled1, tid1 = enc[:1], pooled[:1], text_ids[:1]                       # ONE prompt, repeated num_generations times (TR:369-384)
    enc_r, pooled_r, tid_r = (t.repeat_interleave(G, dim=0) for t in (enc1, pooled1, tid1))
    # ---- by hand
    model.eval()
    rew, all_lat, all_lp, sig, image_ids = trainer.sample_reference_model(args, DEV, model, enc_r, pooled_r, tid_r, lambda lat: rewards, WINDOW,
                                                                          input_latents=lat0, noises=noises)
    adv = grpo.compute_group_advantages(rew, G, weights)
    samples = R.make_samples(all_lat, all_lp, sig, N)
    model.zero_grad(set_to_none=True)
    rows = trainer.train_window(args, model, samples, adv, sig, WINDOW, enc_r, pooled_r, tid_r, image_ids)
    want_stats = rows.sum(dim=0)
    want_grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    assert len(want_grads) > 30 and torch.isfinite(want_stats).all()
    # ---- one call, without and with the fused exchange endpoint (world size 1: same kernels, no peers)
    px = PeerExchange()
    try:
        for exchange in (None, px):
            model.zero_grad(set_to_none=True)
            calls = []
            stats, gathered_res, samples2, adv2 = trainer.train_one_step(
                args, DEV, model, lambda lat: rewards, WINDOW, weights, enc1, pooled1, tid1, exchange=exchange, on_accumulated=calls.append,
                input_latents=lat0, noises=noises)
            assert calls == [1, 3]
            assert torch.equal(adv2, adv) and torch.equal(samples2["latents"], samples["latents"])
            assert torch.allclose(stats, want_stats, rtol=1e-6, atol=1e-9), (stats, want_stats)
            assert set(gathered_res) == {"hps", "pick"} and torch.allclose(gathered_res["pick"], rewards["pick"].mean())
            params = dict(model.named_parameters())
            for n, g_ref in want_grads.items():
                err = (params[n].grad - g_ref).norm() / g_ref.norm().clamp_min(1e-20)
                assert err < 1e-5, (n, err.item())
        # ---- positive/negative re-ranging (TR:527-535): zero-advantage samples are dropped, the rest interleaved
        args_b = _args(use_group=True, advantage_rerange_strategy="balance", trimmed_ratio=0.0)
        model.zero_grad(set_to_none=True)
        calls = []
        stats_b, _, _, adv_b = trainer.train_one_step(args_b, DEV, model, lambda lat: rewards, WINDOW, weights, enc1, pooled1, tid1, exchange=px,
                                                      on_accumulated=calls.append, input_latents=lat0, noises=noises, rng=random.Random(0))
        assert torch.equal(adv_b, adv) and torch.isfinite(stats_b).all()
        assert torch.allclose(stats_b, want_stats, rtol=1e-4, atol=1e-7)            # same samples, another order (none has advantage 0)
        # ---- training_strategy "all" (DanceGRPO style, TR:503-525): per-sample step permutation, a fraction of the steps
        args_a = _args(use_group=True, training_strategy="all", timestep_fraction=0.5, frozen_init_timesteps=0)
        model.zero_grad(set_to_none=True)
        stats_a, _, samples_a, _ = trainer.train_one_step(args_a, DEV, model, lambda lat: rewards, WINDOW, weights, enc1, pooled1, tid1,
                                                          input_latents=lat0, noises=[n for n in noises])
        assert samples_a["latents"].shape == samples["latents"].shape and torch.isfinite(stats_a).all()
        assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
        # ---- use_group False (TR:495-499): one pre-merged reward, global normalisation with the gathered statistics; no prompt repetition
        args_g = _args(use_group=False, multi_reward_mix="reward_aggr")
        merged = rewards["hps"] + 0.5 * rewards["pick"]
        want_adv = GO.group_advantages(merged.cpu(), G, use_group=False, gathered=merged.cpu())
        for exchange in (None, px):
            model.zero_grad(set_to_none=True)
            stats_g, gres, _, adv_g = trainer.train_one_step(args_g, DEV, model, lambda lat: merged, WINDOW, None, enc, pooled, text_ids, exchange=exchange,
                                                             input_latents=lat0, noises=noises)
            assert torch.allclose(adv_g.cpu(), want_adv, atol=1e-6) and torch.isfinite(stats_g).all()
            assert torch.allclose(gres, merged.mean())
    finally:
        px.close()


@pytest.mark.parametrize("drop_last", [False, True])
def test_sample_reference_model_hands_the_vae_its_input(drop_last):
    """TR:284-287 inside the rollout: with ``vae_input=True`` the decode/score callback receives
    ``unpack_latents(latents, h, w, 8) / 0.3611 + 0.1159`` written by the LAST sampler step — bit-identical to the two torch
    ops the reference runs on the device after the loop, the trajectory untouched, one launch fewer."""
    from mixgrpo_b200 import ops, trainer
    args = _args(drop_last_sample=drop_last)
    model, enc, pooled, text_ids, lat0, noises, rewards = _setup(5)
    model.eval()
    seen = {}

    def score_plain(lat):
        seen["plain"] = lat
        return rewards

    def score_vae(x):
        seen["vae"] = x
        return rewards

    before = ops.launch_count
    _, lat_a, lp_a, _, _ = trainer.sample_reference_model(args, DEV, model, enc, pooled, text_ids, score_plain, WINDOW, input_latents=lat0, noises=noises)
    n_plain = ops.launch_count - before
    before = ops.launch_count
    _, lat_b, lp_b, _, _ = trainer.sample_reference_model(args, DEV, model, enc, pooled, text_ids, score_vae, WINDOW, input_latents=lat0, noises=noises,
                                                          vae_input=True)
    assert ops.launch_count - before == n_plain, "the VAE input costs no extra launch"
    assert torch.equal(lat_a, lat_b) and torch.equal(lp_a, lp_b)
    want = GO.unpack(seen["plain"].float(), args.h, args.w, 8) / 0.3611 + 0.1159           # TR:286-287, evaluated on the device like the reference
    assert seen["vae"].shape == (G, 16, 32, 32) and seen["vae"].dtype == torch.float32
    assert torch.equal(seen["vae"], want)
