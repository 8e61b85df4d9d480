"""-m gpu: round-2 additions, all through the C ABI against the CPU oracle / the per-item kernels.

* the log-prob stays finite (and within the 1e-4 gate) however far x_next is from the mean — the reference never
  saturates (SU:201-208); the packed accumulator's side word (csrc/step_math.cuh)
* the window's policy updates batched into one forward + one backward launch: bit-identical to the per-step launches
* the decode-ready second output of the last sampler step vs unpack_latents(...) / 0.3611 + 0.1159 (TR:102-115, TR:286-287)
* graph-safe in-kernel noise: replays of a captured rollout draw different noise; host-seeded Philox refuses capture
* programmatic-dependent-launch early loads change no bit
* the first-order DPM log-prob backward (dpm_apply_strategy == "all", TR:169-180) vs autograd through the oracle
"""
import types

import pytest
import torch

from oracle import grpo_oracle as GO
from oracle import sampling_oracle as O

pytestmark = pytest.mark.gpu

SIG = O.sd3_time_shift(3.0, torch.linspace(1, 0, 26))
ETA = 0.7
CLIP, AMAX, KLC, GA = 1e-4, 5.0, 0.01, 3


def _dev():
    return torch.device("cuda:0")


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


# ------------------------------------------------------------------------------------------ log-prob never saturates
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("far", [1.0, 10.0, 20.0, 25.0, 100.0, 1000.0, 1e5])
def test_logprob_finite_for_far_samples(far, dtype):
    """x_next = mean + far * s * eps: mean(d^2 / 2 s^2) ~ far^2 / 2 = 0.5 ... 5e5.  The packed Q8.32 field holds 255; everything
    above travels through the side accumulators (a warp with a thread beyond the fixed-point range: Q39.24 up to 2^24 per warp,
    integer units up to 2^48).  Reference: finite, = the oracle's value."""
    from mixgrpo_b200 import sampling_utils as su
    d = _dev()
    g = torch.Generator().manual_seed(int(far))
    B, S, idx = 3, 256, 7
    x = torch.randn(B, S, 64, generator=g)
    v = torch.randn(B, S, 64, generator=g).to(dtype)
    eps = torch.randn(B, S, 64, generator=g)
    _, _, _, mean, scale = O.flow_step(v, x, ETA, SIG, idx, x)
    xn = mean + far * scale * eps
    xn[1, :32] += 50 * far * scale                    # sample 1: one CTA's share dwarfs the others (mixed packed / side path)
    ref = O.flow_step(v, x, ETA, SIG, idx, xn)[2]
    got = su.flow_grpo_step(v.to(d), x.to(d), ETA, SIG, idx, xn.to(d), rounding="ref_cpu")[2]
    assert torch.isfinite(got).all(), got
    assert torch.allclose(got.cpu(), ref, rtol=1e-4, atol=0), (got.cpu(), ref)
    again = su.flow_grpo_step(v.to(d), x.to(d), ETA, SIG, idx, xn.to(d), rounding="ref_cpu")[2]
    assert torch.equal(got, again)                    # integer accumulation: bitwise reproducible on the side path too
    # the side word was left zeroed: an ordinary call on the same workspace gives the ordinary answer
    near = mean + scale * eps
    got2 = su.flow_grpo_step(v.to(d), x.to(d), ETA, SIG, idx, near.to(d), rounding="ref_cpu")[2]
    assert torch.allclose(got2.cpu(), O.flow_step(v, x, ETA, SIG, idx, near)[2], rtol=1e-5, atol=0)


def test_logprob_nan_only_for_non_finite_input():
    from mixgrpo_b200 import sampling_utils as su
    d = _dev()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 64, 64, generator=g)
    v = torch.randn(2, 64, 64, generator=g)
    xn = torch.randn(2, 64, 64, generator=g)
    xn[1, 3, 5] = float("inf")
    got = su.flow_grpo_step(v.to(d), x.to(d), ETA, SIG, 4, xn.to(d))[2].cpu()
    assert torch.isfinite(got[0]) and not torch.isfinite(got[1])
    ok = su.flow_grpo_step(v.to(d), x.to(d), ETA, SIG, 4, x.to(d))[2].cpu()       # workspace is clean again
    assert torch.allclose(ok, O.flow_step(v, x, ETA, SIG, 4, x)[2], rtol=1e-5, atol=0)


# ------------------------------------------------------------------------------------------ batched window update
@pytest.mark.parametrize("Bn,S,dtype,flow", [(12, 1024, torch.bfloat16, True), (6, 256, torch.float32, True), (5, 256, torch.bfloat16, False),
                                             (12, 4096, torch.bfloat16, True)])
def test_window_update_two_launches_bit_identical_to_per_step(Bn, S, dtype, flow):
    from mixgrpo_b200 import ops, rollout as R
    d = _dev()
    g = torch.Generator(device=d).manual_seed(11)
    N = 25
    steps = [3, 4, 5, 6]
    traj = torch.randn(Bn, N + 1, S, 64, device=d, generator=g)
    # make the stored transitions plausible: x_{t+1} near the policy mean
    vs = [torch.randn(Bn, S, 64, device=d, generator=g).to(dtype) for _ in steps]
    old = torch.randn(Bn, N, device=d, generator=g) * 0.01 - 1.0
    adv = torch.randn(Bn, device=d, generator=g) * 3
    cfg = R.SamplerConfig(flow_grpo_sampling=flow)
    sig = SIG.to(d)
    rows_ref = torch.zeros(len(steps), Bn, 4, device=d)
    lp_ref, g_ref = [], []
    for j, t in enumerate(steps):
        _, lp, gv = R.policy_update(vs[j], traj[:, t], traj[:, t + 1], old[:, t], adv, sig, t, cfg, clip_range=CLIP, adv_clip_max=AMAX,
                                    kl_coeff=KLC, gradient_accumulation_steps=GA, num_train_timesteps=len(steps), stats_rows=rows_ref[j],
                                    accumulate=False)
        lp_ref.append(lp)
        g_ref.append(gv)
    before = ops.launch_count
    rows = torch.full((len(steps), Bn, 4), 7.0, device=d)
    rows, lp, grads = R.policy_update_window(vs, traj, steps, old, adv, sig, cfg, clip_range=CLIP, adv_clip_max=AMAX, kl_coeff=KLC,
                                             gradient_accumulation_steps=GA, stats_rows=rows, accumulate=False)
    assert ops.launch_count - before == 2, "the whole window must be ONE forward and ONE backward launch"
    assert torch.equal(lp, torch.stack(lp_ref))
    assert torch.equal(rows, rows_ref)
    for a, b in zip(grads, g_ref):
        assert a.dtype == dtype and torch.equal(a, b)
    # accumulate=True adds to the rows; early_loads changes nothing
    rows2 = rows.clone()
    _, lp2, grads2 = R.policy_update_window(vs, traj, steps, old, adv, sig, cfg, clip_range=CLIP, adv_clip_max=AMAX, kl_coeff=KLC,
                                            gradient_accumulation_steps=GA, stats_rows=rows2, accumulate=True, early_loads=True)
    assert torch.equal(lp2, lp) and torch.equal(rows2, rows + rows_ref)
    assert all(torch.equal(a, b) for a, b in zip(grads2, g_ref))


def test_window_update_ragged_falls_back_to_per_step():
    from mixgrpo_b200 import ops, rollout as R
    d = _dev()
    g = torch.Generator(device=d).manual_seed(2)
    Bn, S, steps = 3, 5, [1, 2]                        # n = 5 * 63: not a multiple of 8 -> scalar kernels, per step
    traj = torch.randn(Bn, 6, S, 63, device=d, generator=g)
    vs = [torch.randn(Bn, S, 63, device=d, generator=g) for _ in steps]
    old = torch.randn(Bn, 5, device=d, generator=g)
    adv = torch.randn(Bn, device=d, generator=g)
    cfg = R.SamplerConfig(sampling_steps=5)
    sig = O.sd3_time_shift(3.0, torch.linspace(1, 0, 6))
    before = ops.launch_count
    rows, lp, grads = R.policy_update_window(vs, traj, steps, old, adv, sig, cfg, clip_range=CLIP, adv_clip_max=AMAX, kl_coeff=KLC,
                                             gradient_accumulation_steps=GA)
    assert ops.launch_count - before == 4
    for j, t in enumerate(steps):
        ref = O.flow_step(vs[j].cpu(), traj[:, t].cpu(), ETA, sig, t, traj[:, t + 1].cpu())[2]
        assert torch.allclose(lp[j].cpu(), ref, rtol=1e-5, atol=0)


# ------------------------------------------------------------------------------------------ decode-ready second output
@pytest.mark.parametrize("drop_last", [False, True])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_last_step_emits_the_vae_input(dtype, drop_last):
    """TR:284-287: latents = unpack_latents(latents, h, w, 8); latents = latents / 0.3611 + 0.1159 — here the last sampler
    step's second output, bit-exact against the oracle's unpack + the same fp32 expression, and one launch shorter."""
    from mixgrpo_b200 import ops, rollout as R
    d = _dev()
    g = torch.Generator().manual_seed(4)
    B, h, w, N = 3, 128, 192, 6                        # latent 16 x 24 -> 8 x 12 tokens
    S = (h // 16) * (w // 16)
    z0 = torch.randn(B, S, 64, generator=g).bfloat16()
    vs = [torch.randn(B, S, 64, generator=g).to(dtype) for _ in range(N)]
    sig = O.sd3_time_shift(3.0, torch.linspace(1, 0, N + 1))
    det = [True, False, False, True, True, True]
    nz = [torch.randn(B, S, 64, generator=g).to(dtype) if not det[i] else None for i in range(N)]
    cfg = R.SamplerConfig(sampling_steps=N, drop_last_sample=drop_last, rounding="ref_cpu")
    vd = [t.to(d) for t in vs]
    nd = [t.to(d) if t is not None else None for t in nz]
    # plain rollout + separate unpack launch
    before = ops.launch_count
    _, lat, traj, lps, _ = R.rollout(lambda lt, s, i: vd[i], z0.to(d), sig, det, cfg, noises=nd)
    sep = ops.unpack_latents(lat, h, w, 8, divisor=0.3611, shift=0.1159)
    n_sep = ops.launch_count - before
    # fused
    dec = {"height": h, "width": w, "divisor": 0.3611, "shift": 0.1159}
    before = ops.launch_count
    _, lat2, traj2, lps2, _ = R.rollout(lambda lt, s, i: vd[i], z0.to(d), sig, det, cfg, noises=nd, decode=dec)
    assert ops.launch_count - before == n_sep - 1, "the fused rollout must be one launch shorter"
    assert torch.equal(traj, traj2) and torch.equal(lps, lps2) and torch.equal(lat, lat2)
    want = GO.unpack(lat.cpu(), h, w, 8) / 0.3611 + 0.1159
    assert dec["out"].shape == want.shape == (B, 16, h // 8, w // 8)
    assert torch.equal(dec["out"].cpu(), want)
    assert torch.equal(dec["out"], sep)
    # torch's CUDA div-by-scalar multiplies by the reciprocal: that variant matches the expression evaluated on the device
    dec_r = {"height": h, "width": w, "divisor": 0.3611, "shift": 0.1159, "reciprocal": True}
    R.rollout(lambda lt, s, i: vd[i], z0.to(d), sig, det, cfg, noises=nd, decode=dec_r)
    on_dev = GO.unpack(lat, h, w, 8) / 0.3611 + 0.1159
    assert torch.equal(dec_r["out"], on_dev)


def test_decode_output_rejects_bad_shapes():
    from mixgrpo_b200 import coefs, ops
    from mixgrpo_b200._cabi import SRC_DETERMINISTIC, SRC_GIVEN
    d = _dev()
    v = torch.randn(2, 16, 64, device=d)
    x = torch.randn(2, 16, 64, device=d)
    k, _ = coefs.flow(SIG, 3, ETA, "fp32", False)
    bad = torch.empty(2, 16, 8, 9, device=d)            # C*H*W != n
    with pytest.raises((RuntimeError, ValueError)):
        ops.fused_step(ops.FLOW, v, x, k, src=SRC_DETERMINISTIC, decode={"out": bad})
    ok = torch.empty(2, 16, 8, 8, device=d)
    with pytest.raises(RuntimeError):                   # a stored x_next is not something this launch computes
        ops.fused_step(ops.FLOW, v, x, k, src=SRC_GIVEN, x_next=x, decode={"out": ok})


# ------------------------------------------------------------------------------------------ graph-safe in-kernel noise
def test_inkernel_noise_under_graph_capture_draws_fresh_noise_each_replay():
    from mixgrpo_b200 import ops, rollout as R
    d = _dev()
    g = torch.Generator(device=d).manual_seed(0)
    B, S, N = 2, 64, 4
    z0 = torch.randn(B, S, 64, device=d, generator=g).bfloat16()
    vs = [torch.randn(B, S, 64, device=d, generator=g).bfloat16() for _ in range(N)]
    sig = O.sd3_time_shift(3.0, torch.linspace(1, 0, N + 1))
    det = [True, False, False, True]
    cfg = R.SamplerConfig(sampling_steps=N, inkernel_noise=True)
    st = ops.PhiloxState(d, seed=123)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        R.rollout(lambda lt, sg, i: vs[i], z0, sig, det, cfg, philox_state=st)          # warm-up (workspaces, tables)
        s.synchronize()
        base0 = st.state.clone()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            _, _, traj, lps, _ = R.rollout(lambda lt, sg, i: vs[i], z0, sig, det, cfg, philox_state=st)
        # host-seeded Philox cannot be captured: loud failure instead of silently repeating noise
        with pytest.raises(RuntimeError, match="captured"):
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, stream=s):
                R.rollout(lambda lt, sg, i: vs[i], z0, sig, det, cfg)
    torch.cuda.synchronize()
    gr.replay(); torch.cuda.synchronize()
    a, la = traj.clone(), lps.clone()
    gr.replay(); torch.cuda.synchronize()
    b, lb = traj.clone(), lps.clone()
    assert torch.equal(a[:, 1], b[:, 1])                # the ODE prefix is deterministic
    assert not torch.equal(a[:, 2], b[:, 2])            # SDE steps differ between replays
    assert not torch.equal(la[:, 1], lb[:, 1])
    per = 4 * ((B * S * 64 + 3) // 4) * 2               # two SDE steps per rollout
    assert int(st.state[1] - base0[1]) == 2 * per       # the device-side base moved once per replay
    # and the draws are the documented pure function of (seed, offset, element): replay == eager with the same state
    st2 = ops.PhiloxState(d, seed=123, offset=int(base0[1]) + per)
    _, _, traj_e, _, _ = R.rollout(lambda lt, sg, i: vs[i], z0, sig, det, cfg, philox_state=st2)
    assert torch.equal(traj_e, b)


# ------------------------------------------------------------------------------------------ PDL early loads
@pytest.mark.parametrize("early", [1, 2])
def test_early_loads_change_no_bit(early):
    from mixgrpo_b200 import coefs, ops
    from mixgrpo_b200._cabi import SRC_DETERMINISTIC, SRC_GIVEN, SRC_NOISE
    d = _dev()
    g = torch.Generator(device=d).manual_seed(3)
    B, S = 4, 1024
    x = torch.randn(B, S, 64, device=d, generator=g)
    v = torch.randn(B, S, 64, device=d, generator=g).bfloat16()
    e = torch.randn(B, S, 64, device=d, generator=g).bfloat16()
    k, _ = coefs.flow(SIG, 9, ETA, "ref_cuda", True)
    for src, kw in ((SRC_NOISE, {"noise": e}), (SRC_DETERMINISTIC, {}), (SRC_GIVEN, {"x_next": x.roll(1, 0)})):
        ref = ops.fused_step(ops.FLOW, v, x, k, src=src, round_like_torch=True, **kw)
        for _ in range(3):                                # back to back, so the programmatic edge is really exercised
            got = ops.fused_step(ops.FLOW, v, x, k, src=src, round_like_torch=True, early=early, **kw)
        for a, b in zip(ref[:3], got[:3]):
            assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------ DPM order-1 backward
@pytest.mark.parametrize("algo", ["dpmsolver++", "dpmsolver"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dpm_first_order_logprob_backward_vs_oracle_autograd(algo, dtype):
    """TR:169-180: dpm_step(dpm_state=None, sde_solver=True) inside grpo_one_step — the log-prob of the freshly drawn sample,
    differentiable w.r.t. model_output through the mean."""
    from mixgrpo_b200 import sampling_utils as su
    d = _dev()
    g = torch.Generator().manual_seed(9)
    B, S, idx = 3, 128, 6
    x = torch.randn(B, S, 64, generator=g)
    v = torch.randn(B, S, 64, generator=g).to(dtype)
    nz = torch.randn(B, S, 64, generator=g)
    w = torch.tensor([0.7, -1.3, 2.0])
    args = types.SimpleNamespace(dpm_algorithm_type=algo, dpm_solver_order=2, dpm_solver_type="midpoint")
    vd = v.to(d).requires_grad_(True)
    xn, x0, lp = su.dpm_step(args, vd, x.to(d), idx, SIG[:-1], SIG, dpm_state=None, variance_noise=nz.to(d), sde_solver=True, rounding="ref_cpu")
    (lp * w.to(d)).sum().backward()
    vc = v.clone().requires_grad_(True)
    o_xn, o_x0, o_lp = O.dpm_step(vc, x, idx, 25, SIG, algo=algo, solver_order=2, history=None, noise=nz, sde_solver=True)
    (o_lp * w).sum().backward()
    assert torch.allclose(lp.detach().cpu(), o_lp.detach(), rtol=1e-4, atol=1e-6)
    assert torch.allclose(xn.detach().cpu(), o_xn.detach(), rtol=1e-5, atol=1e-5)
    assert vd.grad.dtype == dtype
    assert _rel(vd.grad.float().cpu(), vc.grad.float()) < (1e-2 if dtype == torch.bfloat16 else 1e-4)
    # higher order / ODE calls have no backward kernel: refuse instead of detaching silently
    with pytest.raises(NotImplementedError):
        su.dpm_step(args, vd, x.to(d), idx, SIG[:-1], SIG, dpm_state=None, sde_solver=False)
    with pytest.raises(NotImplementedError):
        su.flow_grpo_step(vd, x.to(d), ETA, SIG, idx, None, determistic=True)
    with torch.no_grad():
        su.flow_grpo_step(vd, x.to(d), ETA, SIG, idx, None, determistic=True)          # fine without grad mode


# ------------------------------------------------------------------------------------------ ODE-step log-probs are optional
@pytest.mark.parametrize("flash", [False, True])
def test_rollout_without_ode_log_probs_changes_nothing_that_is_read(flash):
    """SamplerConfig.ode_log_probs=False: deterministic steps skip the reduction.  The trajectory and the SDE-window log-probs
    (all train_one_step ever reads, TR:536-553) are bit-identical; the skipped columns are NaN, not stale memory."""
    from mixgrpo_b200 import rollout as R
    d = _dev()
    g = torch.Generator(device=d).manual_seed(9)
    B, S, N = 4, 256, 25
    window = [2, 3, 4, 5]
    kw = dict(dpm_algorithm_type="dpmsolver++", dpm_apply_strategy="post") if flash else {}
    z0 = torch.randn(B, S, 64, device=d, generator=g).bfloat16()
    vs = [torch.randn(B, S, 64, device=d, generator=g).bfloat16() for _ in range(N)]
    nz = [torch.randn(B, S, 64, device=d, generator=g).bfloat16() if i in window else None for i in range(N)]
    det = R.window_mask(N, window)
    sig = R.sigma_schedule(N, 3.0)
    a = R.rollout(lambda lt, s, i: vs[i], z0, sig, det, R.SamplerConfig(**kw), noises=nz)
    b = R.rollout(lambda lt, s, i: vs[i], z0, sig, det, R.SamplerConfig(ode_log_probs=False, **kw), noises=nz)
    assert torch.equal(a[2], b[2]) and torch.equal(a[1], b[1])
    assert torch.equal(a[3][:, window], b[3][:, window])
    ode = [i for i in range(a[3].shape[1]) if i not in window]
    assert torch.isnan(b[3][:, ode]).all() and torch.isfinite(a[3][:, [i for i in ode if i < a[3].shape[1] - 1]]).all()


# ------------------------------------------------------------------------------------------ deferred log-prob finalization
@pytest.mark.parametrize("family", ["flow", "dance", "flash", "dpm_all"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_deferred_log_probs_are_bit_identical_to_the_immediate_path(family, dtype):
    """MIXGRPO_FLAG_DEFER_LOGP + mixgrpo_logp_finalize: the rollout's step launches only accumulate, ONE launch finalizes — the
    same packed integer sums, so all_log_probs (and everything else) must not change by a bit; one launch more in total."""
    from mixgrpo_b200 import ops, rollout as R
    d = _dev()
    g = torch.Generator(device=d).manual_seed(21)
    B, S, N = 5, 320, 12                                # 160 tiles... n = 20480: 10 CTAs per sample
    window = [2, 3, 4, 5]
    kw = {"flow": {}, "dance": dict(flow_grpo_sampling=False), "flash": dict(dpm_algorithm_type="dpmsolver++", dpm_apply_strategy="post"),
          "dpm_all": dict(dpm_algorithm_type="dpmsolver++", dpm_apply_strategy="all", dpm_solver_order=3)}[family]
    z0 = torch.randn(B, S, 64, device=d, generator=g).bfloat16()
    vs = [torch.randn(B, S, 64, device=d, generator=g).to(dtype) for _ in range(N)]
    nd = dtype if family in ("flow", "flash") else torch.float32
    nz = [torch.randn(B, S, 64, device=d, generator=g).to(nd) if i in window else None for i in range(N)]
    det = R.window_mask(N, window)
    sig = R.sigma_schedule(N, 3.0)
    before = ops.launch_count
    a = R.rollout(lambda lt, s, i: vs[i], z0, sig, det, R.SamplerConfig(sampling_steps=N, defer_log_probs=False, **kw), noises=nz)
    n_imm = ops.launch_count - before
    before = ops.launch_count
    b = R.rollout(lambda lt, s, i: vs[i], z0, sig, det, R.SamplerConfig(sampling_steps=N, defer_log_probs=True, **kw), noises=nz)
    assert ops.launch_count - before == n_imm + 1
    assert torch.equal(a[2], b[2]) and torch.equal(a[1], b[1])
    assert torch.equal(torch.nan_to_num(a[3], nan=123.0), torch.nan_to_num(b[3], nan=123.0))      # the last DPM step's NaN (sigma_t = 0) included
    # twice in a row on the same cached records: finalize left them zeroed
    c = R.rollout(lambda lt, s, i: vs[i], z0, sig, det, R.SamplerConfig(sampling_steps=N, **kw), noises=nz)
    assert torch.equal(torch.nan_to_num(c[3], nan=123.0), torch.nan_to_num(a[3], nan=123.0))


def test_deferred_log_probs_far_samples_inactive_rows_and_graph_replay():
    from mixgrpo_b200 import coefs, ops
    from mixgrpo_b200._cabi import SRC_DETERMINISTIC, SRC_GIVEN
    d = _dev()
    g = torch.Generator().manual_seed(8)
    B, S, idx = 3, 256, 7
    x = torch.randn(B, S, 64, generator=g)
    v = torch.randn(B, S, 64, generator=g)
    _, _, _, mean, scale = O.flow_step(v, x, ETA, SIG, idx, x)
    xn = mean + 1000.0 * scale * torch.randn(B, S, 64, generator=g)     # shares beyond the packed field: side accumulators, deferred
    xn[1, :32] += 50 * 1000.0 * scale
    ref = O.flow_step(v, x, ETA, SIG, idx, xn)[2]
    k, _ = coefs.flow(SIG, idx, ETA, "fp32", False)
    vd, xd, xnd = v.to(d), x.to(d), xn.to(d)
    out = torch.zeros(3, B, device=d)

    def run():
        acc = ops.DeferredLogProbs(d, 3, B, S * 64)
        ops.fused_step(ops.FLOW, vd, xd, k, src=SRC_GIVEN, x_next=xnd, want_x0=False, defer=acc.slot(0, k))
        ops.fused_step(ops.FLOW, vd, xd, k, src=SRC_DETERMINISTIC, want_x0=False, defer=acc.slot(2, k))      # launch 1 never takes a slot
        acc.finalize(out)

    run()
    imm = ops.fused_step(ops.FLOW, vd, xd, k, src=SRC_GIVEN, x_next=xnd, want_x0=False)[2]
    assert torch.equal(out[0], imm) and torch.allclose(out[0].cpu(), ref, rtol=1e-4, atol=0)
    assert torch.isnan(out[1]).all()
    assert torch.equal(out[2], ops.fused_step(ops.FLOW, vd, xd, k, src=SRC_DETERMINISTIC, want_x0=False)[2])
    want = out.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        run()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            run()
        for _ in range(3):
            out.zero_()
            gr.replay()
            s.synchronize()
            assert torch.equal(torch.nan_to_num(out, nan=5.0), torch.nan_to_num(want, nan=5.0))
    torch.cuda.current_stream().wait_stream(s)
    with pytest.raises(ValueError):
        ops.DeferredLogProbs(d, 3, B, S * 64).finalize(torch.zeros(2, B, device=d))
    with pytest.raises(RuntimeError):      # the fused loss needs its log-prob inside the launch: the flag is refused there
        from mixgrpo_b200 import _cabi
        import ctypes as C
        la = _cabi.LossArgs()
        la.old_logp, la.advantages = out[0].data_ptr(), out[0].data_ptr()
        ws = torch.zeros(4096, dtype=torch.uint8, device=d)
        rc = _cabi.lib().mixgrpo_policy_fwd(0, vd.data_ptr(), 0, xd.data_ptr(), S * 64, xnd.data_ptr(), S * 64, out[1].data_ptr(), ws.data_ptr(), ws.numel(), B,
                                            S * 64, C.byref(k), C.byref(la), _cabi.FLAG_DEFER_LOGP, torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc, "policy_fwd")


@pytest.mark.parametrize("S", [4096, 100, 17])
@pytest.mark.parametrize("B", [3, 12])
def test_deferred_half_tile_ctas_are_bit_identical_to_both_256_thread_paths(S, B):
    """A deferred launch runs as 128-thread CTAs of ONE half-tile (1024 scalars) whose reductions are spread over 8 sub-records per
    sample; the 256-thread kernels convert each of their two half-tiles to fixed point on its own and add integers.  So the
    deferred launch in either shape (mixgrpo_set_tuning key 6) and the immediate launch give the same bits — at the full size,
    with a ragged last tile (n = 6400: 7 half-tiles, 4 tiles) and below one half-tile; far samples (side accumulators) included."""
    from mixgrpo_b200 import _cabi, coefs, ops
    from mixgrpo_b200._cabi import SRC_DETERMINISTIC, SRC_GIVEN, SRC_NOISE
    d = _dev()
    g = torch.Generator(device=d).manual_seed(77)
    idx = 5
    x = torch.randn(B, S, 64, device=d, generator=g)
    v = torch.randn(B, S, 64, device=d, generator=g).bfloat16()
    e = torch.randn(B, S, 64, device=d, generator=g).bfloat16()
    k, _ = coefs.flow(SIG, idx, ETA, "ref_cuda", True)
    far = x + 3.0e4 * torch.randn(B, S, 64, device=d, generator=g)          # |d|/s ~ 1e5: every part goes to a side accumulator
    far[1] = x[1] + 0.3 * torch.randn(S, 64, device=d, generator=g)         # ... except sample 1's
    cases = [dict(src=SRC_DETERMINISTIC), dict(src=SRC_NOISE, noise=e), dict(src=SRC_GIVEN, x_next=far)]
    lib = _cabi.lib()
    for kw in cases:
        imm = ops.fused_step(ops.FLOW, v, x, k, want_x0=False, round_like_torch=True, **kw)
        got = {}
        for half in (2, 0):                                              # 2 = the 128-thread shape whatever the grid size, 0 = never
            old = lib.mixgrpo_set_tuning(6, half)
            try:
                acc = ops.DeferredLogProbs(d, 2, B, S * 64)
                out = torch.zeros(2, B, device=d)
                xn = torch.empty_like(x) if kw["src"] != SRC_GIVEN else None
                ops.fused_step(ops.FLOW, v, x, k, want_x0=False, round_like_torch=True, out_x_next=xn, defer=acc.slot(1, k), **kw)
                acc.finalize(out)
            finally:
                lib.mixgrpo_set_tuning(6, old)
            assert torch.isnan(out[0]).all()
            got[half] = out[1].clone()
            if xn is not None:
                assert torch.equal(xn, imm[0])
        assert torch.isfinite(imm[2]).all()
        assert torch.equal(got[2], imm[2]) and torch.equal(got[0], imm[2]), (kw["src"], got, imm[2])
    # the default (key 6 = 1) picks the 128-thread shape only when the 256-thread grid is more than one wave of 6 CTAs per SM
    sms = torch.cuda.get_device_properties(d).multi_processor_count
    tiles = (S * 64 + 2047) // 2048
    before = lib.mixgrpo_set_tuning(8, 0)
    acc = ops.DeferredLogProbs(d, 2, B, S * 64)
    ops.fused_step(ops.FLOW, v, x, k, want_x0=False, round_like_torch=True, src=SRC_DETERMINISTIC, defer=acc.slot(0, k))
    acc.finalize(torch.zeros(2, B, device=d))
    assert lib.mixgrpo_set_tuning(8, 0) - before == (1 if tiles * B > 6 * sms else 0)


# ------------------------------------------------------------------------------------------ trajectory seed fused into step 0
@pytest.mark.parametrize("first_sde", [False, True])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_first_step_seeds_the_trajectory(first_sde, dtype):
    """all_latents[:, 0] = float(z) (SU:26, SU:153) written by the first sampler step itself (mixgrpo_step_ext.x_f32_out): the
    whole rollout is bit-identical to seeding with the separate cast launch, and one launch shorter."""
    from mixgrpo_b200 import coefs, ops, rollout as R
    from mixgrpo_b200._cabi import SRC_DETERMINISTIC, SRC_NOISE
    d = _dev()
    g = torch.Generator(device=d).manual_seed(31)
    B, S, N = 4, 512, 6
    window = [0, 1] if first_sde else [2, 3]
    z0 = torch.randn(B, S, 64, device=d, generator=g).bfloat16()
    vs = [torch.randn(B, S, 64, device=d, generator=g).to(dtype) for _ in range(N)]
    nz = [torch.randn(B, S, 64, device=d, generator=g).to(dtype) if i in window else None for i in range(N)]
    det = R.window_mask(N, window)
    sig = R.sigma_schedule(N, 3.0)
    cfg = R.SamplerConfig(sampling_steps=N)
    before = ops.launch_count
    a = R.rollout(lambda lt, s, i: vs[i], z0, sig, det, cfg, noises=nz)
    n_fused = ops.launch_count - before
    assert n_fused == N + 1, "N step launches + 1 finalize: no cast launch"
    assert torch.equal(a[2][:, 0], z0.float())
    # reference composition: cast launch + every step on the fp32 slot
    traj = torch.empty(B, N + 1, S, 64, device=d)
    ops.cast_rows(z0, traj[:, 0])
    lps = torch.empty(N, B, device=d)
    for i in range(N):
        k, _ = coefs.flow(sig, i, cfg.eta, "ref_cuda", dtype == torch.bfloat16)
        ops.fused_step(ops.FLOW, vs[i], traj[:, i], k, src=SRC_DETERMINISTIC if det[i] else SRC_NOISE, noise=nz[i], out_x_next=traj[:, i + 1],
                       out_logp=lps[i], want_x0=False, round_like_torch=True)
    assert torch.equal(a[2], traj) and torch.equal(a[3], lps.t())
    # a float32 z (or an unaligned one) falls back to the cast launch and gives the same trajectory
    b = R.rollout(lambda lt, s, i: vs[i], z0.float(), sig, det, cfg, noises=nz)
    assert torch.equal(b[2], traj)
    # the entry point refuses the seed for the stored-transition source and for non-bf16 latents
    k, _ = coefs.flow(sig, 0, cfg.eta, "ref_cuda", dtype == torch.bfloat16)
    with pytest.raises(ValueError):
        ops.fused_step(ops.FLOW, vs[0], z0.float(), k, src=SRC_DETERMINISTIC, seed_out=traj[:, 0])


def test_deferred_finalize_handles_more_than_one_chunk_of_launches_and_ragged_samples():
    """mixgrpo_logp_finalize works in chunks of 64 launches; 70 sampler steps on ragged samples (n = 5 * 63: scalar kernels, no
    trajectory seed) must still match the self-finalizing launches bit for bit."""
    from mixgrpo_b200 import rollout as R
    d = _dev()
    g = torch.Generator(device=d).manual_seed(77)
    B, S, C, N = 3, 5, 63, 70
    window = [10, 11, 40, 66]
    z0 = torch.randn(B, S, C, device=d, generator=g).bfloat16()
    vs = [torch.randn(B, S, C, device=d, generator=g) for _ in range(N)]
    nz = [torch.randn(B, S, C, device=d, generator=g) if i in window else None for i in range(N)]
    det = R.window_mask(N, window)
    sig = R.sigma_schedule(N, 3.0)
    a = R.rollout(lambda lt, s, i: vs[i], z0, sig, det, R.SamplerConfig(sampling_steps=N, defer_log_probs=False), noises=nz)
    b = R.rollout(lambda lt, s, i: vs[i], z0, sig, det, R.SamplerConfig(sampling_steps=N), noises=nz)
    assert torch.equal(a[2], b[2]) and a[3].shape == (B, N)
    assert torch.equal(torch.nan_to_num(a[3], nan=9.0), torch.nan_to_num(b[3], nan=9.0))
    c = R.rollout(lambda lt, s, i: vs[i], z0, sig, det, R.SamplerConfig(sampling_steps=N, ode_log_probs=False), noises=nz)
    assert torch.equal(c[3][:, window], a[3][:, window]) and torch.isnan(c[3][:, [0, 9, 12, 65, 69]]).all()
