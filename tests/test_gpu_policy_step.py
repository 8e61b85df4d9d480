"""Single-pass policy update (mixgrpo_policy_step, csrc/policy_kernels.cu): log-prob forward + clipped-ratio loss +
log-prob backward in one launch with the residuals kept on chip.  Checked against the oracle evaluated the reference's
way (per-sample autograd through step + loss, TR:536-585) and against the two-launch CUDA pair it replaces."""
import pytest
import torch

from oracle import grpo_oracle as GO
from oracle import sampling_oracle as O

pytestmark = pytest.mark.gpu

SIG = O.sd3_time_shift(3.0, torch.linspace(1, 0, 26))
ETA = 0.7
CLIP, AMAX, KLC, GA, T = 1e-4, 5.0, 0.01, 3, 4


def _dev():
    return torch.device("cuda:0")


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(autouse=True)
def short_timeout():
    from mixgrpo_b200 import _cabi
    old = _cabi.lib().mixgrpo_set_tuning(5, 3000)        # a scheduling bug must fail the test, not hang the box
    yield
    _cabi.lib().mixgrpo_set_tuning(5, old)


def _case(Bn, S, dtype, flow, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(Bn, S, 64, generator=g)
    v = torch.randn(Bn, S, 64, generator=g).to(dtype)
    eps = torch.randn(Bn, S, 64, generator=g).to(dtype if flow else torch.float32)
    idx = 8
    with torch.no_grad():
        if flow:
            xn, _, old_lp, _, _ = O.flow_step(v, x, ETA, SIG, idx, None, eps, False)
        else:
            xn, _, old_lp = O.dance_step(v, x, ETA, SIG, idx, None, eps, True, True)
    v_new = (v.float() + 0.02 * torch.randn(v.shape, generator=g)).to(dtype)
    adv = torch.randn(Bn, generator=g) * 3
    return x, v_new, xn, old_lp, adv, idx


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("flow", [True, False])
def test_single_pass_vs_oracle_autograd(dtype, flow):
    from mixgrpo_b200 import ops, rollout as R
    d = _dev()
    Bn = 6
    x, v_new, xn, old_lp, adv, idx = _case(Bn, 64, dtype, flow, 55)
    cfg = R.SamplerConfig(flow_grpo_sampling=flow, rounding="ref_cpu")
    rows = torch.zeros(Bn, 4, device=d)
    before = ops.launch_count
    _, new_lp, gv = R.policy_update(v_new.to(d), x.to(d), xn.to(d), old_lp.to(d), adv.to(d), SIG, idx, cfg, clip_range=CLIP, adv_clip_max=AMAX,
                                    kl_coeff=KLC, gradient_accumulation_steps=GA, num_train_timesteps=T, stats_rows=rows, single_pass=True)
    assert ops.launch_count - before == 1, "the single-pass kernel must cover this shape in ONE launch"
    vc = v_new.clone().requires_grad_(True)
    ref_rows, lps = torch.zeros(Bn, 4), []
    for i in range(Bn):                                          # the reference's per-sample loop
        if flow:
            lp = O.flow_step(vc[i:i + 1], x[i:i + 1], ETA, SIG, idx, xn[i:i + 1])[2]
        else:
            lp = O.dance_step(vc[i:i + 1], x[i:i + 1], ETA, SIG, idx, xn[i:i + 1], None, True, True)[2]
        out = GO.grpo_loss(lp, old_lp[i:i + 1], adv[i:i + 1], CLIP, AMAX, KLC, GA, T)
        out[0].backward()
        ref_rows[i] = torch.stack([o.detach() for o in out])
        lps.append(lp.detach())
    assert torch.allclose(new_lp.cpu(), torch.cat(lps), rtol=1e-5, atol=0)
    from test_gpu_parity import _assert_rows_at_gate
    _assert_rows_at_gate(rows.cpu(), ref_rows, new_lp.cpu(), torch.cat(lps), old_lp, GA * T)
    assert gv.dtype == dtype
    assert _rel(gv.float().cpu(), vc.grad.float()) < (2e-2 if dtype == torch.bfloat16 else 1e-4)


@pytest.mark.parametrize("Bn,S,dtype,flow", [
    (12, 4096, torch.bfloat16, True),        # BASELINE configs[1]: 1536 tiles over 888 co-resident CTAs
    (24, 4096, torch.bfloat16, True),        # configs[4]: group 24 at 1024^2 (3072 tiles, 3-4 per CTA)
    (24, 1024, torch.bfloat16, True),        # 512^2
    (12, 4096, torch.float32, True),
    (12, 2025, torch.bfloat16, True),        # 720^2 (the reference default): n % 2048 != 0 -> partial last tile per sample
    (4, 256, torch.bfloat16, False),         # configs[0] shape, dance family
    (3, 4096, torch.bfloat16, False),
    (1, 8, torch.float32, True),             # one partial tile
    (2000, 32, torch.bfloat16, True),        # many samples, one tile each: tiles of several samples per CTA
])
def test_single_pass_matches_the_two_launch_pair(Bn, S, dtype, flow):
    """Same bits as mixgrpo_policy_fwd + mixgrpo_policy_bwd: log-probs, stats rows and gradients."""
    from mixgrpo_b200 import ops, rollout as R
    d = _dev()
    g = torch.Generator(device=d).manual_seed(Bn * 7 + S)
    x = torch.randn(Bn, S, 64, device=d, generator=g)
    v = torch.randn(Bn, S, 64, device=d, generator=g).to(dtype)
    xn = x + 0.3 * torch.randn(Bn, S, 64, device=d, generator=g)
    old_lp = torch.randn(Bn, device=d, generator=g) * 0.01 - 1.0
    adv = torch.randn(Bn, device=d, generator=g) * 3
    cfg = R.SamplerConfig(flow_grpo_sampling=flow)
    kw = dict(clip_range=CLIP, adv_clip_max=AMAX, kl_coeff=KLC, gradient_accumulation_steps=GA, num_train_timesteps=T)
    rows1, rows2 = torch.zeros(Bn, 4, device=d), torch.zeros(Bn, 4, device=d)
    before = ops.launch_count
    _, lp1, g1 = R.policy_update(v, x, xn, old_lp, adv, SIG, 9, cfg, stats_rows=rows1, single_pass=True, **kw)
    assert ops.launch_count - before == 1
    _, lp2, g2 = R.policy_update(v, x, xn, old_lp, adv, SIG, 9, cfg, stats_rows=rows2, single_pass=False, **kw)
    torch.cuda.synchronize()
    assert torch.equal(lp1, lp2)
    assert torch.equal(rows1, rows2)
    assert torch.equal(g1, g2)
    # replays are bitwise reproducible and leave the workspace clean (accumulators re-zeroed, epochs advanced)
    _, lp3, g3 = R.policy_update(v, x, xn, old_lp, adv, SIG, 9, cfg, stats_rows=None, single_pass=True, **kw)
    assert torch.equal(lp1, lp3) and torch.equal(g1, g3)


def test_single_pass_in_a_cuda_graph_with_strided_trajectory_views():
    """Captured like bench.py does: inputs are views of the (B, N+1, S, 64) trajectory (batch stride != n)."""
    from mixgrpo_b200 import rollout as R
    d = _dev()
    Bn, S = 12, 1024
    g = torch.Generator(device=d).manual_seed(4)
    traj = torch.randn(Bn, 5, S, 64, device=d, generator=g)
    v = torch.randn(Bn, S, 64, device=d, generator=g).bfloat16()
    old_lp = torch.full((Bn,), -1.0, device=d)
    adv = torch.randn(Bn, device=d, generator=g)
    cfg = R.SamplerConfig()
    kw = dict(clip_range=CLIP, adv_clip_max=AMAX, kl_coeff=KLC, gradient_accumulation_steps=GA, num_train_timesteps=T)
    rows = torch.zeros(Bn, 4, device=d)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        R.policy_update(v, traj[:, 2], traj[:, 3], old_lp, adv, SIG, 9, cfg, stats_rows=rows, accumulate=False, single_pass=True, **kw)
    s.synchronize()
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph, stream=s):
        _, lp, gv = R.policy_update(v, traj[:, 2], traj[:, 3], old_lp, adv, SIG, 9, cfg, stats_rows=rows, accumulate=False, single_pass=True, **kw)
    outs = []
    for _ in range(3):
        gph.replay()
        torch.cuda.synchronize()
        outs.append((lp.clone(), gv.clone(), rows.clone()))
    _, lp_ref, gv_ref = R.policy_update(v, traj[:, 2], traj[:, 3], old_lp, adv, SIG, 9, cfg, stats_rows=None, single_pass=False, **kw)
    for lp_i, gv_i, rows_i in outs:
        assert torch.equal(lp_i, lp_ref) and torch.equal(gv_i, gv_ref) and torch.equal(rows_i, outs[0][2])
    assert torch.isfinite(gv_ref.float()).all()


def test_shapes_outside_the_on_chip_budget_fall_back_to_two_launches():
    from mixgrpo_b200 import ops, rollout as R
    d = _dev()
    cfg = R.SamplerConfig()
    kw = dict(clip_range=CLIP, adv_clip_max=AMAX, kl_coeff=KLC, gradient_accumulation_steps=GA, num_train_timesteps=T)
    for Bn, n_last in ((2, 63),):                                # ragged: n % 8 != 0
        x = torch.randn(Bn, 5, n_last, device=d)
        v = torch.randn(Bn, 5, n_last, device=d).bfloat16()
        before = ops.launch_count
        _, lp, gv = R.policy_update(v, x, x + 0.1, torch.zeros(Bn, device=d), torch.ones(Bn, device=d), SIG, 9, cfg, single_pass=True, **kw)
        assert ops.launch_count - before == 2 and torch.isfinite(lp).all()
    # 40 x 4096 x 64 fp32 residuals = 42 MB: more than the chip's shared memory
    Bn = 40
    x = torch.randn(Bn, 4096, 64, device=d)
    v = torch.randn(Bn, 4096, 64, device=d).bfloat16()
    before = ops.launch_count
    _, lp, gv = R.policy_update(v, x, x + 0.1, torch.zeros(Bn, device=d), torch.ones(Bn, device=d), SIG, 9, cfg, single_pass=True, **kw)
    assert ops.launch_count - before == 2 and torch.isfinite(lp).all() and torch.isfinite(gv.float()).all()
