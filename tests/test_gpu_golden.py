"""-m gpu: the CUDA path against the committed golden vectors — outputs of the REFERENCE ITSELF:
  tests/golden/reference_ops_cpu.npz      reference run on CPU tensors   -> rounding="ref_cpu" must be bit-exact
  tests/golden/cuda_reference_b200.npz    reference run on B200 tensors  -> rounding="ref_cuda" (the default) bit-exact
plus full rollouts (drop-in run_sample_step and the batched native driver) against the pinned oracle for every
schedule family: MixGRPO, DanceGRPO, DPM "all", MixGRPO-Flash "post" (midpoint / heun / dance)."""
import itertools
import sys
import types
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import sampling_oracle as O

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
ETA, SHIFT, N = 0.7, 3.0, 25
SIG = O.sd3_time_shift(SHIFT, torch.linspace(1, 0, N + 1))
DEV = torch.device("cuda:0")


def _t(a, dtype=torch.float32):
    return torch.from_numpy(np.asarray(a)).to(dtype)


def _eq(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return torch.equal(torch.isnan(a), torch.isnan(b)) and torch.equal(torch.nan_to_num(a, posinf=1e38, neginf=-1e38),
                                                                         torch.nan_to_num(b, posinf=1e38, neginf=-1e38))


def _close(a, b, rtol):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    ok = torch.isfinite(b)
    return torch.equal(torch.isfinite(a), ok) and torch.allclose(a[ok], b[ok], rtol=rtol, atol=0)


@pytest.mark.parametrize("dn,dtype", [("bf16", torch.bfloat16), ("f32", torch.float32)])
def test_flow_and_dance_vs_reference_cpu_golden(dn, dtype):
    from mixgrpo_b200 import sampling_utils as su
    z = np.load(GOLD / "reference_ops_cpu.npz")
    for idx, det in itertools.product((0, 1, 3, 12, 23, 24), (0, 1)):
        k = f"flow/{dn}/i{idx}/det{det}"
        x, v, e = _t(z[f"{k}/x"]).to(DEV), _t(z[f"{k}/v"], dtype).to(DEV), _t(z[f"{k}/eps"], dtype).to(DEV)
        out = su.flow_grpo_step(v, x, ETA, SIG, idx, None, determistic=bool(det), noise=e, rounding="ref_cpu")
        assert _eq(out[0], _t(z[f"{k}/prev"])) and _eq(out[1], _t(z[f"{k}/x0"])) and _eq(out[3], _t(z[f"{k}/mean"])), k
        assert _close(out[2], _t(z[f"{k}/logp"]), 1e-5), k
        assert out[4].item() == float(z[f"{k}/scale"])
        if not det:
            vg = v.clone().requires_grad_(True)
            lp = su.flow_grpo_step(vg, x, ETA, SIG, idx, _t(z[f"{k}/xn_train"]).to(DEV), rounding="ref_cpu")[2]
            (lp * torch.tensor([0.7, -1.3], device=DEV)).sum().backward()
            assert _close(lp, _t(z[f"{k}/train_logp"]), 1e-5)
            ref_g = _t(z[f"{k}/train_grad"])
            err = (vg.grad.float().cpu() - ref_g).norm() / ref_g.norm()
            assert err < (1e-2 if dtype == torch.bfloat16 else 1e-5), (k, err)
    for idx, sde in itertools.product((0, 3, 12, 24), (1, 0)):
        k = f"dance/{dn}/i{idx}/sde{sde}"
        x, v = _t(z[f"{k}/x"]).to(DEV), _t(z[f"{k}/v"], dtype).to(DEV)
        out = su.dance_grpo_step(v, x, ETA, SIG, idx, None, True, bool(sde), noise=_t(z[f"{k}/noise"]).to(DEV), rounding="ref_cpu")
        assert _eq(out[0], _t(z[f"{k}/prev"])) and _eq(out[1], _t(z[f"{k}/x0"])), k
        assert torch.allclose(out[2].cpu(), _t(z[f"{k}/logp"]), rtol=1e-4, atol=1e-12)
        vg = v.clone().requires_grad_(bool(sde))        # sde_solver=False has no backward kernel (never trained, TR:159-168): it raises
        lp = su.dance_grpo_step(vg, x, ETA, SIG, idx, _t(z[f"{k}/xn_train"]).to(DEV), True, bool(sde), rounding="ref_cpu")[2]
        assert _close(lp, _t(z[f"{k}/train_logp"]), 1e-4)
        if sde:
            lp.sum().backward()
            ref_g = _t(z[f"{k}/train_grad"])
            err = (vg.grad.float().cpu() - ref_g).norm() / ref_g.norm()
            assert err < (1e-2 if dtype == torch.bfloat16 else 1e-5), (k, err)


@pytest.mark.parametrize("dn,dtype", [("bf16", torch.bfloat16), ("f32", torch.float32)])
def test_dpm_vs_reference_cpu_golden(dn, dtype):
    from mixgrpo_b200 import sampling_utils as su
    z = np.load(GOLD / "reference_ops_cpu.npz")
    n = 0
    for algo, stype, order, sde, idx in itertools.product(("dpmsolver++", "dpmsolver"), ("midpoint", "heun"), (1, 2, 3), (0, 1), (0, 1, 2, 13, 24)):
        k = f"dpm/{dn}/{algo}/{stype}/o{order}/sde{sde}/i{idx}"
        if f"{k}/x" not in z:
            continue
        args = types.SimpleNamespace(dpm_algorithm_type=algo, dpm_solver_type=stype, dpm_solver_order=order)
        st = su.DPMState(order=order)
        st.model_outputs = [(_t(z[f"{k}/hist{j}"]).to(DEV) if f"{k}/hist{j}" in z else None) for j in range(order)]
        st.lower_order_nums = min(idx, order)
        out = su.dpm_step(args, _t(z[f"{k}/v"], dtype).to(DEV), _t(z[f"{k}/x"]).to(DEV), idx, SIG[:-1], SIG, dpm_state=st,
                          variance_noise=_t(z[f"{k}/eps"]).to(DEV), sde_solver=bool(sde), rounding="ref_cpu")
        assert _eq(out[1], _t(z[f"{k}/x0"])), k
        ref = _t(z[f"{k}/prev"])
        fin = torch.isfinite(ref)
        assert torch.equal(torch.isfinite(out[0].cpu()), fin), k
        if fin.all():
            assert (out[0].cpu() - ref).norm() / ref.norm() < 1e-5, k                     # bar: 1e-5 (fp32) / 1e-3 (bf16)
        if sde:
            assert _close(out[2], _t(z[f"{k}/logp"]), 1e-4), k
        n += 1
    assert n >= 80


def test_default_rounding_vs_reference_on_b200_golden():
    """Golden vectors recorded by running the reference on a B200 (tools/probe_cuda_rounding.py): the default
    rounding mode must reproduce them bit-for-bit (log-prob to 1e-5; bar 1e-4)."""
    from mixgrpo_b200 import sampling_utils as su
    path = GOLD / "cuda_reference_b200.npz"
    if not path.exists():
        pytest.skip("CUDA golden vectors not generated yet")
    z = np.load(path)
    keys = sorted({k.rsplit("/", 1)[0] for k in z.files})
    n = 0
    for k in keys:
        parts = k.split("_")
        dtype = torch.bfloat16 if parts[1] == "bf16" else torch.float32
        if parts[0] == "flow":
            idx, det = int(parts[2][1:]), bool(int(parts[3][3:]))
            x, v, e = _t(z[f"{k}/x"]).to(DEV), _t(z[f"{k}/v"], dtype).to(DEV), _t(z[f"{k}/eps"], dtype).to(DEV)
            out = su.flow_grpo_step(v, x, ETA, SIG, idx, None, determistic=det, noise=e)
            assert _eq(out[0], _t(z[f"{k}/prev"])) and _eq(out[1], _t(z[f"{k}/x0"])) and _eq(out[3], _t(z[f"{k}/mean"])), k
            assert _close(out[2], _t(z[f"{k}/logp"]), 1e-5), k
            if f"{k}/train_grad" in z.files:
                vg = v.clone().requires_grad_(True)
                lp = su.flow_grpo_step(vg, x, ETA, SIG, idx, _t(z[f"{k}/xn_train"]).to(DEV))[2]
                lp.sum().backward()
                assert _close(lp, _t(z[f"{k}/train_logp"]), 1e-5), k
                assert _eq(vg.grad, _t(z[f"{k}/train_grad"])), k          # gradient is bit-exact, bf16 included
            n += 1
        elif parts[0] == "dance":
            idx, sde = int(parts[2][1:]), bool(int(parts[3][3:]))
            out = su.dance_grpo_step(_t(z[f"{k}/v"], dtype).to(DEV), _t(z[f"{k}/x"]).to(DEV), ETA, SIG, idx, _t(z[f"{k}/xn"]).to(DEV), True, sde)
            assert _eq(out[1], _t(z[f"{k}/x0"])) and _close(out[2], _t(z[f"{k}/logp"]), 1e-4), k
            n += 1
    assert n >= 40


# ------------------------------------------------------------------------------ full rollouts
class ExactStandIn(torch.nn.Module):
    """Stand-in DiT built from IEEE-exact ops only (mul/add/roll), so CPU and CUDA produce identical bf16 outputs."""

    def forward(self, hidden_states, encoder_hidden_states, timestep, guidance, txt_ids, pooled_projections, img_ids,
                joint_attention_kwargs, return_dict):
        z = hidden_states.float()
        # timestep = int(sigma*1000)/1000 is a division torch evaluates as x*(1/1000) on CUDA and x/1000 on CPU (1 ulp
        # apart): recover the integer and rescale by a power of two so both devices agree bit-for-bit
        t = (torch.round(timestep.float() * 1000.0) * 0.0009765625).view(-1, 1, 1)
        return ((z.roll(1, dims=-1) * 0.75 + z * 0.125 + t * 0.5).to(torch.bfloat16),)


def _cases():
    sys.path.insert(0, str(Path(__file__).parent.parent / "tools"))
    from make_golden import ROLLOUT_CASES
    return ROLLOUT_CASES


@pytest.mark.parametrize("name", ["mixgrpo_w4_8", "mixgrpo_w0_drop", "dance_w5", "dpm_all_o2", "flash_mid_04", "flash_heun_02", "flash_dance_04"])
def test_rollout_dropin_and_native_vs_oracle(name):
    from mixgrpo_b200 import rollout as R
    from mixgrpo_b200 import sampling_utils as su
    args, window = _cases()[name]
    det = [i not in window for i in range(N)]
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    Br, Sr = 3, 20
    z0 = torch.randn(Br, Sr, 64, generator=g).bfloat16()
    all_sde = args.dpm_algorithm_type != "null" and args.dpm_apply_strategy == "all"
    ndt = torch.float32 if (all_sde or not args.flow_grpo_sampling) else torch.bfloat16
    noises = [torch.randn(Br, Sr, 64, generator=g).to(ndt) for _ in range(N)]
    model = ExactStandIn()

    def cpu_model(zz, s, i):
        return model(zz, None, torch.full([Br], int(s * 1000)) / 1000, None, None, None, None, None, False)[0]

    ref = O.rollout(cpu_model, z0, SIG, det, noises, eta=args.eta, shift=args.shift, flow_grpo_sampling=args.flow_grpo_sampling,
                    dpm_algorithm_type=args.dpm_algorithm_type, dpm_apply_strategy=args.dpm_apply_strategy,
                    dpm_post_compress_ratio=args.dpm_post_compress_ratio, dpm_solver_order=args.dpm_solver_order,
                    dpm_solver_type=args.dpm_solver_type, drop_last_sample=args.drop_last_sample)
    dn = [t.to(DEV) for t in noises]
    # (1) drop-in run_sample_step with the reference's positional signature
    out = su.run_sample_step(args, z0.to(DEV), range(N), SIG.to(DEV), model.to(DEV), torch.zeros(Br, 4, 8, device=DEV),
                             torch.zeros(Br, 8, device=DEV), torch.zeros(Br, 3, device=DEV), torch.zeros(Sr, 3, device=DEV), True, det,
                             noises=dn, rounding="ref_cpu")
    # (2) batched native driver
    cfg = R.SamplerConfig(sampling_steps=N, eta=args.eta, shift=args.shift, flow_grpo_sampling=args.flow_grpo_sampling,
                          dpm_algorithm_type=args.dpm_algorithm_type, dpm_apply_strategy=args.dpm_apply_strategy,
                          dpm_post_compress_ratio=args.dpm_post_compress_ratio, dpm_solver_order=args.dpm_solver_order,
                          dpm_solver_type=args.dpm_solver_type, drop_last_sample=args.drop_last_sample, rounding="ref_cpu")

    def gpu_model(zz, s, i):
        return model(zz, None, torch.full([Br], int(s * 1000), device=DEV) / 1000, None, None, None, None, None, False)[0]

    nat = R.rollout(gpu_model, z0.to(DEV), SIG, det, cfg, noises=dn)
    exact = args.dpm_algorithm_type == "null"          # dpm coefficients use host exp/log: 1e-5 instead of bit-exact
    for got in (out, nat[:4]):
        assert got[2].shape == ref[2].shape and got[3].shape == ref[3].shape
        if exact:
            assert _eq(got[0], ref[0]) and _eq(got[1], ref[1]) and _eq(got[2], ref[2]), name
        else:
            for a, b in zip(got[:3], ref[:3]):
                assert (a.float().cpu() - b.float()).norm() / b.float().norm() < 1e-5, name
        lp, rlp = got[3].cpu(), ref[3]
        sde_cols = [i for i in range(rlp.shape[1]) if i < N and not det[i]]
        assert torch.allclose(lp[:, sde_cols], rlp[:, sde_cols], rtol=1e-4, atol=0), name      # the trained log-probs
        fin = torch.isfinite(rlp)
        assert torch.allclose(lp[fin], rlp[fin], rtol=2e-3, atol=1e-3), name                    # ODE-step values (unused)
    assert nat[4].numel() == ref[2].shape[1]


@pytest.mark.parametrize("name", ["mixgrpo_w4_8", "flash_mid_04"])
def test_full_size_rollouts_vs_reference_ops_on_device(name):
    """BASELINE configs[1] / configs[3] at FULL size — group 12, (4096, 64) latents, 25 steps (MixGRPO window 4; Flash with
    the compressed DPM-Solver++ tail) — through the batched driver, against the reference's op sequence rolled out on the
    same B200 (oracle.rollout on CUDA tensors, default CUDA rounding).  Trajectory bit-exact for MixGRPO."""
    from mixgrpo_b200 import rollout as R
    args, window = _cases()[name]
    det = [i not in window for i in range(N)]
    g = torch.Generator(device=DEV).manual_seed(11)
    Br, Sr = 12, 4096
    z0 = torch.randn(Br, Sr, 64, device=DEV, generator=g).bfloat16()
    noises = [torch.randn(Br, Sr, 64, device=DEV, generator=g).bfloat16() for i in range(N)]   # the reference draws noise on ODE steps too (SU:188-195)
    model = ExactStandIn().to(DEV)

    def gpu_model(zz, s, i):
        return model(zz, None, torch.full([Br], int(float(s) * 1000), device=DEV) / 1000, None, None, None, None, None, False)[0]

    ref = O.rollout(gpu_model, z0, SIG.to(DEV), det, noises, eta=args.eta, shift=args.shift, flow_grpo_sampling=args.flow_grpo_sampling,
                    dpm_algorithm_type=args.dpm_algorithm_type, dpm_apply_strategy=args.dpm_apply_strategy,
                    dpm_post_compress_ratio=args.dpm_post_compress_ratio, dpm_solver_order=args.dpm_solver_order,
                    dpm_solver_type=args.dpm_solver_type, drop_last_sample=args.drop_last_sample)
    cfg = R.SamplerConfig(sampling_steps=N, eta=args.eta, shift=args.shift, flow_grpo_sampling=args.flow_grpo_sampling,
                          dpm_algorithm_type=args.dpm_algorithm_type, dpm_apply_strategy=args.dpm_apply_strategy,
                          dpm_post_compress_ratio=args.dpm_post_compress_ratio, dpm_solver_order=args.dpm_solver_order,
                          dpm_solver_type=args.dpm_solver_type, drop_last_sample=args.drop_last_sample)
    nat = R.rollout(gpu_model, z0, SIG, det, cfg, noises=noises)
    assert nat[2].shape == ref[2].shape and nat[3].shape == ref[3].shape
    if args.dpm_algorithm_type == "null":
        assert torch.equal(nat[2], ref[2]) and torch.equal(nat[0], ref[0].float()), name
    else:
        assert ((nat[2] - ref[2]).norm() / ref[2].norm()).item() < 1e-5, name
    sde_cols = [i for i in range(ref[3].shape[1]) if i < N and not det[i]]
    assert sde_cols
    assert torch.allclose(nat[3][:, sde_cols], ref[3][:, sde_cols], rtol=1e-5, atol=0), name
