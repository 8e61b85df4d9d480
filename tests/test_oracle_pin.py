"""Pins the oracle (oracle/*.py) to the reference.

* ``test_*_vs_live_reference``: execute the UNMODIFIED reference functions where they lie (oracle/ref_loader.py,
  oracle/ref_extract.py) and require bit-identical tensors.  Skipped on boxes without the reference tree.
* ``test_*_vs_golden``: the same comparisons against tests/golden/*.npz — outputs of the reference itself
  (tools/make_golden.py), so the oracle stays pinned on the GPU box where the tree does not exist.
CPU only; no CUDA library involved.
"""
import itertools
import types
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import grpo_oracle as GO
from oracle import ref_extract, ref_loader
from oracle import sampling_oracle as O

GOLD = Path(__file__).parent / "golden"
ETA, SHIFT, N = 0.7, 3.0, 25
SIG = O.sd3_time_shift(SHIFT, torch.linspace(1, 0, N + 1))


def _t(a, dtype=torch.float32):
    return torch.from_numpy(np.asarray(a)).to(dtype)


def _eq(a, b):
    a, b = a.detach().float(), b.detach().float()
    return torch.equal(torch.isnan(a), torch.isnan(b)) and torch.equal(torch.nan_to_num(a, posinf=1e38, neginf=-1e38),
                                                                         torch.nan_to_num(b, posinf=1e38, neginf=-1e38))


def _mk(g, dtype, shape=(3, 16, 64)):
    x = torch.randn(*shape, generator=g)
    v = torch.randn(*shape, generator=g).to(dtype)
    e = torch.randn(*shape, generator=g).to(dtype)
    xn = torch.randn(*shape, generator=g)
    return x, v, e, xn


# ------------------------------------------------------------------------------ live reference
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_flow_and_dance_vs_live_reference(ref_su, dtype):
    g = torch.Generator().manual_seed(0)
    for idx in (0, 1, 3, 12, 23, 24):
        for det in (False, True):
            x, v, e, xn = _mk(g, dtype)
            ref_loader.NOISE_QUEUE[:] = [e]
            r = ref_su.flow_grpo_step(v, x, ETA, SIG, idx, None, determistic=det)
            o = O.flow_step(v, x, ETA, SIG, idx, None, e, det)
            assert all(_eq(a, b) for a, b in zip(r, o)), ("flow rollout", idx, det)
            r = ref_su.flow_grpo_step(v, x, ETA, SIG, idx, xn)
            o = O.flow_step(v, x, ETA, SIG, idx, xn)
            assert all(_eq(a, b) for a, b in zip(r, o)), ("flow train", idx)
    for idx in (0, 3, 12, 24):
        for sde in (True, False):
            x, v, e, xn = _mk(g, dtype)
            assert all(_eq(a, b) for a, b in zip(ref_su.dance_grpo_step(v, x, ETA, SIG, idx, xn, True, sde),
                                                 O.dance_step(v, x, ETA, SIG, idx, xn, None, True, sde)))
            torch.manual_seed(5)
            r = ref_su.dance_grpo_step(v, x, ETA, SIG, idx, None, True, sde)
            torch.manual_seed(5)
            nz = torch.randn_like(x)
            assert all(_eq(a, b) for a, b in zip(r, O.dance_step(v, x, ETA, SIG, idx, None, nz, True, sde)))
            assert all(_eq(a, b) for a, b in zip(ref_su.dance_grpo_step(v, x, ETA, SIG, idx, None, False, sde),
                                                 O.dance_step(v, x, ETA, SIG, idx, None, None, False, sde)))


@pytest.mark.parametrize("algo,stype,order", list(itertools.product(["dpmsolver++", "dpmsolver"], ["midpoint", "heun"], [1, 2, 3])))
def test_dpm_vs_live_reference(ref_su, algo, stype, order):
    g = torch.Generator().manual_seed(order)
    for dtype in (torch.bfloat16, torch.float32):
        for sde in (False, True):
            args = types.SimpleNamespace(dpm_algorithm_type=algo, dpm_solver_type=stype, dpm_solver_order=order)
            rs, oh = ref_su.DPMState(order=order), O.History(order)
            for idx in range(N):
                x, v, e, _ = _mk(g, dtype, (2, 4, 64))
                e = e.float()
                rerr = oerr = None
                try:
                    r = ref_su.dpm_step(args, v, x, idx, SIG[:-1], SIG, dpm_state=rs, variance_noise=e, sde_solver=sde)
                except Exception as ex:  # noqa: BLE001
                    rerr = type(ex).__name__
                try:
                    o = O.dpm_step(v, x, idx, N, SIG, algo=algo, solver_order=order, solver_type=stype, history=oh, noise=e, sde_solver=sde)
                except Exception as ex:  # noqa: BLE001
                    oerr = type(ex).__name__
                assert rerr == oerr, (algo, stype, order, idx, rerr, oerr)
                if rerr:
                    break
                assert all(_eq(a, b) for a, b in zip(r, o)), (algo, stype, order, dtype, idx, sde)


def test_rollout_vs_live_reference(ref_su):
    import sys
    sys.path.insert(0, str(Path(__file__).parent.parent / "tools"))
    from make_golden import ROLLOUT_CASES, StandIn
    model = StandIn()
    g = torch.Generator().manual_seed(3)
    Br, Sr = 2, 4
    for name, (args, window) in ROLLOUT_CASES.items():
        if not args.flow_grpo_sampling:
            continue                                   # dance draws from the global generator; covered by the golden test
        det = [i not in window for i in range(N)]
        z = torch.randn(Br, Sr, 64, generator=g).bfloat16()
        all_sde = args.dpm_algorithm_type != "null" and args.dpm_apply_strategy == "all"
        noises = [torch.randn(Br, Sr, 64, generator=g).to(torch.float32 if all_sde else torch.bfloat16) for _ in range(N)]
        ref_loader.NOISE_QUEUE[:] = [noises[i] for i in range(N) if not det[i]] if all_sde else list(noises)
        with pytest.warns(UserWarning):
            r = ref_su.run_sample_step(args, z, range(N), SIG, model, torch.zeros(Br, 4, 8), torch.zeros(Br, 8), torch.zeros(Br, 3),
                                       torch.zeros(Sr, 3), True, det)
        ref_loader.NOISE_QUEUE[:] = []
        o = O.rollout(lambda zz, s, i: model(zz, None, torch.full([Br], int(s * 1000)) / 1000, None, None, None, None, None, False)[0],
                      z, SIG, det, noises, eta=args.eta, shift=args.shift, flow_grpo_sampling=True,
                      dpm_algorithm_type=args.dpm_algorithm_type, dpm_apply_strategy=args.dpm_apply_strategy,
                      dpm_post_compress_ratio=args.dpm_post_compress_ratio, dpm_solver_order=args.dpm_solver_order,
                      dpm_solver_type=args.dpm_solver_type, drop_last_sample=args.drop_last_sample)
        assert all(_eq(a, b) for a, b in zip(r, o)), name


def test_grpo_oracle_source_pin():
    """oracle/grpo_oracle.py against the reference's own inline statements (TR:440-501, TR:560-583) run via ast."""
    if ref_extract.reference_advantages(torch.zeros(4), None, use_group=True, num_generations=4, trimmed_ratio=0.0,
                                        multi_reward_mix="reward_aggr", reward_weights=None) is None:
        pytest.skip("train_grpo_flux.py not present on this box")
    g = torch.Generator().manual_seed(11)
    for trial in range(5):
        G = (4, 12, 24, 7, 12)[trial]
        r3 = {"a": torch.randn(2 * G, generator=g), "b": torch.randn(2 * G, generator=g) * 0.01 + 0.3, "c": torch.randn(2 * G, generator=g)}
        r3["c"][G:] = 1.5
        w = {"a": 1.0, "b": 0.25, "c": 3.0}
        for ratio in (0.0, 0.1, 0.34, 0.99):
            ref = ref_extract.reference_advantages({k: t.clone() for k, t in r3.items()}, None, use_group=True, num_generations=G,
                                                   trimmed_ratio=ratio, multi_reward_mix="advantage_aggr", reward_weights=w)
            assert _eq(ref, GO.group_advantages(r3, G, w, trimmed_ratio=ratio))
            ref = ref_extract.reference_advantages(r3["a"].clone(), None, use_group=True, num_generations=G, trimmed_ratio=ratio,
                                                   multi_reward_mix="reward_aggr", reward_weights=None)
            assert _eq(ref, GO.group_advantages(r3["a"], G, trimmed_ratio=ratio))
        gathered = torch.randn(4 * G, generator=g)
        ref = ref_extract.reference_advantages(gathered[:G].clone(), gathered, use_group=False, num_generations=G, trimmed_ratio=0.0,
                                               multi_reward_mix="reward_aggr", reward_weights=None)
        assert _eq(ref, GO.group_advantages(gathered[:G], G, use_group=False, gathered=gathered))
        for Bn in (1, G):
            old = -1 + 0.1 * torch.randn(Bn, generator=g)
            new = old + 2e-4 * torch.randn(Bn, generator=g)
            adv = 3 * torch.randn(Bn, generator=g)
            for kl in (0.0, 0.02):
                a, b = new.clone().requires_grad_(True), new.clone().requires_grad_(True)
                ref = ref_extract.reference_loss(a, old, adv, clip_range=1e-4, adv_clip_max=5.0, kl_coeff=kl,
                                                 gradient_accumulation_steps=3, n_train_timesteps=4)
                mine = GO.grpo_loss(b, old, adv, 1e-4, 5.0, kl, 3, 4)
                assert all(_eq(x, y) for x, y in zip(ref, mine))
                ref[0].backward()
                mine[0].backward()
                assert _eq(a.grad, b.grad)


# ------------------------------------------------------------------------------ golden vectors (always run)
def _ops():
    return np.load(GOLD / "reference_ops_cpu.npz")


def test_sigmas_vs_golden():
    assert torch.equal(_t(_ops()["sigmas"]), SIG)


@pytest.mark.parametrize("dn,dtype", [("bf16", torch.bfloat16), ("f32", torch.float32)])
def test_flow_dance_oracle_vs_golden(dn, dtype):
    z = _ops()
    for idx in (0, 1, 3, 12, 23, 24):
        for det in (0, 1):
            k = f"flow/{dn}/i{idx}/det{det}"
            x, v, e = _t(z[f"{k}/x"]), _t(z[f"{k}/v"], dtype), _t(z[f"{k}/eps"], dtype)
            o = O.flow_step(v, x, ETA, SIG, idx, None, e, bool(det))
            for nm, t in zip(("prev", "x0", "logp", "mean", "scale"), o):
                assert _eq(t, _t(z[f"{k}/{nm}"])), (k, nm)
            if not det:
                vg = v.clone().requires_grad_(True)
                lp = O.flow_step(vg, x, ETA, SIG, idx, _t(z[f"{k}/xn_train"]))[2]
                (lp * torch.tensor([0.7, -1.3])).sum().backward()
                assert _eq(lp, _t(z[f"{k}/train_logp"])) and _eq(vg.grad, _t(z[f"{k}/train_grad"]))
    for idx in (0, 3, 12, 24):
        for sde in (1, 0):
            k = f"dance/{dn}/i{idx}/sde{sde}"
            x, v = _t(z[f"{k}/x"]), _t(z[f"{k}/v"], dtype)
            o = O.dance_step(v, x, ETA, SIG, idx, None, _t(z[f"{k}/noise"]), True, bool(sde))
            assert _eq(o[0], _t(z[f"{k}/prev"])) and _eq(o[1], _t(z[f"{k}/x0"])) and _eq(o[2], _t(z[f"{k}/logp"]))
            vg = v.clone().requires_grad_(True)
            lp = O.dance_step(vg, x, ETA, SIG, idx, _t(z[f"{k}/xn_train"]), None, True, bool(sde))[2]
            lp.sum().backward()
            assert _eq(lp, _t(z[f"{k}/train_logp"])) and _eq(vg.grad, _t(z[f"{k}/train_grad"]))


@pytest.mark.parametrize("dn,dtype", [("bf16", torch.bfloat16), ("f32", torch.float32)])
def test_dpm_oracle_vs_golden(dn, dtype):
    z = _ops()
    n = 0
    for algo, stype, order, sde, idx in itertools.product(("dpmsolver++", "dpmsolver"), ("midpoint", "heun"), (1, 2, 3), (0, 1), (0, 1, 2, 13, 24)):
        k = f"dpm/{dn}/{algo}/{stype}/o{order}/sde{sde}/i{idx}"
        if f"{k}/x" not in z:
            continue
        hist = O.History(order)
        hs = [(_t(z[f"{k}/hist{j}"]) if f"{k}/hist{j}" in z else None) for j in range(order)]
        hist.model_outputs = hs
        hist.lower_order_nums = min(idx, order)
        o = O.dpm_step(_t(z[f"{k}/v"], dtype), _t(z[f"{k}/x"]), idx, N, SIG, algo=algo, solver_order=order, solver_type=stype,
                       history=hist, noise=_t(z[f"{k}/eps"]), sde_solver=bool(sde))
        for nm, t in zip(("prev", "x0", "logp"), o):
            assert _eq(t, _t(z[f"{k}/{nm}"])), (k, nm)
        n += 1
    assert n >= 80


def test_rollout_oracle_vs_golden():
    import sys
    sys.path.insert(0, str(Path(__file__).parent.parent / "tools"))
    from make_golden import ROLLOUT_CASES, StandIn
    z = np.load(GOLD / "reference_rollout_cpu.npz")
    model = StandIn()
    for name, (args, window) in ROLLOUT_CASES.items():
        det = [i not in window for i in range(N)]
        z0 = _t(z[f"{name}/z"], torch.bfloat16)
        Br = z0.shape[0]
        all_sde = args.dpm_algorithm_type != "null" and args.dpm_apply_strategy == "all"
        ndt = torch.float32 if (all_sde or not args.flow_grpo_sampling) else torch.bfloat16
        noises = [_t(a, ndt) for a in z[f"{name}/noises"]]
        o = O.rollout(lambda zz, s, i: model(zz, None, torch.full([Br], int(s * 1000)) / 1000, None, None, None, None, None, False)[0],
                      z0, SIG, det, noises, eta=args.eta, shift=args.shift, flow_grpo_sampling=args.flow_grpo_sampling,
                      dpm_algorithm_type=args.dpm_algorithm_type, dpm_apply_strategy=args.dpm_apply_strategy,
                      dpm_post_compress_ratio=args.dpm_post_compress_ratio, dpm_solver_order=args.dpm_solver_order,
                      dpm_solver_type=args.dpm_solver_type, drop_last_sample=args.drop_last_sample)
        for nm, t in zip(("z_out", "latents", "all_latents", "all_log_probs"), o):
            assert _eq(t, _t(z[f"{name}/{nm}"])), (name, nm)


def test_grpo_oracle_vs_golden():
    z = np.load(GOLD / "reference_grpo_cpu.npz")
    r3 = {k: _t(z[f"adv/rewards/{k}"]) for k in ("hps", "pick", "ir")}
    w3 = {"hps": 1.0, "pick": 0.5, "ir": 2.0}
    for ratio in (0.0, 0.2, 0.5):
        assert _eq(GO.group_advantages(r3, 12, w3, trimmed_ratio=ratio), _t(z[f"adv/advantage_aggr/trim{ratio}"]))
        assert _eq(GO.group_advantages(r3["hps"], 12, trimmed_ratio=ratio), _t(z[f"adv/reward_aggr/trim{ratio}"]))
    gathered = _t(z["adv/gathered"])
    assert _eq(GO.group_advantages(gathered[12:24], 12, use_group=False, gathered=gathered), _t(z["adv/nogroup"]))
    for Bn in (1, 12):
        for kl in (0.0, 0.01):
            k = f"loss/B{Bn}/kl{kl}"
            new = _t(z[f"{k}/new"]).requires_grad_(True)
            out = GO.grpo_loss(new, _t(z[f"{k}/old"]), _t(z[f"{k}/adv"]), 1e-4, 5.0, kl, 3, 4)
            out[0].backward()
            for nm, t in zip(("loss", "policy", "kl", "clip_frac", "grad"), (*out, new.grad)):
                assert _eq(t, _t(z[f"{k}/{nm}"])), (k, nm)


def test_survey_sanity_values():
    """values recorded from the reference during the survey."""
    for i, s in ((0, 0.7000), (1, 0.7145), (12, 0.2186), (23, 0.1107), (24, 0.0825)):
        x = torch.zeros(1, 4, 64)
        sc = O.flow_step(x.bfloat16(), x, ETA, SIG, i, x)[4]
        assert abs(sc.item() - s) < 5e-5
    a = GO.group_advantages(torch.tensor([0.1, 0.4, 0.2, 0.9]), 4)
    assert torch.allclose(a, torch.tensor([-0.84293, 8.4e-08, -0.56195, 1.40488]), atol=2e-5)
    assert torch.equal(GO.group_advantages(torch.full((4,), 0.3), 4), torch.zeros(4))
    x = torch.randn(2, 8, 64)
    assert torch.equal(O.dance_step(x.bfloat16(), x, ETA, SIG, 5, None, None, True, False)[2], torch.zeros(2))
