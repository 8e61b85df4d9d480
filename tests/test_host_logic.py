"""Host-side logic of the product package (no GPU): per-step coefficient tables vs the oracle's zero-dim
scalars, schedule construction, window masks, sample bookkeeping, reward stacking."""
import random
import types

import pytest
import torch

from mixgrpo_b200 import coefs, grpo, rollout as R
from mixgrpo_b200 import sampling_utils as su
from oracle import grpo_oracle as GO
from oracle import sampling_oracle as O

SIG = O.sd3_time_shift(3.0, torch.linspace(1, 0, 26))


def _bf(x):
    return torch.tensor(x).bfloat16().float().item()


@pytest.mark.parametrize("index", [0, 1, 5, 12, 23, 24])
def test_flow_coefficients_match_oracle_scalars(index):
    x = torch.zeros(1, 2, 64)
    _, _, _, _, scale = O.flow_step(x, x, 0.7, SIG, index, x)
    k, sc = coefs.flow(SIG, index, 0.7, "fp32", False)
    assert sc == scale.item()
    s, sn = SIG[index], SIG[index + 1]
    dt = sn - s
    std = torch.sqrt(s / (1 - torch.where(s == 1, SIG[1].item(), s))) * 0.7
    assert k.c[0] == s.item() and k.c[3] == dt.item() and k.c[5] == dt.item()
    assert k.c[1] == (1 + std ** 2 / (2 * s) * dt).item()
    assert k.c[2] == (1 + std ** 2 * (1 - s) / (2 * s)).item()
    assert k.two_var == (2 * (scale ** 2)).item()
    assert k.log_scale == torch.log(scale).item()
    # rounding modes only touch the factors torch would cast to bf16
    kc, _ = coefs.flow(SIG, index, 0.7, "ref_cpu", True)
    kg, _ = coefs.flow(SIG, index, 0.7, "ref_cuda", True)
    assert (kc.c[0], kc.c[4], kc.c[5]) == (_bf(k.c[0]), _bf(k.c[4]), _bf(k.c[5]))
    assert (kc.c[1], kc.c[2], kc.c[3]) == (k.c[1], k.c[2], k.c[3])
    assert (kg.c[2], kg.c[3]) == (_bf(k.c[2]), _bf(k.c[3])) and kg.c[1] == k.c[1]
    kf, _ = coefs.flow(SIG, index, 0.7, "ref_cuda", False)          # fp32 model output: nothing is rounded
    assert list(kf.c) == list(k.c)


def test_schedule_cache_is_value_safe():
    a = O.sd3_time_shift(3.0, torch.linspace(1, 0, 26))
    k1, _ = coefs.flow(a, 3, 0.7, "fp32", False)
    a[4] = 0.5                                    # in-place edit bumps _version -> new host copy, new coefficients
    k2, _ = coefs.flow(a, 3, 0.7, "fp32", False)
    assert k1.c[3] != k2.c[3]
    b = O.sd3_time_shift(2.0, torch.linspace(1, 0, 26))
    k3, _ = coefs.flow(b, 3, 0.7, "fp32", False)
    assert k3.c[0] == b[3].item()


def test_dpm_coefficients_reproduce_oracle_update_on_cpu():
    """Evaluate the folded-sign coefficient form with plain torch and compare with the oracle's formulas."""
    g = torch.Generator().manual_seed(0)
    for algo in ("dpmsolver++", "dpmsolver"):
        for stype in ("midpoint", "heun"):
            for order in (1, 2, 3):
                if algo == "dpmsolver" and order == 3:
                    with pytest.raises(UnboundLocalError):
                        coefs.dpm(SIG, 5, 3, algo, stype, "fp32", False)
                    continue
                i = 7
                x = torch.randn(2, 4, 64, generator=g)
                m = [torch.randn(2, 4, 64, generator=g) for _ in range(3)]          # m[-1] = m0 (current x0)
                nz = torch.randn(2, 4, 64, generator=g)
                k, sc = coefs.dpm(SIG, i, order, algo, stype, "fp32", False)
                c = [torch.tensor(v) for v in k.c]
                m0, m1, m2 = m[2], m[1], m[0]
                d1 = d2 = torch.zeros_like(x)
                if order == 2:
                    d1 = c[1] * (m0 - m1)
                if order == 3:
                    d10, d11 = c[1] * (m0 - m1), c[2] * (m1 - m2)
                    d1, d2 = d10 + c[3] * (d10 - d11), c[4] * (d10 - d11)
                mean = c[5] * x + c[6] * m0
                ode = c[9] * x + c[10] * m0
                if order >= 2:
                    mean, ode = mean + c[7] * d1, ode + c[11] * d1
                if order == 3:
                    mean, ode = mean + c[8] * d2, ode + c[12] * d2
                for sde in (False, True):
                    if order == 1:
                        ref = O.dpm_first_order(algo, m0, SIG, i, x, nz, sde)
                    elif order == 2:
                        ref = O.dpm_second_order(algo, stype, m, SIG, i, x, nz, sde)
                    else:
                        ref = O.dpm_third_order(algo, m, SIG, i, x, nz, sde)
                    got = mean + c[13] * nz if sde else ode
                    assert torch.equal(got, ref[0]), (algo, stype, order, sde)
                    assert torch.equal(mean, ref[1])
                    assert sc == (ref[2] * ref[3]).item()


def test_flash_schedule_and_order_selection():
    args = types.SimpleNamespace(dpm_post_compress_ratio=0.4, shift=3.0, dpm_solver_order=2)
    det = [True] * 25
    for i in (2, 3, 4, 5):
        det[i] = False
    mine, last = su._flash_schedule(args, SIG, det)
    ref, rlast = O.flash_schedule(SIG, det, 3.0, 0.4)
    assert last == rlast == 5 and torch.equal(mine, ref) and mine.numel() == 6 + 8
    st = su.DPMState(order=2)
    assert su._dpm_order(args, 0, 12, st) == 1
    st.lower_order_nums = 1
    assert su._dpm_order(args, 6, 12, st) == 2 and su._dpm_order(args, 11, 12, st) == 1
    assert su._dpm_order(args, 6, 12, None) == 1
    args.dpm_solver_order = 3
    st = su.DPMState(order=3)
    st.lower_order_nums = 2
    assert su._dpm_order(args, 5, 25, st) == 3 and su._dpm_order(args, 10, 12, st) == 2


def test_window_mask_and_schedule():
    assert R.window_mask(6, [1, 2]) == [True, False, False, True, True, True]
    assert R.window_mask(3, [], "all") == [False] * 3
    assert torch.equal(R.sigma_schedule(25, 3.0), SIG)


def test_balance_pos_neg_matches_reference():
    from oracle import ref_loader
    ru = ref_loader.load_reward_utils()
    samples = [{"advantages": torch.tensor([a]), "id": i} for i, a in enumerate([0.3, -1.2, 0.0, 2.0, -0.1, -0.7, 0.9, -3.0, -0.2])]
    mine = R.balance_pos_neg(samples, rng=random.Random(5))
    ids = [s["id"] for s in mine]
    assert sorted(ids) == [0, 1, 3, 4, 5, 6, 7, 8]                      # the zero-advantage sample is dropped
    signs = [samples[i]["advantages"].item() > 0 for i in ids]
    assert signs[:6] == [True, False] * 3 and not any(signs[6:])         # interleaved, then the larger group's rest
    if ru is not None:
        for use_random in (False, True):
            random.seed(123)
            ref = ru.balance_pos_neg(list(samples), use_random=use_random)
            random.seed(123)
            got = R.balance_pos_neg(list(samples), use_random=use_random)
            assert [s["id"] for s in ref] == [s["id"] for s in got]


def test_make_samples_slices_and_reward_stacking():
    lat = torch.arange(2 * 6 * 3, dtype=torch.float32).view(2, 6, 3)
    lp = torch.arange(10, dtype=torch.float32).view(2, 5)
    sig = O.sd3_time_shift(3.0, torch.linspace(1, 0, 6))
    s = R.make_samples(lat, lp, sig, 5)
    a, b, c = GO.sample_slices(lat, lp)
    assert torch.equal(s["latents"], a) and torch.equal(s["next_latents"], b) and torch.equal(s["log_probs"], c)
    assert s["timesteps"].tolist() == [[int(x * 1000) for x in sig][:5][:-1]] * 2
    mat, names = grpo.stack_rewards({"a": torch.ones(4), "b": torch.zeros(4)})
    assert names == ["a", "b"] and mat.shape == (2, 4)
    single = grpo.gather_rewards(torch.arange(4.0))
    assert torch.equal(single, torch.arange(4.0))                        # no process group: identity (TR:333-334)
    d = grpo.gather_rewards({"a": torch.ones(3)})
    assert list(d) == ["a"] and torch.equal(d["a"], torch.ones(3))


def test_philox_oracle_known_answers():
    """Random123 known-answer vectors for Philox4x32-10 (zeros / ones / pi) pin the host restatement of the in-kernel
    noise generator; the GPU test then pins the kernel to the host restatement."""
    import numpy as np
    from oracle import philox_oracle as P
    z = np.zeros(1, dtype=np.uint64)
    f = np.full(1, 0xffffffff, dtype=np.uint64)
    assert [int(a[0]) for a in P.philox4x32_10(z, z, z, z, 0, 0)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert [int(a[0]) for a in P.philox4x32_10(f, f, f, f, 0xffffffff, 0xffffffff)] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    pi = [np.array([w], dtype=np.uint64) for w in (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344)]
    assert [int(a[0]) for a in P.philox4x32_10(*pi, 0xa4093822, 0x299f31d0)] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    n = P.normal(7, 12, 1 << 18)
    assert abs(n.mean()) < 5e-3 and abs(n.std() - 1) < 5e-3 and np.isfinite(n).all()
    assert not np.array_equal(P.normal(7, 16, 64), n[:64]) and np.array_equal(P.normal(7, 12, 64), n[:64])


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` works on a CPU-only box and prints the contract keys: the unmodified reference when a tree
    is reachable (kind "reference"), the oracle restatement otherwise (kind "port")."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=str(root))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "GB/s" and line["value"] > 0 and line["gpu_launches"] == 0
    from oracle import ref_loader
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_loader.reference_root() is not None else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["reference_control_flow"]["value"] > 0
    assert line["config"]["group_size"] == 12 and line["warmup"] == 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


@pytest.mark.parametrize("shift", [1.0, 2.0, 3.0, 5.0])
def test_train_timesteps_use_fp32_product_like_the_reference(shift):
    """TR:401 / SU:63-65: ``int(sigma * 1000)`` on an fp32 0-dim tensor.  A python-double product is off by one for many
    (N, shift) pairs (e.g. N=20, shift=3 -> 899 vs 900 at i=5), which would make the policy update evaluate the DiT at a
    different timestep than the rollout did."""
    differs_from_double = 0
    for n in range(4, 51):
        sig = O.sd3_time_shift(shift, torch.linspace(1, 0, n + 1))
        ref = [int(s * 1000) for s in sig][:n]                                  # the reference expression, verbatim
        assert R.timestep_values(sig, n) == ref
        host = coefs.host_schedule(sig)
        assert [int(host[i] * 1000) for i in range(n)] == ref                   # what run_sample_step feeds the transformer
        differs_from_double += [int(s * 1000) for s in sig.tolist()][:n] != ref
        x = torch.zeros(2, n + 1, 1, 64)
        smp = R.make_samples(x, torch.zeros(2, n), sig, n)
        assert smp["timesteps"].tolist() == [ref[:-1]] * 2
    if shift in (1.0, 3.0):
        assert differs_from_double > 0                                          # the sweep does cover the off-by-one cases


def test_fixed_point_logprob_reduction_error_model():
    """The kernels' log-prob reduction (csrc/step_math.cuh) is integer from the thread up: a thread's fp32 sum of d^2 over its 8 scalars
    is scaled by 2^32 / (n * 2 s^2) and rounded to an integer, everything above is exact integer addition.  Restated here on the CPU
    (torch fp32 for the per-thread part, int64 above) against the fp64 mean: the claimed 0.29 * sqrt(n / 8) units of 2^-32 (1.2e-8 at
    1024^2) of quantisation holds on top of the fp32 rounding of the scale itself, for on-policy residuals and for a policy that has drifted to |d|/s ~ 5; the per-thread cap leaves the fast path
    more than a factor 10 of headroom over the largest on-policy thread share at both ends of the size range."""
    import torch
    g = torch.Generator().manual_seed(5)
    for n, spread in ((4096 * 64, 1.0), (4096 * 64, 5.0), (256 * 64, 1.0), (1024 * 64, 3.0)):
        B, two_var = 3, 2.0 * 0.31 ** 2
        d = torch.randn(B, n, generator=g) * (0.31 * spread)
        denom = torch.tensor(float(n), dtype=torch.float32) * torch.tensor(two_var, dtype=torch.float32)     # lp_quant(): three IEEE single ops
        scale = torch.tensor(4294967296.0, dtype=torch.float32) / denom
        fit = 1095216660480.0 / float((n + 7) // 8)
        cap = min(fit, 67108864.0)
        dd = (d * d).view(B, n // 8, 4, 2)
        acc = torch.zeros(B, n // 8)
        for j in range(4):                                     # acc += dd[2j] + dd[2j+1], pair-wise, in order
            acc = acc + (dd[:, :, j, 0] + dd[:, :, j, 1])
        u = acc * scale
        assert float(u.max()) * 10 < cap or spread > 1.0       # on-policy: the LARGEST thread share is > 10x below the cap (19x at 256^2, 150x at 1024^2)
        assert float(u.max()) < cap                            # all of these stay on the fast path
        q = torch.round(u.double()).to(torch.int64).sum(dim=1)  # round-to-nearest-even like cvt.rni, then exact integer sums
        got = q.double() / 4294967296.0
        want = (d.double() ** 2).sum(dim=1) / (float(n) * two_var)
        err = (got - want).abs().max().item()
        # quantisation 0.29 * sqrt(n/8) * 2^-32 (random, absolute) plus what any fp32 evaluation has: the scale is two fp32 roundings
        # of n * 2 s^2 (systematic, relative <= 2^-23) and each thread's own 8-term sum carries relative 2^-24 (random)
        bound = 4 * (0.29 * (n / 8) ** 0.5 / 4294967296.0) + float(want.max()) * 2.0 ** -23
        assert err < bound, (n, spread, err, bound)
        assert err < 1.2e-7 * max(1.0, float(want.max()))      # one fp32 ulp of the quantity it feeds
