"""The compiled binding (mixgrpo_b200/_lib/_torchbind.so, csrc_bind/torch_bind.cpp) is a second LOADER of the same C ABI,
not a second implementation: it must load, agree on the ABI version, refuse CPU tensors like the ctypes layer, and — on a
GPU — give bit-identical results to the ctypes path for every entry point it covers, autograd included."""
import contextlib
import os

import pytest
import torch

from mixgrpo_b200 import _cabi, coefs, ops
from mixgrpo_b200 import sampling_utils as su

SIG = su.sd3_time_shift(3.0, torch.linspace(1, 0, 26))
pytestmark = pytest.mark.skipif(os.environ.get("MIXGRPO_BINDING") == "ctypes", reason="the compiled binding is switched off (MIXGRPO_BINDING=ctypes)")


@contextlib.contextmanager
def ctypes_loader():
    """Route every ops.* call through the ctypes loader for the duration (the compiled binding stays loaded)."""
    ops.binding()
    saved = ops._binding_mod
    ops._binding_mod = None
    try:
        yield
    finally:
        ops._binding_mod = saved


def test_compiled_binding_loads_and_matches_the_abi():
    tb = ops.binding()
    assert tb is not None, "the compiled binding must build and load wherever a C++ compiler and torch headers are present"
    assert tb.abi_version() == _cabi.ABI_VERSION == _cabi.lib().mixgrpo_abi_version()
    for name in ("fused_step", "logprob_backward", "transition_logprob", "policy_forward", "policy_backward"):
        assert callable(getattr(tb, name))


def test_compiled_binding_refuses_cpu_tensors():
    tb = ops.binding()
    k, _ = coefs.flow(SIG, 3, 0.7, "ref_cuda", False)
    x = torch.randn(2, 4, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.fused_step(ops.FLOW, x, x, k, src=_cabi.SRC_DETERMINISTIC)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.logprob_backward(ops.FLOW, x, x, x, torch.randn(2), k)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.policy_forward(ops.FLOW, x, x, x, k, torch.randn(2), torch.randn(2), 1e-4, 5.0, 0.0, 12.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tb.transition_logprob(x, x, x, 0, 0, False, False)          # the CUDA check fires before the coefficient block is read


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_both_loaders_are_bit_identical(dtype):
    d = torch.device("cuda:0")
    g = torch.Generator(device=d).manual_seed(3)
    B, S = 5, 256
    traj = torch.randn(B, 3, S, 64, device=d, generator=g)
    x = traj[:, 0]                                                       # strided rows: batch stride 3 * S * 64
    v = torch.randn(B, S, 64, device=d, generator=g).to(dtype)
    eps = torch.randn(B, S, 64, device=d, generator=g).to(dtype)
    old = torch.randn(B, device=d, generator=g) * 0.01 - 1
    adv = torch.randn(B, device=d, generator=g)

    def run():
        out = list(su.flow_grpo_step(v, x, 0.7, SIG, 4, None, noise=eps)[:4])
        out += list(su.flow_grpo_step(v, x, 0.7, SIG, 4, None, determistic=True)[:3])
        out += list(su.dance_grpo_step(v, x, 0.7, SIG, 4, None, True, True, noise=eps.float()))
        xn = out[0]
        vg = v.clone().requires_grad_(True)
        lp = su.flow_grpo_step(vg, x, 0.7, SIG, 4, xn)[2]
        (lp * adv).sum().backward()
        out += [lp.detach(), vg.grad]
        vg2 = v.clone().requires_grad_(True)
        lp2 = su.dance_grpo_step(vg2, x, 0.7, SIG, 4, xn, True, True)[2]
        lp2.sum().backward()
        out += [lp2.detach(), vg2.grad]
        k, _ = coefs.flow(SIG, 4, 0.7, "ref_cuda", dtype == torch.bfloat16)
        rows = torch.zeros(B, 4, device=d)
        nl = ops.policy_forward(ops.FLOW, v, x, xn, k, old, adv, 1e-4, 5.0, 0.01, 12.0, stats_rows=rows, round_like_torch=True)
        gv = ops.policy_backward(ops.FLOW, v, x, xn, nl, k, old, adv, 1e-4, 5.0, 0.01, 12.0, round_like_torch=True, early_loads=True)
        out += [nl, rows, gv, ops.logprob_backward(ops.FLOW, v, x, xn, adv, k, True)]
        return out

    assert ops.binding() is not None
    a = run()
    with ctypes_loader():
        b = run()
    assert len(a) == len(b)
    for i, (p, q) in enumerate(zip(a, b)):
        assert p.dtype == q.dtype and p.shape == q.shape and torch.equal(p, q), i


@pytest.mark.gpu
def test_compiled_binding_error_types_match_the_ctypes_layer():
    d = torch.device("cuda:0")
    k, _ = coefs.flow(SIG, 3, 0.7, "ref_cuda", False)
    x = torch.randn(2, 8, 64, device=d)
    for ctx in (contextlib.nullcontext(), ctypes_loader()):
        with ctx:
            with pytest.raises(ValueError):
                ops.fused_step(ops.FLOW, x, x[:, :4], k, src=_cabi.SRC_DETERMINISTIC)             # shape mismatch
            with pytest.raises(TypeError):
                ops.fused_step(ops.FLOW, x.half(), x, k, src=_cabi.SRC_DETERMINISTIC)             # fp16 model output
            with pytest.raises(ValueError):
                ops.fused_step(ops.FLOW, x, x, k, src=_cabi.SRC_NOISE)                            # rollout without noise
            with pytest.raises(ValueError):
                ops.fused_step(ops.FLOW, x, x, k, src=_cabi.SRC_DETERMINISTIC, out_logp=torch.empty(3, device=d))
            assert ops.fused_step(ops.FLOW, x[:0], x[:0], k, src=_cabi.SRC_DETERMINISTIC)[2].shape == (0,)   # empty batch: no launch
