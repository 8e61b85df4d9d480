"""-m gpu: bench.py's metric is algorithmic bytes / time, so its byte count must describe what a step really launches.
Each scenario's `step_plan` (SURVEY §8d vocabulary) is held against the launches `native_step` actually issues, at a small
token count; and the e2e / graph legs' shared helpers are exercised once."""
import collections
import importlib
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


@pytest.mark.parametrize("name", ["mixgrpo", "flash", "large_b24_512sq_f32", "mixgrpo_ode_logp_off"])
def test_step_plan_matches_the_launches(name, monkeypatch):
    bench = importlib.import_module("bench")
    from mixgrpo_b200 import ops
    from mixgrpo_b200._cabi import SRC_DETERMINISTIC, SRC_GIVEN
    B, S, dt, flash = bench.SCENARIOS[name]
    monkeypatch.setitem(bench.SCENARIOS, name, (B, 64, dt, flash))          # 64 tokens: same launches, tiny tensors
    dev = torch.device("cuda:0")
    w = bench.Workload(dev, 0, name)
    seen = collections.Counter()
    real = ops.fused_step

    def spy(family, v, x, coefs, *, src, want_x0=True, order=1, **kw):
        if family == ops.DPM:
            kind = f"dpm{order}_ode_x0"
        elif src == SRC_GIVEN:
            kind = "train_fwd"
        else:
            kind = ("ode" if src == SRC_DETERMINISTIC else "sde") + ("_x0" if want_x0 else "")
        seen[kind] += 1
        return real(family, v, x, coefs, src=src, want_x0=want_x0, order=order, **kw)

    monkeypatch.setattr(ops, "fused_step", spy)
    multi = collections.Counter()
    real_f, real_b = ops.policy_forward_multi, ops.policy_backward_multi
    monkeypatch.setattr(ops, "policy_forward_multi", lambda fam, vs, *a, **k: (multi.update(train_fwd=len(vs)), real_f(fam, vs, *a, **k))[1])
    monkeypatch.setattr(ops, "policy_backward_multi", lambda fam, vs, *a, **k: (multi.update(bwd=len(vs)), real_b(fam, vs, *a, **k))[1])
    before = ops.launch_count
    stats, logps, grads, adv = bench.native_step(w)
    torch.cuda.synchronize()
    launched = ops.launch_count - before
    want = dict(w.plan)
    got = dict(seen)
    got.update(multi)
    assert got == want, (got, want)
    # launches: the sampler steps + 1 finalize + 1 advantage kernel + window forward + window backward (+ a cast when the first
    # step cannot seed the trajectory: fp32 model output still seeds — z is bf16 in every scenario)
    assert launched == w.n_steps + 4, launched
    assert len(grads) == bench.WINDOW and logps.shape == (B, w.n_steps) and torch.isfinite(logps[:, w.window]).all()
    assert w.bytes_per_step == B * 64 * bench.C * sum(bench.bytes_per_elem(k, dt == torch.float32) * c for k, c in w.plan)


def test_h2d_ceiling_and_time_graph_helpers_run():
    bench = importlib.import_module("bench")
    dev = torch.device("cuda:0")
    gbs = bench.h2d_ceiling(dev, 8 << 20, reps=2)
    assert 1.0 < gbs < 200.0
    x = torch.zeros(1 << 20, device=dev)
    us = bench._time_graph(lambda: x.add_(1.0), 1, torch.cuda.Stream(device=dev), reps=5)
    assert 0.5 < us < 1000.0
