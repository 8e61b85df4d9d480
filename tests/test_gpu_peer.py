"""Fused peer-memory exchange (csrc/peer_kernels.cu): reward gather + advantages and the logging all-reduce.

Single-GPU tests drive the real protocol with several endpoints ("ranks") on ONE device, one CUDA stream each — the
kernels push into each other's regions, publish flags and spin exactly as they do across NVLink.  The two-process test
(needs >= 2 GPUs, skipped otherwise) maps the regions through CUDA IPC under torchrun-style ranks.  The checker is the
oracle's restatement of TR:439-501 (oracle/grpo_oracle.py), like every other parity test."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
WEIGHTS = {"hps": 1.0, "pick": 0.5, "ir": 2.0}


def _mk(world, local_B, n_models, seed):
    g = torch.Generator().manual_seed(seed)
    names = list(WEIGHTS)[:n_models]
    return [{k: torch.randn(local_B, generator=g) for k in names} for _ in range(world)]


def _stage(local, dev):
    """Inputs on the device BEFORE the concurrent launches: no host-blocking copy between rank 0's and rank 1's launch."""
    from mixgrpo_b200 import grpo
    out = [{k: v.to(dev) for k, v in d.items()} for d in local]
    grpo._weights_on(dev, tuple(WEIGHTS[k] for k in local[0]))
    torch.cuda.synchronize()
    return out


def _run_all(eps, fn):
    """Launch fn(endpoint, rank) for every endpoint on its own stream, then join."""
    dev = eps[0].device
    streams = [torch.cuda.Stream(dev) for _ in eps]
    cur = torch.cuda.current_stream(dev)
    out = []
    for r, (e, s) in enumerate(zip(eps, streams)):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            out.append(fn(e, r))
    for s in streams:
        cur.wait_stream(s)
    torch.cuda.synchronize(dev)
    return out


@pytest.fixture
def short_timeout():
    from mixgrpo_b200 import _cabi
    old = _cabi.lib().mixgrpo_set_tuning(2, 4000)       # a protocol bug must fail the test, not hang the box
    yield
    _cabi.lib().mixgrpo_set_tuning(2, old)


@pytest.mark.parametrize("world,local_B,G,n_models,trim", [
    (1, 12, 12, 3, 0.0), (2, 12, 12, 3, 0.0), (2, 24, 12, 1, 0.0), (4, 12, 4, 2, 0.0), (8, 12, 12, 3, 0.0),
    (8, 24, 24, 3, 0.25), (2, 13, 4, 2, 0.0),
])
def test_gather_advantages_local_groups(world, local_B, G, n_models, trim, short_timeout):
    from mixgrpo_b200.peer import PeerExchange
    from oracle import grpo_oracle as GO
    dev = torch.device("cuda:0")
    eps = PeerExchange.local_ranks(world, dev)
    try:
        for call in range(3):                                       # three calls: both buffer parities and the wrap
            local = _mk(world, local_B, n_models, 10 * call + world)
            names = list(local[0])
            w = {k: WEIGHTS[k] for k in names}
            on_dev = _stage(local, dev)
            res = _run_all(eps, lambda e, r: e.gather_advantages(on_dev[r], G, w, trim))
            for r, (adv, gathered) in enumerate(res):
                want = GO.group_advantages(local[r], G, w, trimmed_ratio=trim)
                assert torch.allclose(adv.cpu(), want, rtol=0, atol=1e-6), (r, call)
                for k in names:                                     # torch.cat order of TR:338
                    assert torch.equal(gathered[k].cpu(), GO.gather_cat([local[q][k] for q in range(world)]))
            assert all(e.status() == (call + 1, 0, 0) for e in eps)
    finally:
        for e in eps:
            e.close()


@pytest.mark.parametrize("world,local_B,G", [(2, 12, 24), (8, 3, 24), (4, 6, 12), (4, 5, 4)])
def test_gather_advantages_group_split_across_ranks(world, local_B, G, short_timeout):
    from mixgrpo_b200.peer import PeerExchange
    from oracle import grpo_oracle as GO
    dev = torch.device("cuda:0")
    eps = PeerExchange.local_ranks(world, dev)
    try:
        local = _mk(world, local_B, 3, 7)
        on_dev = _stage(local, dev)
        res = _run_all(eps, lambda e, r: e.gather_advantages(on_dev[r], G, WEIGHTS, mode="split"))
        full = GO.group_advantages({k: torch.cat([local[q][k] for q in range(world)]) for k in WEIGHTS}, G, WEIGHTS)
        for r, (adv, _) in enumerate(res):
            assert torch.allclose(adv.cpu(), full[r * local_B:(r + 1) * local_B], rtol=0, atol=1e-6), r
    finally:
        for e in eps:
            e.close()


def test_gather_advantages_global_and_plain_tensor(short_timeout):
    from mixgrpo_b200.peer import PeerExchange
    from oracle import grpo_oracle as GO
    dev = torch.device("cuda:0")
    world, local_B = 4, 12
    eps = PeerExchange.local_ranks(world, dev)
    try:
        g = torch.Generator().manual_seed(3)
        local = [torch.randn(local_B, generator=g) for _ in range(world)]
        cat = torch.cat(local)
        on_dev = [t.to(dev) for t in local]
        torch.cuda.synchronize()
        res = _run_all(eps, lambda e, r: e.gather_advantages(on_dev[r], 12, mode="global"))
        for r, (adv, gathered) in enumerate(res):
            assert torch.allclose(adv.cpu(), GO.group_advantages(local[r], 12, use_group=False, gathered=cat), rtol=0, atol=1e-6)
            assert torch.equal(gathered.cpu(), cat)
        res = _run_all(eps, lambda e, r: e.gather_advantages(on_dev[r], 4, trimmed_ratio=0.3))
        for r, (adv, _) in enumerate(res):
            assert torch.allclose(adv.cpu(), GO.group_advantages(local[r], 4, trimmed_ratio=0.3), rtol=0, atol=1e-6)
        with pytest.raises(ValueError):
            eps[0].gather_advantages({"hps": local[0].to(dev)}, 12, WEIGHTS, mode="global")
    finally:
        for e in eps:
            e.close()


def test_matches_the_two_launch_path_bitwise(short_timeout):
    """The fused kernel and all_gather + mixgrpo_group_advantages share their arithmetic: identical bits."""
    from mixgrpo_b200 import grpo
    from mixgrpo_b200.peer import PeerExchange
    dev = torch.device("cuda:0")
    eps = PeerExchange.local_ranks(2, dev)
    try:
        local = _mk(2, 24, 3, 11)
        on_dev = _stage(local, dev)
        res = _run_all(eps, lambda e, r: e.gather_advantages(on_dev[r], 12, WEIGHTS, 0.1))
        for r, (adv, _) in enumerate(res):
            two = grpo.compute_group_advantages(on_dev[r], 12, WEIGHTS, 0.1)
            assert torch.equal(adv, two)
    finally:
        for e in eps:
            e.close()


@pytest.mark.parametrize("world", [1, 2, 8])
def test_allreduce_rank_order_sum_is_bitwise_reproducible(world, short_timeout):
    from mixgrpo_b200.peer import PeerExchange
    dev = torch.device("cuda:0")
    eps = PeerExchange.local_ranks(world, dev)
    try:
        g = torch.Generator().manual_seed(5)
        for call, (count, avg) in enumerate([(4, True), (4, False), (256, True), (1, True)]):
            vals = [torch.randn(count, generator=g) * 10 ** (r % 3) for r in range(world)]
            want = vals[0].clone()
            for q in range(1, world):
                want = want + vals[q]                                # fp32, rank order
            if avg:
                want = want / world
            on_dev = [v.to(dev) for v in vals]
            torch.cuda.synchronize()
            res = _run_all(eps, lambda e, r: e.allreduce_stats(on_dev[r], average=avg))
            for out in res:
                assert torch.equal(out.cpu(), want), call
        # longer vectors travel as 256-entry messages (one launch each): [4 x 24 x 4] stats rows of a group-24 window
        vals = [torch.randn(4, 24, 4, generator=g) * (r + 1) for r in range(world)]
        want = vals[0].clone()
        for q in range(1, world):
            want = want + vals[q]
        on_dev = [v.to(dev) for v in vals]
        torch.cuda.synchronize()
        for out in _run_all(eps, lambda e, r: e.allreduce_stats(on_dev[r], average=False)):
            assert torch.equal(out.cpu(), want)
        with pytest.raises(ValueError):
            eps[0].allreduce_stats(torch.zeros(0, device=dev))
        with pytest.raises(ValueError):
            eps[0].allreduce_stats(torch.zeros(8, device=dev, dtype=torch.float64))
    finally:
        for e in eps:
            e.close()


def test_graph_capture_and_replay(short_timeout):
    """Pointers and arguments are fixed and the call counter lives in the region: one captured graph per rank replays."""
    from mixgrpo_b200.peer import PeerExchange
    from oracle import grpo_oracle as GO
    dev = torch.device("cuda:0")
    world, local_B = 2, 12
    eps = PeerExchange.local_ranks(world, dev)
    try:
        bufs = [{k: torch.zeros(local_B, device=dev) for k in WEIGHTS} for _ in range(world)]
        stats = [torch.zeros(4, device=dev) for _ in range(world)]
        w_dev = torch.tensor([WEIGHTS[k] for k in WEIGHTS], device=dev)
        graphs, outs, streams = [], [], [torch.cuda.Stream(dev) for _ in range(world)]
        _run_all(eps, lambda e, r: (e.gather_advantages(bufs[r], 12, w_dev), e.allreduce_stats(stats[r])))   # warm-up, eager
        torch.cuda.synchronize()
        # capture one rank at a time: capture only records, so no peer needs to run concurrently
        for r in range(world):
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph, stream=streams[r]):
                adv, gathered = eps[r].gather_advantages(bufs[r], 12, w_dev)
                eps[r].allreduce_stats(stats[r])
            graphs.append(gph)
            outs.append((adv, gathered))
        for it in range(4):
            local = _mk(world, local_B, 3, 50 + it)
            for r in range(world):
                for k in WEIGHTS:
                    bufs[r][k].copy_(local[r][k])
                stats[r].copy_(torch.tensor([1.0, 2.0, 3.0, 4.0]) * (r + 1 + it))
            torch.cuda.synchronize()
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    graphs[r].replay()
            torch.cuda.synchronize()
            for r in range(world):
                assert torch.allclose(outs[r][0].cpu(), GO.group_advantages(local[r], 12, WEIGHTS), rtol=0, atol=1e-6)
                assert torch.equal(outs[r][1]["pick"].cpu(), torch.cat([local[q]["pick"] for q in range(world)]))
                want = torch.tensor([1.0, 2.0, 3.0, 4.0]) * ((1 + it) + (2 + it)) / 2
                assert torch.allclose(stats[r].cpu(), want)
        assert all(e.status()[2] == 0 for e in eps)
    finally:
        for e in eps:
            e.close()


def test_missing_peer_times_out_with_nan_instead_of_hanging():
    from mixgrpo_b200 import _cabi
    from mixgrpo_b200.peer import PeerExchange
    dev = torch.device("cuda:0")
    old = _cabi.lib().mixgrpo_set_tuning(2, 200)
    eps = PeerExchange.local_ranks(2, dev)
    try:
        adv, gathered = eps[0].gather_advantages(torch.randn(12, device=dev), 12)     # rank 1 never shows up
        torch.cuda.synchronize()
        assert torch.isnan(gathered).all() and torch.isnan(adv).all()
        assert eps[0].status()[2] == 1
    finally:
        _cabi.lib().mixgrpo_set_tuning(2, old)
        for e in eps:
            e.close()


def test_argument_validation():
    from mixgrpo_b200.peer import PeerExchange
    dev = torch.device("cuda:0")
    (e,) = PeerExchange.local_ranks(1, dev, cap_floats=16)
    try:
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            e.gather_advantages(torch.randn(12), 12)
        with pytest.raises(RuntimeError):                                            # 3*12 floats > cap 16
            e.gather_advantages({k: torch.randn(12, device=dev) for k in WEIGHTS}, 12, WEIGHTS)
        with pytest.raises(ValueError):
            e.gather_advantages({k: torch.randn(4, device=dev) for k in WEIGHTS}, 4)  # multi-reward without weights
    finally:
        e.close()


# ------------------------------------------------------------------ real peers: two processes, two GPUs, CUDA IPC
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_two_processes_over_cuda_ipc():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tests" / "peer_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("PEER_OK") == 2, r.stdout[-3000:] + r.stderr[-3000:]
