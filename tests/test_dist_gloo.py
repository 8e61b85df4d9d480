"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: ONE all_gather_into_tensor of the reward matrix in
rank order (replacing the per-model list all_gather + cat, TR:332-338, TR:417-425), ONE averaged [4] stats
reduction, and prompt-group partitioning with rank-local statistics (TR:443-461)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mixgrpo_b200 import grpo
        from oracle import grpo_oracle as GO
        g = torch.Generator().manual_seed(100 + rank)
        local = {"hps": torch.randn(12, generator=g), "pick": torch.randn(12, generator=g), "ir": torch.randn(12, generator=g)}
        gathered = grpo.gather_rewards(local)
        # reference: per-model list all_gather + cat
        ref = {}
        for k, t in local.items():
            parts = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            ref[k] = GO.gather_cat(parts)
        ok = all(torch.equal(gathered[k], ref[k]) for k in local) and list(gathered) == list(local)
        flat = grpo.gather_rewards(local["hps"])
        ok = ok and torch.equal(flat, ref["hps"])
        ok = ok and torch.equal(grpo.gather_tensor(local["pick"]), ref["pick"])          # TR:332-338 drop-in
        # rank-local group statistics: this rank's slice of the gathered vector gives back its own advantages
        adv_local = GO.group_advantages(local, 12, {"hps": 1.0, "pick": 0.5, "ir": 2.0})
        adv_from_gather = GO.group_advantages({k: v[rank * 12:(rank + 1) * 12] for k, v in gathered.items()}, 12, {"hps": 1.0, "pick": 0.5, "ir": 2.0})
        ok = ok and torch.equal(adv_local, adv_from_gather)
        # extended mode: ONE group of 24 split over the 2 ranks (12 each): statistics come from the gathered vector
        adv_full = GO.group_advantages(gathered["hps"], 24)
        mine = grpo.split_group_slice(adv_full, rank, world)
        ok = ok and mine.shape == (12,) and torch.equal(mine, adv_full[rank * 12:(rank + 1) * 12])
        ok = ok and not torch.equal(mine, GO.group_advantages(local["hps"], 12))          # differs from rank-local statistics
        # prompt partition: DistributedSampler(shuffle=False) order, padded by wrap-around
        from torch.utils.data import DistributedSampler
        for n_prompts in (7, 8, 1):
            ds = list(DistributedSampler(list(range(n_prompts)), num_replicas=world, rank=rank, shuffle=False))
            ok = ok and grpo.partition_prompts(n_prompts, rank, world) == ds
        stats = torch.tensor([1.0, 2.0, 3.0, 4.0]) * (rank + 1)
        red = grpo.reduce_step_stats(stats.clone())
        want = torch.tensor([1.0, 2.0, 3.0, 4.0]) * (sum(range(1, world + 1)) / world)
        ok = ok and torch.allclose(red, want)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_reward_gather_and_stats_reduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, True), (1, True)]
