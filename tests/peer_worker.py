"""Worker of tests/test_gpu_peer.py::test_two_processes_over_cuda_ipc and of `bench.py`-independent N>1 checks:
launched under torch.distributed.run, one rank per GPU.  Exercises the fused peer exchange across REAL peers (regions
mapped with CUDA IPC, pushes travel over NVLink) against the oracle and against NCCL, eager and as replayed CUDA graphs."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    from mixgrpo_b200 import _cabi, grpo
    from mixgrpo_b200.peer import PeerExchange
    from oracle import grpo_oracle as GO
    _cabi.lib().mixgrpo_set_tuning(2, 20000)
    weights = {"hps": 1.0, "pick": 0.5, "ir": 2.0}
    px = PeerExchange()
    assert px.world == world and px.rank == rank
    local_B = 12

    def rewards_of(r, it):
        g = torch.Generator().manual_seed(1000 * it + r)
        return {k: torch.randn(local_B, generator=g) for k in weights}

    for it in range(5):
        mine = rewards_of(rank, it)
        allr = [rewards_of(q, it) for q in range(world)]
        mine_dev = {k: v.to(dev) for k, v in mine.items()}
        adv, gathered = px.gather_advantages(mine_dev, 12, weights)
        nccl = grpo.gather_rewards(mine_dev)                                   # the NCCL path it replaces
        assert torch.allclose(adv.cpu(), GO.group_advantages(mine, 12, weights), rtol=0, atol=1e-6)
        for k in weights:
            assert torch.equal(gathered[k], nccl[k])
            assert torch.equal(gathered[k].cpu(), torch.cat([allr[q][k] for q in range(world)]))
        # a group of 12*world samples split across the ranks
        adv_s, _ = px.gather_advantages(mine_dev, 12 * world, weights, mode="split")
        full = GO.group_advantages({k: torch.cat([allr[q][k] for q in range(world)]) for k in weights}, 12 * world, weights)
        assert torch.allclose(adv_s.cpu(), full[rank * local_B:(rank + 1) * local_B], rtol=0, atol=1e-6)
        # no-group normalisation (TR:498)
        adv_g, _ = px.gather_advantages(mine_dev["hps"], 12, mode="global")
        cat = torch.cat([allr[q]["hps"] for q in range(world)])
        assert torch.allclose(adv_g.cpu(), GO.group_advantages(mine["hps"], 12, use_group=False, gathered=cat), rtol=0, atol=1e-6)
        stats = (torch.tensor([1.0, 2.0, 3.0, 4.0]) * (rank + 1 + it)).to(dev)
        ref = stats.clone()
        px.allreduce_stats(stats)
        dist.all_reduce(ref, op=dist.ReduceOp.AVG)
        assert torch.allclose(stats, ref, rtol=1e-6, atol=0), (stats, ref)
    # CUDA graph: 20 replays, no NCCL inside
    buf = {k: torch.zeros(local_B, device=dev) for k in weights}
    st = torch.zeros(4, device=dev)
    w_dev = torch.tensor(list(weights.values()), device=dev)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        px.gather_advantages(buf, 12, w_dev)
        px.allreduce_stats(st)
    torch.cuda.synchronize()
    dist.barrier()
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph, stream=s):
        adv, gathered = px.gather_advantages(buf, 12, w_dev)
        px.allreduce_stats(st)
    for it in range(20):
        mine = rewards_of(rank, 100 + it)
        for k in weights:
            buf[k].copy_(mine[k])
        st.copy_(torch.tensor([1.0, 2.0, 3.0, 4.0]) * (rank + it))
        gph.replay()
        torch.cuda.synchronize()
        assert torch.allclose(adv.cpu(), GO.group_advantages(mine, 12, weights), rtol=0, atol=1e-6)
        assert torch.equal(gathered["ir"].cpu(), torch.cat([rewards_of(q, 100 + it)["ir"] for q in range(world)]))
        want = torch.tensor([1.0, 2.0, 3.0, 4.0]) * (sum(q + it for q in range(world)) / world)
        assert torch.allclose(st.cpu(), want, rtol=1e-6)
    # end to end: one training step (BASELINE configs[0] shape: tiny FLUX-shaped transformer, 256^2 latents, group 4) per rank, with NCCL
    # collectives and with the fused exchange — same advantages, same local gradients, same rank-averaged statistics
    import types
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    from tiny_flux import TinyFluxTransformer
    from mixgrpo_b200 import trainer
    targs = types.SimpleNamespace(w=256, h=256, sampling_steps=10, shift=3.0, eta=0.7, flow_grpo_sampling=True, dpm_algorithm_type="null",
                                  dpm_apply_strategy="post", dpm_post_compress_ratio=0.4, dpm_solver_order=2, dpm_solver_type="midpoint",
                                  sample_strategy="progressive", drop_last_sample=False, training_strategy="part", init_same_noise=False,
                                  clip_range=1e-4, adv_clip_max=5.0, kl_coeff=0.01, gradient_accumulation_steps=2, num_generations=4,
                                  use_group=True, advantage_rerange_strategy="null", trimmed_ratio=0.0)
    torch.manual_seed(0)
    model = TinyFluxTransformer().to(dev)
    gg = torch.Generator(device=dev).manual_seed(77 + rank)
    enc, pooled, tids = torch.randn(1, 6, 32, device=dev, generator=gg), torch.randn(1, 16, device=dev, generator=gg), torch.zeros(1, 3, device=dev)
    lat0 = torch.randn(4, 16, 32, 32, device=dev, generator=gg).bfloat16()
    nzs = [torch.randn(4, 256, 64, device=dev, generator=gg).bfloat16() for _ in range(10)]
    rw = {"hps": torch.randn(4, device=dev, generator=gg), "pick": torch.randn(4, device=dev, generator=gg)}
    res = []
    for exchange in (None, px):
        model.zero_grad(set_to_none=True)
        stats, gres, _, adv_t = trainer.train_one_step(targs, dev, model, lambda lat: rw, [3, 4, 5, 6], {"hps": 1.0, "pick": 0.5}, enc, pooled, tids,
                                                       exchange=exchange, input_latents=lat0, noises=nzs)
        res.append((stats.clone(), adv_t.clone(), gres["hps"].clone(), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}))
    assert torch.equal(res[0][1], res[1][1])
    assert torch.allclose(res[0][0], res[1][0], rtol=1e-5, atol=1e-9), (res[0][0], res[1][0])
    assert torch.allclose(res[0][2], res[1][2], rtol=1e-6)
    for n, g0 in res[0][3].items():
        assert ((res[1][3][n] - g0).norm() / g0.norm().clamp_min(1e-20)).item() < 1e-5, n
    # extended mode: ONE group of 4*world samples spread over the ranks (4 each): statistics from the gathered rewards
    targs.num_generations = 4 * world
    adv_split = []
    for exchange in (None, px):
        model.zero_grad(set_to_none=True)
        _, _, _, adv_t = trainer.train_one_step(targs, dev, model, lambda lat: rw, [3, 4, 5, 6], {"hps": 1.0, "pick": 0.5}, enc, pooled, tids,
                                                exchange=exchange, input_latents=lat0, noises=nzs, split_groups=True)
        adv_split.append(adv_t.clone())
    all_rw = [None] * world
    dist.all_gather_object(all_rw, {k: v.cpu() for k, v in rw.items()})
    full = GO.group_advantages({k: torch.cat([r[k] for r in all_rw]) for k in rw}, 4 * world, {"hps": 1.0, "pick": 0.5})
    assert torch.equal(adv_split[0], adv_split[1])
    assert torch.allclose(adv_split[1].cpu(), full[rank * 4:(rank + 1) * 4], rtol=0, atol=1e-6)
    # latency: fused exchange vs NCCL all_gather + advantage kernel (CUDA events, 200 calls each)
    mine_dev = {k: torch.randn(local_B, device=dev) for k in weights}
    def timed(fn, n=200):
        for _ in range(20):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n * 1e3
    t_fused = timed(lambda: px.gather_advantages(mine_dev, 12, w_dev))
    t_nccl = timed(lambda: (grpo.gather_rewards(mine_dev), grpo.compute_group_advantages(mine_dev, 12, w_dev)))
    t_graph = timed(gph.replay)
    st2 = torch.zeros(4, device=dev)
    t_red = timed(lambda: px.allreduce_stats(st2))
    t_red_nccl = timed(lambda: dist.all_reduce(st2, op=dist.ReduceOp.AVG))
    assert px.status()[2] == 0
    print(f"PEER_OK rank {rank}/{world}: gather+adv fused {t_fused:.1f} us vs NCCL all_gather + adv kernel {t_nccl:.1f} us; "
          f"allreduce fused {t_red:.1f} us vs NCCL {t_red_nccl:.1f} us; graph(gather+adv, allreduce) {t_graph:.1f} us", flush=True)
    px.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
