#!/usr/bin/env python
"""What a plain pinned-host -> device copy gets on THIS box with N processes copying at once — the ceiling of bench.py's e2e leg,
whose step time is the upload of 157 MB of model outputs (e2e scaled 1.00 / 1.00 / 0.55 / 0.45 at N = 1/2/4/8).

    python tools/h2d_ceiling.py                                                          # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/h2d_ceiling.py

One cudaMemcpyAsync per repetition from ONE pinned block (no batch-copy APIs), every rank starting together behind a barrier,
device-timed with CUDA events.  Prints one JSON line: per-rank GB/s for the e2e upload size and for 1 / 4 x that size, the
GPU's PCI address, its NUMA node and the CPUs this process may run on."""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 25 * 12 * 4096 * 64 * 2 + 3 * 12 * 4                     # bench.py's h2d_bytes_per_step
    res = {}
    for mult, name in ((1, "e2e_upload_157MB"), (4, "4x")):
        mine = bench.h2d_ceiling(dev, nbytes * mult, reps=5)
        res[name] = [round(x, 2) for x in bench.per_rank(mine, dev)]
    pci = torch.cuda.get_device_properties(dev)
    bdf = f"{pci.pci_domain_id:04x}:{pci.pci_bus_id:02x}:{pci.pci_device_id:02x}.0" if hasattr(pci, "pci_bus_id") else "?"
    numa = "?"
    try:
        numa = Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text().strip()
    except Exception:  # noqa: BLE001
        pass
    info = [None] * world
    me = {"rank": rank, "pci": bdf, "numa_node": numa, "cpus_allowed": len(os.sched_getaffinity(0))}
    if world > 1:
        dist.all_gather_object(info, me)
    else:
        info = [me]
    if rank == 0:
        print(json.dumps({"n_gpus": world, "gbs_per_rank": res, "min_gbs": {k: min(v) for k, v in res.items()}, "ranks": info}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
