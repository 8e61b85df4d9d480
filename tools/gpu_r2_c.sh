#!/bin/bash
# round-2 GPU session C (1 GPU): final default bench + reference arm, ncu launch list of the timed region, eager host-cost table, sync cost
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --profile-only > gpurun_out/r2_profile_only.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --profile-only > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu rc=$?"
python tools/eager_reference_gpu.py > gpurun_out/r2_eager.jsonl 2> gpurun_out/r2_eager.err; echo "eager rc=$?"
python tools/host_overhead.py > gpurun_out/r2_host_overhead.log 2>&1
python - > gpurun_out/r2_sync_cost.log 2>&1 <<'PY'
import time, torch
from mixgrpo_b200 import grpo
d = torch.device("cuda:0")
r = torch.randn(3, 12, device=d); w = torch.tensor([1.0, 0.5, 2.0], device=d)
for _ in range(50): a = grpo.compute_group_advantages(r, 12, w).tolist()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(500): a = grpo.compute_group_advantages(r, 12, w).tolist()
print("advantages kernel + .tolist() (the only host sync of train_one_step's re-ranging, TR:527-535): %.1f us" % ((time.perf_counter() - t0) / 500 * 1e6))
t0 = time.perf_counter()
for _ in range(500): a = grpo.compute_group_advantages(r, 12, w)
torch.cuda.synchronize(); print("advantages kernel alone, async: %.1f us" % ((time.perf_counter() - t0) / 500 * 1e6))
PY
cat gpurun_out/r2_sync_cost.log
