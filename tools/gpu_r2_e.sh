#!/bin/bash
# round-2 GPU session E (after ABI v6): ncu --set full of the step's instantiations as the final step launches them + launch list of the bench
set -u
mkdir -p gpurun_out
python tools/ncu_targets.py > gpurun_out/r2e_targets_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:'step_kernel|finalize|multi_kernel|logprob_bwd' -f -o /tmp/r02e_targets python tools/ncu_targets.py > gpurun_out/r2e_ncu_targets.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r02e_targets.ncu-rep --page raw --csv > gpurun_out/r02e_targets_raw.csv 2>/dev/null
ls -la gpurun_out/r02e_targets_raw.csv; tail -2 gpurun_out/r2e_targets_plain.log
python bench.py --steps 2 --warmup 3 --profile-only > gpurun_out/r2e_bench_short.json 2> gpurun_out/r2e_bench_short.err && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02e_bench_launches.csv python bench.py --steps 2 --warmup 3 --profile-only > gpurun_out/r2e_ncu_bench.log 2>&1; echo "launch list rc=$?"
wc -l gpurun_out/r02e_bench_launches.csv
