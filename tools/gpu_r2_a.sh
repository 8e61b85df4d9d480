#!/bin/bash
# round-2 GPU session A: default bench, reference arm, ncu captures of the step's instantiations (raw CSV pages come back, not the reports)
set -u
mkdir -p gpurun_out
python bench.py --steps 100 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
python tools/ncu_targets.py > gpurun_out/r2_targets_plain.log 2>&1 && \
ncu --set full --clock-control none -f -o /tmp/r02_targets python tools/ncu_targets.py > gpurun_out/r2_ncu_targets.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r02_targets.ncu-rep --page raw --csv > gpurun_out/r02_targets_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:step_kernel -c 1 -f -o gpurun_out/r02_ode_src python tools/ncu_targets.py > gpurun_out/r2_ncu_ode_src.log 2>&1; echo "ncu-src rc=$?"
ls -la gpurun_out/ | tail -12
