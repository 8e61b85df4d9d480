#!/bin/bash
# round-2 GPU session D: ncu --set full of the step's instantiations exactly as the final step launches them
set -u
mkdir -p gpurun_out
python tools/ncu_targets.py > gpurun_out/r2_targets_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:'step_kernel|finalize|multi_kernel|logprob_bwd' -f -o /tmp/r02_targets python tools/ncu_targets.py > gpurun_out/r2_ncu_targets.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r02_targets.ncu-rep --page raw --csv > gpurun_out/r02_targets_raw.csv 2>/dev/null
ls -la gpurun_out/r02_targets_raw.csv; tail -2 gpurun_out/r2_targets_plain.log
