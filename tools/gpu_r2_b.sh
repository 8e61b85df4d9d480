#!/bin/bash
# round-2 GPU session B (1 GPU): full gpu test suite (compiled binding is the default loader), eager host-cost table
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_tests.log
python tools/eager_reference_gpu.py > gpurun_out/r2_eager.jsonl 2> gpurun_out/r2_eager.err; echo "eager rc=$?"; cat gpurun_out/r2_eager.jsonl
python tools/host_overhead.py > gpurun_out/r2_host_overhead.log 2>&1; cat gpurun_out/r2_host_overhead.log
