#!/bin/bash
# round-2 multi-GPU session: N = $1 ranks — cross-process peer tests, bench (default line), pinned-copy ceiling
set -u
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then python -m pytest tests/test_gpu_peer.py -m gpu -x -q > gpurun_out/r2_peer_tests_n$N.log 2>&1; echo "peer tests rc=$?"; tail -3 gpurun_out/r2_peer_tests_n$N.log; fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/h2d_ceiling.py > gpurun_out/r2_h2d_ceiling_n$N.json 2> gpurun_out/r2_h2d_ceiling_n$N.err; echo "ceiling rc=$?"; cat gpurun_out/r2_h2d_ceiling_n$N.json
