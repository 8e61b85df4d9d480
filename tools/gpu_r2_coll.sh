#!/bin/bash
set -u
for c in none peer; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 200 --skip-e2e --collectives $c 2>/dev/null | tail -1 | cut -c1-200
done
NCCL_P2P_DISABLE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 200 --skip-e2e --collectives none 2>/dev/null | tail -1 | cut -c1-200
python bench.py --steps 200 --skip-e2e | cut -c1-200
CUDA_VISIBLE_DEVICES=1 python bench.py --steps 200 --skip-e2e | cut -c1-200
(CUDA_VISIBLE_DEVICES=0 python bench.py --steps 3000 --skip-e2e | cut -c1-200) & (CUDA_VISIBLE_DEVICES=1 python bench.py --steps 3000 --skip-e2e | cut -c1-200); wait
