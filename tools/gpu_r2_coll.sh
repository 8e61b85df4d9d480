#!/bin/bash
set -u
echo "N=1 plain"; python bench.py --steps 200 --skip-e2e | cut -c1-100
echo "N=1 + px(world 1) side"; MIXGRPO_BENCH_PEER_N1=1 python bench.py --steps 200 --skip-e2e | cut -c1-100
echo "N=1 + px(world 1) head"; MIXGRPO_BENCH_PEER_N1=1 MIXGRPO_BENCH_PX_PLACE=head python bench.py --steps 200 --skip-e2e | cut -c1-100
for pl in side head; do echo "N=2 peer $pl"; MIXGRPO_BENCH_PX_PLACE=$pl python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 200 --skip-e2e --collectives peer 2>/dev/null | tail -1 | cut -c1-130; done
echo "N=2 none"; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 200 --skip-e2e --collectives none 2>/dev/null | tail -1 | cut -c1-130
