#!/bin/bash
# One 8-GPU box visit: the headline bench at N = 8 as the driver launches it, then BASELINE configs[4]
# (1 048 576 robots = 8 x 131 072, 500 closed-loop ticks on device).  gpurun --gpus 8 --timeout 400 -- 'bash tools/gpu_scale8.sh TAG'
TAG=${1:-r01}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29511 bench.py --gpus 8 --steps 50 --warmup 5 > $O/${TAG}_bench8.json 2> $O/${TAG}_bench8.err || tail -5 $O/${TAG}_bench8.err
cat $O/${TAG}_bench8.json | cut -c1-600
timeout 150 $TR --master-port 29512 bench.py --gpus 8 --workload sweep --batch 131072 --steps 500 --warmup 5 > $O/${TAG}_sweep8.json 2> $O/${TAG}_sweep8.err || tail -5 $O/${TAG}_sweep8.err
cat $O/${TAG}_sweep8.json | cut -c1-900
