#!/bin/bash
# One GPU box visit: plain bench, in-kernel phase timers, ncu launch list, one ncu --set full capture
# of a steady-state active-set launch.  Usage (here): gpurun --timeout 900 -- 'bash tools/gpu_profile.sh TAG'
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
python bench.py --steps 50 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err || { tail -20 $O/${TAG}_bench.err; exit 1; }
cat $O/${TAG}_bench.json

ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --settle 14 --no-cpu-baseline > $O/${TAG}_ncu_list.log 2>&1
# launches of riccati_kernel: 2 per tick of the (host-input, chunked) generation loop = 42, then one per device-resident tick;
# skip 59 -> the capture is tick 17 of the timed device-resident replay (steady state, full 4096-robot launch)
ncu --set full --clock-control none --import-source on -k regex:riccati_kernel --launch-skip 59 --launch-count 1 \
    -o $O/${TAG}_solve -f python bench.py --steps 4 --warmup 3 --settle 14 --no-cpu-baseline > $O/${TAG}_ncu_full.log 2>&1
tail -3 $O/${TAG}_ncu_full.log | cut -c1-300
ls -la $O
