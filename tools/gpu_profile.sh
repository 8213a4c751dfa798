#!/bin/bash
# One GPU box visit: the profiling driver plain (must exit 0), its ncu launch list, one ncu --set full capture (with source) of a
# steady-state launch of the dominant kernel.  Usage (here): gpurun --timeout 900 -- 'bash tools/gpu_profile.sh TAG'
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
python tools/prof_tick.py > $O/${TAG}_prof_plain.log 2>&1 || { tail -20 $O/${TAG}_prof_plain.log; exit 1; }
cat $O/${TAG}_prof_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $O/${TAG}_launches.csv \
    python tools/prof_tick.py > $O/${TAG}_ncu_list.log 2>&1
python tools/prof_tick.py > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:riccati_kernel --launch-skip 85 --launch-count 1 \
    -o $O/${TAG}_riccati -f python tools/prof_tick.py > $O/${TAG}_ncu_full.log 2>&1
tail -3 $O/${TAG}_ncu_full.log | cut -c1-300
ls -la $O | grep $TAG
