#!/bin/bash
# One GPU box visit at the end of a round: GPU test suite, smoke, bench (both arms), the bench's launch list under ncu, and one
# ncu --set full capture of the dominant kernel.  Usage (here): gpurun --timeout 900 -- 'bash tools/gpu_round_end.sh TAG'
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q > $O/${TAG}_gputest.log 2>&1; tail -3 $O/${TAG}_gputest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
timeout 300 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err || { tail -5 $O/${TAG}_bench.err; exit 1; }
timeout 200 python bench.py --impl reference > $O/${TAG}_ref.json 2> $O/${TAG}_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${TAG}_bench_launches.csv \
    python bench.py --steps 10 --warmup 3 --no-other-configs --no-cpu-baseline --no-dropin --latency-ticks 10 > $O/${TAG}_ncu_bench.log 2>&1
python tools/prof_tick.py > $O/${TAG}_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:riccati_kernel --launch-skip 85 --launch-count 1 \
    -o $O/${TAG}_riccati -f python tools/prof_tick.py > $O/${TAG}_ncu_full.log 2>&1
tail -2 $O/${TAG}_ncu_full.log | cut -c1-200
python - $TAG <<'P'
import json,sys
d=json.load(open("gpurun_out/%s_bench.json" % sys.argv[1] if len(sys.argv)>1 else "gpurun_out/r02_bench.json"))
print(d["value"], d["value_one_tick_at_a_time"], d["e2e"]["value"], d["roofline"]["frac"], {k[:12]: round(v["value"]/1e6,2) for k,v in d["other_configs"].items() if isinstance(v,dict)})
P
ls -la $O | grep $TAG
