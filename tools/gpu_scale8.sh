#!/bin/bash
# One 8-GPU box visit (gpurun --gpus 8 --timeout 600 -- 'bash tools/gpu_scale8.sh TAG'):
#   1. host <-> device copy bandwidth with 1 / 2 / 4 / 8 GPUs copying at once (tools/copy_concurrent.cu) + the box's topology,
#   2. the headline bench at 1, 2, 4, 8 GPUs as the driver launches it,
#   3. BASELINE configs[4]: 1 048 576 robots = 8 x 131 072, 500 closed-loop ticks on the device.
TAG=${1:-r02}
O=gpurun_out; mkdir -p $O
NG=$(nvidia-smi -L | wc -l)
{ echo "== nproc $(nproc)"; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core"; echo "== nvidia-smi topo -m"; nvidia-smi topo -m; } > $O/${TAG}_host_topology.txt 2>&1
./tools/copy_concurrent > $O/${TAG}_copy_concurrent.jsonl 2> $O/${TAG}_copy_concurrent.err
cat $O/${TAG}_copy_concurrent.jsonl
python bench.py --gpus 1 --steps 50 --warmup 5 --no-other-configs --no-cpu-baseline > $O/${TAG}_bench1.json 2> $O/${TAG}_bench1.err || tail -5 $O/${TAG}_bench1.err
for n in 2 4 8; do
  [ $n -le $NG ] || continue
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29510 + n)) \
      bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench${n}.json 2> $O/${TAG}_bench${n}.err || tail -5 $O/${TAG}_bench${n}.err
done
python - <<PY
import json
for n in (1, 2, 4, 8):
    try:
        d = json.load(open("$O/${TAG}_bench%d.json" % n))
    except Exception as e:
        print(n, "missing", e); continue
    lw = d["latency_window"]
    print("gpus %d: device %.2f M/s (%.4f ms)  e2e %.2f M/s (%.4f ms)  dropin %.2f M/s  window: dev p50 %.3f p99 %.3f  e2e p50 %.3f p99 %.3f max %.3f" % (
        n, d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"], d["e2e_dropin"]["value"] / 1e6,
        lw["device_ms_p50"], lw["device_ms_p99"], lw["e2e_ms_p50"], lw["e2e_ms_p99"], lw["e2e_ms_max"]))
PY
if [ $NG -ge 8 ]; then
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
      bench.py --gpus 8 --workload sweep --batch 131072 --steps 500 --warmup 5 > $O/${TAG}_sweep8.json 2> $O/${TAG}_sweep8.err || tail -5 $O/${TAG}_sweep8.err
  cut -c1-400 $O/${TAG}_sweep8.json
fi
