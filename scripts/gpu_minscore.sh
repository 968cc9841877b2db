#!/bin/bash
# headline workloads with the reference's default threshold (minScore 0.5), then the per-GPU view of an 8-way shard
export PYTHONUNBUFFERED=1; mkdir -p gpurun_out
for wl in cfg3 cfg2; do
  timeout 600 python bench.py --workload $wl --min-score 0.5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_${wl}_min05.json | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$wl min_score 0.5: q/s', round(j['value']), 'e2e', round(j['e2e']['value']), 'scan_ms', round(j['roofline']['kernel_ms'],3), 'frac', round(j['roofline']['frac'],3), j['clocks']['sm_mhz'], j['clocks']['reasons'])"
done
bash scripts/gpu_shard.sh
