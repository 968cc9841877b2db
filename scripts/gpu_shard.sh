#!/bin/bash
# per-GPU behaviour of an 8-way shard of cfg3 on ONE gpu: 1.25M rows, B=1024
export PYTHONUNBUFFERED=1
timeout 600 python bench.py --rows 1250000 --steps 50 --no-cpu-baseline | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('shard 1.25M: q/s', round(j['value']), 'ms/step', round(j['ms_per_step'],3), 'scan_ms', round(j['roofline']['kernel_ms'],3), 'e2e ms', round(j['e2e']['ms_per_step'],3), j['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan|finalize_kernel|prep_queries" -c 30 --csv --log-file gpurun_out/launches_shard.csv python bench.py --rows 1250000 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_shard.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows: agg[r[4].split('(')[0][-40:]].append(float(r[-1]))
for k,v in agg.items(): print(f"{k:42s} n={len(v):3d} avg_us={sum(v)/len(v)/1e3:9.1f}")
PY
