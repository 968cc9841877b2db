#!/bin/bash
# per-launch device times (ncu, cold-cache/serialised: compare shares) for a workload
WL=${1:-cfg3}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan|finalize_kernel|prep_queries|exact_|merge_shards" -c 60 --csv --log-file gpurun_out/launches_$WL.csv python bench.py --workload $WL --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$WL.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_$WL.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows: agg[r[4].split('(')[0][-40:]].append(float(r[-1]))
for k,v in agg.items(): print(f"{k:42s} n={len(v):3d} avg_us={sum(v)/len(v)/1e3:9.1f} min_us={min(v)/1e3:9.1f}")
PY
