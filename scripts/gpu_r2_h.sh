#!/bin/bash
# round-2 visit H (1 GPU, default build): validation of the final tree - whole GPU suite, smoke, memcheck of the new
# code paths, fixed-cost launch list, default bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== compute-sanitizer memcheck (bulk overwrite, group, mutation, tie fallback, pair kernel shapes)"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bulk or group or mutation or exact_ties or threshold_inclusive or 4000-100-200 or 255-64" 2>&1 | tail -6 | tee gpurun_out/r02_sanitizer_memcheck.log
echo "== launch list, cfg5 8-GPU shard shape"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_cfg5_shard8_v2.csv python bench.py --workload cfg5 --rows 625000 --steps 3 --warmup 3 --no-cpu-baseline --no-parity > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_cfg5_shard8_v2.csv')))
agg = collections.defaultdict(list)
for r in rows:
    if len(r) < 15 or not r[0].isdigit(): continue
    agg[r[4].split('(')[0][-36:]].append(float(r[-1]))
for k, v in agg.items(): print(f"{k:38s} n={len(v):3d} mean={sum(v)/len(v)/1000:8.2f} us")
PY
echo "== cfg5 shard / cfg1 / default bench"
timeout 300 python bench.py --workload cfg5 --rows 625000 --no-cpu-baseline --no-parity --steps 50 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('cfg5 shard', j['ms_per_step'], j['roofline']['kernel_ms'], j['e2e']['ms_per_step'])"
timeout 300 python bench.py --workload cfg1 --steps 200 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('cfg1', j['ms_per_step'], j['roofline']['kernel_ms'], j['e2e']['ms_per_step'], j['parity'])"
timeout 1200 python bench.py 2>gpurun_out/bench_cfg3.err | tee gpurun_out/bench_cfg3_final.json | cut -c1-400
tail -3 gpurun_out/bench_cfg3.err
