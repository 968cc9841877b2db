#!/bin/bash
# final check of the shipped (default) build: what the driver runs at round end
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-260
echo "== bench"; timeout 1200 python bench.py 2>gpurun_out/bench_final.err | tee gpurun_out/bench_final.json | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print('cfg3', round(j['value']), round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']), 'parity', j['parity']['id_mismatch'], j['parity']['score_mismatch'], j['parity']['queries'], 'frac', round(j['roofline']['frac'],3), 'launches', j['gpu_launches'], j['clocks'])
print('cpu_baseline', round(j['cpu_baseline']['value'],3), j['cpu_baseline']['cores'], 'ref_1T', round(j['cpu_baseline']['ref_1T']['value'],4))
for k, e in j.get('extra_workloads', {}).items(): print(k, round(e['value']), round(e['ms_per_step'],4), 'e2e', round(e['e2e']['value']), 'parity', e['parity']['id_mismatch'], e['parity']['score_mismatch'])"
tail -2 gpurun_out/bench_final.err
