#!/bin/bash
export PYTHONUNBUFFERED=1
run() { env "$@" timeout 600 python bench.py --no-cpu-baseline --steps 30 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('$*', 'cfg3', round(j['value']), 'scan_ms', round(j['roofline']['kernel_ms'],2), j['clocks']['sm_mhz'])"; env "$@" timeout 600 python scripts/gpu_check.py time1 time2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except: continue
    if 'scan_ms' in j: print('   ', j['stage'], 'b',j['b'],'n',j['n'], 'scan_ms', min(j['scan_ms'][1:]), 'tflops', round(j['tflops'],1))
"; }
for rep in 1 2; do
run RBK_KNN_MAX_LEAD=1
run RBK_KNN_MAX_LEAD=2
run RBK_KNN_MAX_LEAD=3
done
