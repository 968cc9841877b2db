#!/bin/bash
export PYTHONUNBUFFERED=1
run() { echo "== $*"; env "$@" timeout 600 python scripts/gpu_check.py time1 time2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except: continue
    print(j['stage'], 'b',j['b'],'n',j['n'], 'scan_ms', min(j['scan_ms'][1:]), 'tflops', round(j['tflops'],1), 'gbps', round(j['gbps']))
"; }
run RBK_KNN_TS=1
