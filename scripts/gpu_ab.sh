#!/bin/bash
export PYTHONUNBUFFERED=1
run() { echo "== $*"; env "$@" timeout 600 python bench.py --no-cpu-baseline --steps 30 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('cfg3', round(j['value']), 'scan_ms', round(j['roofline']['kernel_ms'],2), j['clocks'])"; }
run RBK_KNN_TS=0
run RBK_KNN_HYBRID_KB=6 RBK_KNN_HYBRID_SLOTS=8
run RBK_KNN_HYBRID_KB=8 RBK_KNN_HYBRID_SLOTS=6
run RBK_KNN_TS=1
run RBK_KNN_RESIDENT=1
