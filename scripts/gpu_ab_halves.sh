#!/bin/bash
# A/B of the epilogue width (4 vs 8 warps per CTA) over scan lengths, interleaved so both see the same box state
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
for h in 1 2 1 2; do
  echo "== RBK_KNN_HALVES=$h"
  RBK_KNN_HALVES=$h timeout 300 python scripts/gpu_check.py sweep256 sweep 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print(j['b'], j['n'], round(j['scan_ms_min'],4), round(j['scan_ms_med'],4))"
done
