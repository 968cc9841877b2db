#!/bin/bash
# parity tests with the automatic seeding mode, then an interleaved A/B of per-chunk (0) vs per-tile (1) seeds
export PYTHONUNBUFFERED=1; mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 0 1 0 1; do
  echo "== RBK_KNN_SEED_TILE=$v"
  RBK_KNN_SEED_TILE=$v timeout 120 python scripts/gpu_check.py sweep256 sweep1 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    if j['n'] in (65536, 1048576): print(j['b'], j['n'], round(j['scan_ms_min'],4), round(j['scan_ms_med'],4))"
done
