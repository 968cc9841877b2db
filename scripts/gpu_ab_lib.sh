#!/bin/bash
# interleaved A/B of two builds of the library (ab/base.so, ab/new.so) on one box: parity suite on the new one,
# then the headline bench alternating
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
cp ab/new.so runbookai_b200/lib/librbk_knn.so
echo "== parity (new)"; timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
for v in base new base new base new; do
  cp ab/$v.so runbookai_b200/lib/librbk_knn.so
  timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
j=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1])
x=j.get('extra_workloads',{})
print('$v', round(j['value']), 'e2e', round(j['e2e']['value']), 'k_ms', j['roofline']['kernel_ms'], 'par', j['parity']['id_mismatch'], j['parity']['score_mismatch'], j['clocks']['sm_mhz'], j['clocks']['reasons'], {k: round(e['value']) for k,e in x.items()})"
done
cp ab/new.so runbookai_b200/lib/librbk_knn.so
for v in base new base new; do
  cp ab/$v.so runbookai_b200/lib/librbk_knn.so
  echo "-- $v small-batch"; timeout 300 python scripts/small_batch_sweep.py 2>&1 | tail -6
done
