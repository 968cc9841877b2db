#!/bin/bash
# interleaved A/B of an env knob on the headline bench: usage gpu_ab_cfg3.sh VAR a b
VAR=${1:-RBK_KNN_HALVES}; A=${2:-1}; B=${3:-2}
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
for v in $A $B $A $B; do
  env $VAR=$v timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$VAR=$v', round(j['value']), round(j['e2e']['value']), j['roofline']['kernel_ms'], j['clocks']['sm_mhz'], j['clocks']['reasons'])"
done
