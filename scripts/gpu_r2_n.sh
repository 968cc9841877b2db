#!/bin/bash
# visit N: compute-sanitizer racecheck / synccheck / initcheck on small shapes of both scan kernels + finalize
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for tool in racecheck synccheck initcheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "255-64 or 4000-100-200 or threshold_inclusive" 2>&1 | tail -7 | tee gpurun_out/r02_sanitizer_$tool.log
done
