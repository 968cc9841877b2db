#!/bin/bash
# short GPU visit: parity tests + scan-time sweeps + the two bench lines (no ncu)
mkdir -p gpurun_out; rm -f gpurun_out/check.jsonl
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 300 python scripts/gpu_check.py sweep256 sweep 2>&1 | tail -14
timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/bench_cfg3.err | tee gpurun_out/bench_cfg3.json
timeout 300 python bench.py --workload cfg2 --steps 50 --no-cpu-baseline 2>gpurun_out/bench_cfg2.err | tee gpurun_out/bench_cfg2.json
