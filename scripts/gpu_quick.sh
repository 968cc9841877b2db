#!/bin/bash
# quick GPU visit: parity tests + a few timing stages + bench lines (no ncu)
mkdir -p gpurun_out; rm -f gpurun_out/check.jsonl
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
timeout 600 python scripts/gpu_check.py time1 time2 time3 2>&1 | tail -5
timeout 900 python bench.py --no-cpu-baseline 2>gpurun_out/bench_cfg3.err | tee gpurun_out/bench_cfg3.json
timeout 600 python bench.py --workload cfg2 --steps 50 --no-cpu-baseline 2>gpurun_out/bench_cfg2.err | tee gpurun_out/bench_cfg2.json
