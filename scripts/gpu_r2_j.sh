#!/bin/bash
# visit J: finalize phase timing probe (build with RBK_EXTRA_NVCC_FLAGS=-DRBK_FIN_PROFILE) + parity
export PYTHONUNBUFFERED=1
timeout 300 python bench.py --workload cfg5 --rows 625000 --steps 2 --warmup 3 --no-cpu-baseline --no-parity 2>&1 | grep "\[fin\]" | tail -4
timeout 300 python bench.py --workload cfg1 --steps 2 --warmup 3 --no-cpu-baseline --no-parity 2>&1 | grep "\[fin\]" | tail -2
timeout 300 python bench.py --workload cfg3 --rows 1250000 --steps 2 --warmup 3 --no-cpu-baseline --no-parity 2>&1 | grep "\[fin\]" | tail -2
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | grep -v "\[fin\]" | tail -5
