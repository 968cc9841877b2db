#!/bin/bash
# round-2 visit I: where does the finalize kernel's time go?  ncu --set full with source view on the 8-GPU cfg5 shard shape
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:finalize_kernel -s 4 -c 1 -f -o gpurun_out/r02_finalize_cfg5s python bench.py --workload cfg5 --rows 625000 --steps 2 --warmup 3 --no-cpu-baseline --no-parity > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prep_queries -s 4 -c 1 -f -o gpurun_out/r02_prep_cfg5s python bench.py --workload cfg5 --rows 625000 --steps 2 --warmup 3 --no-cpu-baseline --no-parity > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
