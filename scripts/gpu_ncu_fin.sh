#!/bin/bash
timeout 600 ncu --set full --clock-control none --import-source on -k regex:finalize -s 3 -c 1 -f -o gpurun_out/finalize_cfg2 python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_fin.log 2>&1
tail -1 gpurun_out/ncu_fin.log
