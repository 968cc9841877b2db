#!/bin/bash
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan -s 3 -c 1 -f -o gpurun_out/scan_small python bench.py --rows 131072 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_small.log 2>&1
tail -2 gpurun_out/ncu_small.log
