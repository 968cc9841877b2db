#!/bin/bash
# ncu --set full capture of the scan kernel for one workload: scripts/gpu_ncu.sh cfg3 [tag]
WL=${1:-cfg3}; TAG=${2:-$WL}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"scan" -s 3 -c 1 -f -o gpurun_out/scan_$TAG python bench.py --workload $WL --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_${TAG}_stdout.log 2>&1
tail -3 gpurun_out/ncu_${TAG}_stdout.log
