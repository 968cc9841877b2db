#!/bin/bash
# multi-GPU visit: sharded parity vs oracle + bench at N GPUs.   usage: gpu_multi.sh N [quick]
G=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR scripts/dist_check.py 2>&1 | grep -v "^W\|^\*\*\*\|Setting OMP" | tail -4 | tee gpurun_out/dist_check_$G.log
timeout 900 $TR bench.py --gpus $G --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_g$G.err | grep '^{' | tee gpurun_out/bench_g$G.json | cut -c1-300
if [ "$2" != "quick" ]; then
timeout 900 $TR bench.py --gpus $G --workload cfg5 --steps 30 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench_g$G.err | grep '^{' | tee gpurun_out/bench_cfg5_g$G.json | cut -c1-300
fi
