#!/bin/bash
# cfg4 (50M x 1024, B=512, k=32) needs 8 GPUs; this runs ONE GPU's 6.25M-row shard of it alone
export PYTHONUNBUFFERED=1; mkdir -p gpurun_out
timeout 600 python bench.py --workload cfg4 --rows 6250000 --steps 20 --no-cpu-baseline 2>gpurun_out/bench_cfg4_shard.err | tee gpurun_out/bench_cfg4_shard.json | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('cfg4 shard 6.25M x 1024, B=512: q/s', round(j['value']), 'ms/step', round(j['ms_per_step'],3), 'scan_ms', round(j['roofline']['kernel_ms'],3), 'TF', round(j['roofline'].get('achieved',0)), j['roofline']['bound'], round(j['roofline']['frac'],3), 'e2e', round(j['e2e']['value']), j['clocks'])"
tail -3 gpurun_out/bench_cfg4_shard.err
