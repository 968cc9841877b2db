#!/bin/bash
# visit M (8 GPUs): the SCALE-shaped line again with the final tree (fixed-cost cuts), cfg5 as main workload too
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
timeout 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 20 2>gpurun_out/bench_g$N.err | grep "^{" | tee gpurun_out/bench_cfg3_g${N}_final.json | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print('cfg3', round(j['value']), round(j['ms_per_step'],4), 'e2e', round(j['e2e']['value']), 'parity', j['parity']['id_mismatch'], j['parity']['score_mismatch'], 'kernel', round(j['roofline']['kernel_ms'],4), 'frac', round(j['roofline']['frac'],3))
for k, e in j.get('extra_workloads', {}).items(): print(k, round(e['value']), round(e['ms_per_step'],4), 'e2e', round(e['e2e']['value']), 'parity', e['parity']['id_mismatch'], e['parity']['score_mismatch'], 'kernel', round(e['roofline']['kernel_ms'],4))"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $N --workload cfg5 --steps 50 --no-cpu-baseline 2>>gpurun_out/bench_g$N.err | grep "^{" | tee gpurun_out/bench_cfg5_g${N}_final.json | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('cfg5 main', round(j['value']), round(j['ms_per_step'],4), 'e2e', round(j['e2e']['value']), j['parity']['id_mismatch'], round(j['roofline']['kernel_ms'],4))"
timeout 600 python bench.py --workload cfg5 --steps 50 --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('cfg5 1 GPU same box', round(j['value']), round(j['ms_per_step'],4), 'e2e', round(j['e2e']['value']))"
timeout 600 python bench.py --no-extras --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read()); print('cfg3 1 GPU same box', round(j['value']), round(j['ms_per_step'],4), 'e2e', round(j['e2e']['value']), round(j['roofline']['frac'],3))"
tail -2 gpurun_out/bench_g$N.err
