#!/bin/bash
# visit K (4 GPUs): the driver's SCALE shape at N=4 (both arms under torchrun)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
echo "== reference arm under torchrun"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>gpurun_out/bench_ref_g$N.err | grep "^{" | cut -c1-400; echo "rc=$?"
echo "== our arm"
timeout 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 20 2>gpurun_out/bench_g$N.err | grep "^{" | tee gpurun_out/bench_cfg3_g$N.json | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print('cfg3', round(j['value']), round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']), 'parity', j['parity']['id_mismatch'], j['parity']['score_mismatch'], 'frac', round(j['roofline']['frac'],3))
for k, e in j.get('extra_workloads', {}).items(): print(k, round(e['value']), round(e['ms_per_step'],4), 'e2e', round(e['e2e']['value']), 'parity', e['parity']['id_mismatch'], e['parity']['score_mismatch'])"
tail -2 gpurun_out/bench_g$N.err
