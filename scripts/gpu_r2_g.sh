#!/bin/bash
# round-2 visit G (8 GPUs): NCCL parity at 8 ranks, the SCALE-shaped bench (cfg3 + cfg5 + cfg4 extras), rbk_group on 8 GPUs
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
echo "== dist_check ($N ranks)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 scripts/dist_check.py 2>gpurun_out/dist_check_$N.err | grep -v "^NCCL" | tee gpurun_out/dist_check_$N.log
echo "== bench --gpus $N"
timeout 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 20 2>gpurun_out/bench_g$N.err | grep "^{" | tee gpurun_out/bench_cfg3_g$N.json | cut -c1-600
tail -3 gpurun_out/bench_g$N.err
echo "== rbk_group, one process, $N GPUs"
timeout 900 python scripts/group_check.py 2000000 768 1024 32 2>gpurun_out/group_check.err | tee gpurun_out/group_check_g$N.json
timeout 900 python scripts/group_check.py 2000000 768 256 16 2>>gpurun_out/group_check.err | tee -a gpurun_out/group_check_g$N.json
tail -3 gpurun_out/group_check.err
