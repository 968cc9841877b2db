#!/bin/bash
# round-2 visit F (2 GPUs): the real NCCL paths - torchrun sharded search (2 ranks) and the in-library rbk_group -
# then the norms-kernel variants and a 2-GPU bench line.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L
echo "== pytest: NCCL + group + shard tests"; timeout 1500 python -m pytest tests -m gpu -x -q -k "two_rank or group or logical_shards or multi_device or retriever or sidecar or bulk" 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_2gpu.log
echo "== norms kernel variants (0 = staged, 1 = +register prefetch, 2 = +integer squares, 3 = integer squares only)"
for v in 0 1 2 3; do RBK_NORMS_VARIANT=$v timeout 300 python scripts/ingest_bench.py 4000000 768 2>&1 | grep bf16 | sed "s/^/variant $v: /"; done | tee gpurun_out/norms_variants.log
echo "== bench --gpus 2 (cfg3 + extras)"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 20 2>gpurun_out/bench_g2.err | tee gpurun_out/bench_cfg3_g2.json | cut -c1-1800
tail -3 gpurun_out/bench_g2.err
