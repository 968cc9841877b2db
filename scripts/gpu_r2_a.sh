#!/bin/bash
# round-2 visit A: full GPU suite, smoke, default bench (cfg3 + extras), reference arm, 8-GPU-shard shapes on one GPU
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench default"; timeout 1200 python bench.py 2>gpurun_out/bench_cfg3.err | tee gpurun_out/bench_cfg3.json; tail -5 gpurun_out/bench_cfg3.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json
echo "== cfg5 shard of an 8-GPU run (625k rows)"; timeout 300 python bench.py --workload cfg5 --rows 625000 --no-cpu-baseline --no-parity --steps 50 2>gpurun_out/bench_cfg5s.err | tee gpurun_out/bench_cfg5_shard8.json
echo "== cfg3 shard of an 8-GPU run (1.25M rows)"; timeout 300 python bench.py --workload cfg3 --rows 1250000 --no-cpu-baseline --no-parity --steps 50 2>gpurun_out/bench_cfg3s.err | tee gpurun_out/bench_cfg3_shard8.json
echo "== ncu launch list cfg5 shard"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_cfg5_shard8.csv python bench.py --workload cfg5 --rows 625000 --steps 3 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_launch_stdout.log 2>&1
tail -40 gpurun_out/launches_cfg5_shard8.csv | cut -c1-220
ls -la gpurun_out
