#!/bin/bash
# One GPU-box visit: tests, smoke, bench lines, ncu launch list + full capture of the scan kernel.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench cfg3"; timeout 900 python bench.py 2>gpurun_out/bench_cfg3.err | tee gpurun_out/bench_cfg3.json
echo "== bench cfg2"; timeout 600 python bench.py --workload cfg2 --steps 50 2>gpurun_out/bench_cfg2.err | tee gpurun_out/bench_cfg2.json
echo "== bench cfg1"; timeout 600 python bench.py --workload cfg1 --steps 200 2>gpurun_out/bench_cfg1.err | tee gpurun_out/bench_cfg1.json
echo "== bench cfg5"; timeout 600 python bench.py --workload cfg5 --steps 30 --no-cpu-baseline 2>gpurun_out/bench_cfg5.err | tee gpurun_out/bench_cfg5.json
echo "== scan timings"; timeout 300 python scripts/gpu_check.py time1 time2 time3 sweep1 2>&1 | tail -12 | tee gpurun_out/timings.log
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json
if [ "$1" != "noncu" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan|finalize_kernel|prep_queries|exact_|merge_shards" -c 200 --csv --log-file gpurun_out/launches_cfg3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_stdout.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan|finalize_kernel|prep_queries|exact_|merge_shards" -c 200 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch2_stdout.log 2>&1
echo "== ncu full (scan kernel, cfg3)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:scan -s 3 -c 1 -f -o gpurun_out/scan_cfg3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_stdout.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan -s 3 -c 1 -f -o gpurun_out/scan_cfg2 python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full2_stdout.log 2>&1
fi
ls -la gpurun_out
