#!/bin/bash
# round-2 visit E (1 GPU, DEFAULT build): whole GPU suite incl. group / retriever / bulk-overwrite tests, ingest
# throughput of the rewritten norms kernel, bench lines of every single-GPU config, ncu captures for profiles/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -16 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== ingest bench"; timeout 600 python scripts/ingest_bench.py 4000000 768 2>&1 | tee gpurun_out/ingest_bench.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"convert_rows|row_norms" -c 24 --csv --log-file gpurun_out/launches_ingest.csv python scripts/ingest_bench.py 4000000 768 > gpurun_out/ncu_ingest_stdout.log 2>&1
echo "== bench cfg1"; timeout 600 python bench.py --workload cfg1 --steps 200 2>gpurun_out/bench_cfg1.err | tee gpurun_out/bench_cfg1.json
echo "== bench cfg2"; timeout 600 python bench.py --workload cfg2 --steps 50 2>gpurun_out/bench_cfg2.err | tee gpurun_out/bench_cfg2.json
echo "== bench default (cfg3 + extras)"; timeout 1200 python bench.py 2>gpurun_out/bench_cfg3.err | tee gpurun_out/bench_cfg3.json
tail -3 gpurun_out/bench_cfg3.err
echo "== ncu launch list cfg3"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extras > /dev/null 2>&1
echo "== ncu full scan cfg3 / cfg2"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:scan2 -s 3 -c 1 -f -o gpurun_out/r02_scan_cfg3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extras > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan2 -s 3 -c 1 -f -o gpurun_out/r02_scan_cfg2 python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extras > /dev/null 2>&1
ls -la gpurun_out | tail -12
