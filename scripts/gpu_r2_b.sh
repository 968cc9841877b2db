#!/bin/bash
# round-2 visit B: cluster occupancy probe, ingest tests + K3 throughput (events + ncu launch list)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== cluster probe"; ./scripts/probes/cluster_probe 2>&1 | tee gpurun_out/cluster_probe.log
echo "== pytest (ingest/mutation)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bulk or mutation or keep_f64 or input_formats or round_to_nearest or logical_shards" 2>&1 | tail -8
echo "== ingest bench"; timeout 600 python scripts/ingest_bench.py 4000000 768 2>&1 | tee gpurun_out/ingest_bench.jsonl
echo "== ingest ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"convert_rows|row_norms" -c 24 --csv --log-file gpurun_out/launches_ingest.csv python scripts/ingest_bench.py 4000000 768 > gpurun_out/ncu_ingest_stdout.log 2>&1
grep -E "convert_rows|row_norms" gpurun_out/launches_ingest.csv | cut -d, -f5,13- | cut -c1-200 | tail -24
