#!/bin/bash
# visit O: CUDA-graph small-batch path - parity suite, cfg1 latency with / without (EXPERIMENTAL build for the switch), graph_replays
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python - <<'PY'
import numpy as np, time
from runbookai_b200 import Index, synth
c = synth.random_corpus(10000, 384, 1); q = synth.random_queries(1, 384, 2)
with Index(384) as ix:
    ix.append_bf16(c)
    for _ in range(20): ix.search(q, 10, None)
    t = time.perf_counter()
    for _ in range(2000): r = ix.search(q, 10, None)
    dt = (time.perf_counter() - t) / 2000
    print("warm-L2 search(B=1, 10k x 384): %.1f us per call, device %.1f us, graph_replays %d" % (dt * 1e6, r[3] * 1e3, ix.stats()["graph_replays"]))
    q8 = synth.random_queries(8, 384, 3)
    for _ in range(10): ix.search(q8, 10, 0.5)
    t = time.perf_counter()
    for _ in range(1000): r = ix.search(q8, 10, 0.5)
    print("B=8: %.1f us per call" % ((time.perf_counter() - t) / 1000 * 1e6), ix.stats()["graph_replays"])
PY
for g in 1 0; do RBK_KNN_GRAPH=$g timeout 300 python bench.py --workload cfg1 --steps 300 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('cfg1 graph=$g', 'value ms', round(j['ms_per_step'],4), 'e2e ms', round(j['e2e']['ms_per_step'],4), j['parity']['id_mismatch'])"; done
