#!/bin/bash
# round-2 visit D: cluster kernel + concurrent pair-kernel tail on the spare SMs.  EXPERIMENTAL build (A/B switches).
mkdir -p gpurun_out; rm -f gpurun_out/ab_tail.jsonl
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu (hybrid on)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_tail.log
run() {  # workload, extra env...
  wl=$1; shift
  env "$@" timeout 600 python bench.py --workload $wl --no-extras --no-cpu-baseline --no-parity --steps 20 2>>gpurun_out/ab_tail.err \
    | python -c "import sys,json; j=json.loads(sys.stdin.read()); r=j['roofline']; print(json.dumps({'wl':'$wl','env':'$*','value':round(j['value']),'e2e':round(j['e2e']['value']),'kernel_ms':round(r['kernel_ms'],4),'frac':round(r['frac'],4),'clk':j['clocks']['sm_mhz'],'pw':j['clocks']['power_w']}))" | tee -a gpurun_out/ab_tail.jsonl
}
for wl in cfg3 cfg2 cfg5; do
  run $wl RBK_KNN_TAIL_PAIRS=0
  run $wl RBK_KNN_TAIL_RHO=1.05
  run $wl RBK_KNN_TAIL_RHO=1.15
  run $wl RBK_KNN_TAIL_RHO=1.30
  run $wl RBK_KNN_CLUSTER4=0
done
run cfg3 RBK_KNN_TAIL_RHO=1.15
run cfg3 RBK_KNN_TAIL_PAIRS=0
tail -5 gpurun_out/ab_tail.err
echo "== cfg4 shard"
for e in RBK_KNN_TAIL_PAIRS=0 RBK_KNN_TAIL_RHO=1.15 RBK_KNN_CLUSTER4=0; do
env $e timeout 600 python bench.py --workload cfg4 --rows 6250000 --no-extras --no-cpu-baseline --no-parity --steps 10 2>>gpurun_out/ab_tail.err \
  | python -c "import sys,json; j=json.loads(sys.stdin.read()); r=j['roofline']; print(json.dumps({'wl':'cfg4shard','env':'$e','value':round(j['value']),'kernel_ms':round(r['kernel_ms'],4),'frac':round(r['frac'],4),'clk':j['clocks']['sm_mhz']}))" | tee -a gpurun_out/ab_tail.jsonl
done
echo "== ncu launch list (hybrid, cfg3): do the two kernels overlap?"
timeout 600 nsys --version >/dev/null 2>&1 || true
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan" -c 12 --csv --log-file gpurun_out/launches_hybrid_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extras > /dev/null 2>&1
grep -E "scan" gpurun_out/launches_hybrid_cfg3.csv | cut -d, -f5,8,9,15 | cut -c1-160 | tail -8
