#!/bin/bash
# visit L: L2 prefetch of the corpus ahead of the smem ring in the pair kernel (EXPERIMENTAL build switches)
mkdir -p gpurun_out; rm -f gpurun_out/ab_prefetch.jsonl
export PYTHONUNBUFFERED=1
run() {
  wl=$1; rows=$2; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl $rows --no-extras --no-cpu-baseline --no-parity --steps 30 2>>gpurun_out/ab_pf.err \
    | python -c "import sys,json; j=json.loads(sys.stdin.read()); r=j['roofline']; print(json.dumps({'wl':'$wl $rows','env':'$*','value':round(j['value']),'kernel_ms':round(r['kernel_ms'],4),'frac':round(r['frac'],4)}))" | tee -a gpurun_out/ab_prefetch.jsonl
}
for pf in 0 1 2 4 8 0; do
  run cfg2 "" RBK_KNN_PREFETCH_TILES=$pf
  run cfg5 "" RBK_KNN_PREFETCH_TILES=$pf
  run cfg5 "--rows 625000" RBK_KNN_PREFETCH_TILES=$pf
done
run cfg3 "" RBK_KNN_PREFETCH_TILES=0
run cfg3 "" RBK_KNN_PREFETCH_TILES=2
tail -3 gpurun_out/ab_pf.err
