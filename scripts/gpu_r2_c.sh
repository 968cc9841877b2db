#!/bin/bash
# round-2 visit C: the cluster-multicast scan kernel (rbk_scan4.cu): parity suite, then A/B against the pair kernel.
# Needs the EXPERIMENTAL build (RBK_EXPERIMENTAL=1 python -m runbookai_b200.build --force) for the RBK_KNN_CLUSTER4 switch.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu (cluster kernel on)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_c4.log
for wl in cfg3 cfg2 cfg5; do
  for c4 in 1 0 1 0; do
    echo "== bench $wl cluster4=$c4"
    RBK_KNN_CLUSTER4=$c4 timeout 600 python bench.py --workload $wl --no-extras --no-cpu-baseline --no-parity --steps 20 2>>gpurun_out/ab_c4.err \
      | python -c "import sys,json; j=json.loads(sys.stdin.read()); r=j['roofline']; print(json.dumps({'wl':'$wl','c4':$c4,'value':round(j['value']),'e2e':round(j['e2e']['value']),'kernel_ms':round(r['kernel_ms'],4),'frac':round(r['frac'],4),'clk':j['clocks']['sm_mhz'],'reasons':j['clocks']['reasons']}))" | tee -a gpurun_out/ab_c4.jsonl
  done
done
tail -5 gpurun_out/ab_c4.err
echo "== cfg4 shard (6.25M x 1024, B=512) c4=1/0"
for c4 in 1 0; do
RBK_KNN_CLUSTER4=$c4 timeout 600 python bench.py --workload cfg4 --rows 6250000 --no-extras --no-cpu-baseline --no-parity --steps 10 2>>gpurun_out/ab_c4.err \
  | python -c "import sys,json; j=json.loads(sys.stdin.read()); r=j['roofline']; print(json.dumps({'wl':'cfg4shard','c4':$c4,'value':round(j['value']),'kernel_ms':round(r['kernel_ms'],4),'frac':round(r['frac'],4),'clk':j['clocks']['sm_mhz']}))" | tee -a gpurun_out/ab_c4.jsonl
done
echo "== ncu full (scan4, cfg3)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:scan4 -s 3 -c 1 -f -o gpurun_out/r02_scan4_cfg3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu_full_c4_stdout.log 2>&1
ls -la gpurun_out | tail -8
