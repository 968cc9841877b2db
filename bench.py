#!/usr/bin/env python
"""bench.py — kNN queries/sec of the knowledge-base search path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--impl ours|reference]

A "step" is one pass of the hot path (batched cosine scan + exact top-k) over one batch of
B synthetic queries against the resident synthetic corpus.  Workloads are BASELINE.json's
configs; the default is the one the target is quoted on (10M x 768 bf16, B=1024, k=16).
For N>1 (torchrun, one rank per GPU) the SAME corpus is row-sharded across the ranks
(strong scaling) and each step ends with the NCCL all-gather + merge of the per-rank lists.

Prints ONE JSON line (rank 0).  `value` = queries/s with the batch already resident in
HBM; `e2e` = the same through the host-facing C-ABI call (host query buffer in, host
results out, copies inside the timed region).  `--impl reference` times the reference's
CPU algorithm (oracle/, strict fp64 restatement; Node is absent so the TypeScript itself
cannot run) on the box's host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (N_docs, d, B, k, description)
    "cfg1": (10_000, 384, 1, 5, "10k x 384, B=1, k=5 (reference-scale, latency-bound)"),
    "cfg2": (1_000_000, 768, 256, 16, "1M x 768 bf16, B=256, k=16"),
    "cfg3": (10_000_000, 768, 1024, 16, "10M x 768 bf16, B=1024, k=16"),
    "cfg4": (50_000_000, 1024, 512, 32, "50M x 1024 bf16, B=512, k=32 (needs 8 GPUs)"),
    "cfg5": (5_000_000, 768, 256, 8, "5M x 768 bf16, 32 investigations x 8 queries, k=8"),
}
SEED = 0x5EED0003
GEN_CHUNK = 65536


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return dict(hbm=j["hbm_gbs"], tf=j["bf16_tflops"], tf_sus=j.get("bf16_tflops_sustained"), src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([time.monotonic()] + [x.strip() for x in line.split(",")])

    def window(self, t0: float, t1: float):
        """Keep only the samples taken inside [t0, t1] (monotonic clock); if the window was shorter than the
        sampling period, the sample nearest to it."""
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        if not inside and self.rows:
            inside = [min(self.rows, key=lambda r: min(abs(r[0] - t0), abs(r[0] - t1)))]
        self.rows = inside

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        rows = [r[1:] for r in self.rows]   # drop the timestamp
        self.rows = rows
        sm = [float(r[1]) for r in rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6),
                          ("sw_power_cap", 7)):
            if any(len(r) >= 8 and r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def gen_shard(ix, lo: int, hi: int, d: int, device):
    """Fill `ix` with global rows [lo, hi): N(0,1) rounded to bf16, generated on the device
    chunk by chunk from a per-chunk seed, so any sharding yields the same global corpus."""
    import torch
    c0 = lo // GEN_CHUNK
    c1 = -(-hi // GEN_CHUNK)
    g = torch.Generator(device=device)
    for c in range(c0, c1):
        g.manual_seed(SEED * 1_000_003 + c)
        t = torch.randn(GEN_CHUNK, d, device=device, generator=g, dtype=torch.float32).to(torch.bfloat16)
        a = max(lo, c * GEN_CHUNK) - c * GEN_CHUNK
        b = min(hi, (c + 1) * GEN_CHUNK) - c * GEN_CHUNK
        t = t[a:b].contiguous()
        torch.cuda.synchronize(device)
        ix.append_bf16_device(t.data_ptr(), b - a)
        del t


def cpu_reference_qps(corpus_bits, queries, k_fetch, n_total, threads):
    """Oracle, rows split over all host cores; returns (qps scaled to n_total rows, seconds)."""
    import oracle
    t0 = time.perf_counter()
    oracle.search_batch_mt(corpus_bits, queries.astype(np.float64), k_fetch, None, n_threads=threads)
    dt = time.perf_counter() - t0
    return queries.shape[0] / dt * (corpus_bits.shape[0] / n_total), dt


def run_reference(args, wl):
    """--impl reference: the reference's CPU algorithm on this box's cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from runbookai_b200 import synth  # numpy-only helpers; no CUDA involved on this arm
    import oracle
    oracle.build()
    n, d, B, k, desc = wl
    cores = os.cpu_count() or 1
    ns = min(n, 100_000)
    bs = min(B, max(8, cores // 2))
    corpus = synth.random_corpus(ns, d, SEED)
    queries = synth.random_queries(bs, d, SEED + 1)
    for _ in range(args.warmup):
        cpu_reference_qps(corpus, queries, 2 * k, n, cores)
    t_all, qps = 0.0, []
    for _ in range(args.steps):
        v, dt = cpu_reference_qps(corpus, queries, 2 * k, n, cores)
        qps.append(v)
        t_all += dt
    value = float(np.mean(qps))
    sample = f"{ns} rows x {bs} queries per step, all {cores} host threads, scaled linearly to {n} rows"
    print(json.dumps({
        "impl": "reference", "metric": "knn_queries_per_sec", "value": value, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_all / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "n_docs": n, "dim": d, "batch": B, "k": k, "k_fetch": 2 * k},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("RBK_BENCH_WORKLOAD", "cfg3"), choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override N_docs (debug only; marks the run reduced)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--min-score", type=float, default=None,
                    help="cosine threshold (the reference's default is 0.5); default: none (-inf), the headline runs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    wl = list(WORKLOADS[args.workload])
    reduced = bool(args.rows)
    if args.rows:
        wl[0] = args.rows
    if args.impl == "reference":
        run_reference(args, wl)
        return
    n, d, B, k, desc = wl
    k_fetch = 2 * k   # the reference's over-fetch (vector-store.ts:221)

    import torch
    import torch.distributed as dist
    from runbookai_b200 import Index, synth
    from runbookai_b200.sharded import ShardedSearcher, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this engine has no CPU path")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()   # nvidia-smi takes a second to start: launch it now, keep only the timed region's samples
    lo, hi = shard_bounds(n, world, rank)
    ix = Index(d, device=local, capacity_hint=hi - lo)
    ix.set_slot_base(lo)
    gen_shard(ix, lo, hi, d, device)
    stream = torch.cuda.Stream(device)  # a real (non-legacy) stream shared by torch, NCCL and the engine
    torch.cuda.set_stream(stream)
    ix.set_stream(stream.cuda_stream)   # so torch CUDA events bracket the engine's kernels
    searcher = ShardedSearcher(ix)

    q_host = torch.from_numpy(synth.random_queries(B, d, SEED + 1)).pin_memory()
    q_dev = q_host.to(device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---------------- value: batch resident in HBM ----------------
    for _ in range(args.warmup):
        searcher.search_device(q_dev, k_fetch, args.min_score)
    launches0 = ix.stats()["kernel_launches"]
    barrier()
    t_region0 = time.monotonic()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_ms = []
    # timing rule: inputs larger than L2, or L2 flushed between timed iterations.  A shard that could stay
    # (partly) resident in the 126 MB L2 is timed step by step with a 256 MB memset between steps (outside
    # the events); the large workloads stream far more than L2 per step and are timed as one region.
    l2_flush = 2.0 * (hi - lo) * d < 4 * 126e6
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=device) if l2_flush else None
    if l2_flush:
        ms = 0.0
        for _ in range(args.steps):
            flush_buf.zero_()
            ev0.record(stream)
            searcher.search_device(q_dev, k_fetch, args.min_score)
            ev1.record(stream)
            torch.cuda.synchronize(device)
            ms += ev0.elapsed_time(ev1)
            scan_ms.append(ix.stats()["last_scan_ms"])
        barrier()
    else:
        ev0.record(stream)
        for _ in range(args.steps):
            searcher.search_device(q_dev, k_fetch, args.min_score)
            scan_ms.append(ix.stats()["last_scan_ms"])
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
    launches = ix.stats()["kernel_launches"] - launches0 + (args.steps if world > 1 else 0)
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        t_region1 = time.monotonic()
        time.sleep(0.05)   # let the last sample of the region arrive
        sampler.window(t_region0, t_region1)
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- e2e: host buffers through the public call ----------------
    def e2e_step():
        if world == 1:
            return ix.search(q_host.numpy(), k_fetch, args.min_score)      # rbk_index_search_f32: H2D + D2H inside
        return searcher.search(q_host, k_fetch, args.min_score, device)
    for _ in range(args.warmup):
        e2e_step()
    barrier()
    if l2_flush:
        e2e_s = 0.0
        for _ in range(args.steps):
            flush_buf.zero_()
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            e2e_step()                      # returns host results: the call has synchronised
            e2e_s += time.perf_counter() - t0
        barrier()
    else:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    fallback = ix.stats()["fallback_queries"]

    if rank == 0:
        pk = peaks()
        value = B * args.steps / (ms * 1e-3)
        n_local = hi - lo
        scan_avg_ms = float(np.mean(scan_ms))
        flops = 2.0 * B * n_local * d          # dot products only
        bytes_ = 2.0 * n_local * d + 2.0 * B * d + 8.0 * B * k_fetch
        tensor_bound = B > pk["tf"] * 1e12 / (pk["hbm"] * 1e9)   # arithmetic intensity ~ B flop/byte vs ridge
        # B200_PROFILING.md: burst peak for a kernel timed alone, sustained peak for a kernel timed inside a
        # long step.  The scan kernels here run back to back for the whole timed region; when that region is
        # long enough for the 1 kW power cap to engage (sw_power_cap seen) the sustained figure applies.
        capped = bool(clocks and "sw_power_cap" in clocks["reasons"]) and ms > 150.0 and pk["tf_sus"]
        if tensor_bound:
            ach = flops / (scan_avg_ms * 1e-3) / 1e12
            peak = pk["tf_sus"] if capped else pk["tf"]
            roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "peak_kind": "sustained (long power-capped loop)" if capped else "burst",
                    "frac_of_burst": ach / pk["tf"], "frac_of_sustained": ach / pk["tf_sus"] if pk["tf_sus"] else None,
                    "hbm_gbs": bytes_ / (scan_avg_ms * 1e-3) / 1e9}
        else:
            ach = bytes_ / (scan_avg_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                    "tflops": flops / (scan_avg_ms * 1e-3) / 1e12}
        roof.update({"kernel": ("scan2_kernel<streamed>: CTA-pair tcgen05 cta_group::2 QxC^T + fused top-k'" if B > 128
                                else "scan_kernel: tcgen05 QxC^T + fused top-k'"), "kernel_ms": scan_avg_ms,
                     "peak_source": pk["src"], "traffic": _profiled_traffic(args.workload)})
        out = {
            "metric": "knn_queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": desc + (" [REDUCED rows: debug run]" if reduced else ""), "n_docs": n, "dim": d,
                       "batch": B, "k": k, "k_fetch": k_fetch, "min_score": args.min_score,
                       "parallelism": f"rows sharded over {world} GPU(s), all-gather of top-k" if world > 1 else "1 GPU",
                       "rerank": "exact fp64 re-rank of k' candidates, ids/scores identical to the fp64 oracle",
                       "l2": (f"corpus {2.0 * n_local * d / 1e6:.0f} MB per GPU could stay in the 126 MB L2: L2 flushed "
                              "(256 MB memset) before every timed step, steps timed one by one" if l2_flush else
                              f"corpus {2.0 * n_local * d / 1e9:.1f} GB per GPU >> 126 MB L2, no flush needed"),
                       "fallback_queries": int(fallback)},
            "clocks": clocks,
            "e2e": {"value": B * args.steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": B * d * 4,
                    "d2h_bytes_per_step": B * k_fetch * 16 + B * 8, "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": int(launches),
            "roofline": roof,
        }
        if not args.no_cpu_baseline:
            import oracle
            oracle.build()
            cores = os.cpu_count() or 1
            ns = min(n_local, 200_000)
            bs = min(B, 64)
            sample_rows = ix.read_rows_bf16(0, ns)
            v, dt = cpu_reference_qps(sample_rows, q_host.numpy()[:bs], k_fetch, n, cores)
            out["cpu_baseline"] = {"value": v, "unit": "queries/s", "cores": cores, "kind": "port",
                                   "sample": f"first {ns} rows x {bs} queries on all {cores} host threads "
                                             f"({dt:.2f} s), scaled linearly to {n} rows"}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ix.close()


def _profiled_traffic(workload: str):
    """dram bytes per scan launch from the committed ncu --set full capture, if any."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        return json.loads(p.read_text()).get(workload)
    return None


if __name__ == "__main__":
    main()
