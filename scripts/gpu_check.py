"""First-contact GPU validation: each stage runs in its own process (a trapped kernel
poisons the CUDA context) and appends one JSON line to gpurun_out/check.jsonl.

    python scripts/gpu_check.py            # all stages
    python scripts/gpu_check.py gemm       # one stage in-process
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)


def emit(stage, **kw):
    rec = {"stage": stage, **kw}
    with open(OUT / "check.jsonl", "a") as f:
        f.write(json.dumps(rec) + "\n")
    print(json.dumps(rec), flush=True)


def stage_gemm(n=1000, d=384, b=5, tag="gemm"):
    import numpy as np
    from runbookai_b200 import Index, synth
    c = synth.random_corpus(n, d, 1)
    q = synth.random_queries(b, d, 2)
    ix = Index(d)
    ix.append_bf16(c)
    got = ix.debug_scores(q)
    cf = synth.bf16_bits_to_f32(c).astype(np.float64)
    qf = q.astype(np.float64)
    ref = (qf @ cf.T) / (np.linalg.norm(qf, axis=1)[:, None] * np.linalg.norm(cf, axis=1)[None, :])
    err = np.abs(got - ref)
    bad = np.argwhere(~(err < 1e-3))
    emit(tag, n=n, d=d, b=b, max_err=float(np.nanmax(err)), n_bad=int(len(bad)), n_nan=int(np.isnan(got).sum()),
         first_bad=bad[:8].tolist(), sample_got=got[0, :6].tolist(), sample_ref=ref[0, :6].tolist(),
         bad_rows_mod8=sorted(set(int(x[1]) % 8 for x in bad[:2000])),
         bad_cols_hist=[int(((bad[:, 1] // 32) % 8 == i).sum()) for i in range(8)] if len(bad) else [])
    ix.close()


def stage_search(n=10000, d=384, b=1, k=10, planted=10, min_score=0.5, tag="search"):
    import numpy as np
    import oracle
    from runbookai_b200 import Index, synth
    c = synth.random_corpus(n, d, 3)
    q = synth.random_queries(b, d, 4)
    if planted:
        synth.plant_neighbours(c, q, planted, 5)
    ix = Index(d)
    ix.append_bf16(c)
    t0 = time.time()
    slots, scores, counts, ms = ix.search(q.astype(np.float64), k, min_score)
    wall = time.time() - t0
    st = ix.stats()
    nq = min(b, 64)
    os_, ov, oc = oracle.search_batch_mt(c, q[:nq].astype(np.float64), k, min_score)
    id_ok = sc_ok = 0
    for i in range(nq):
        same = counts[i] == oc[i] and (slots[i, :oc[i]] == os_[i, :oc[i]]).all()
        id_ok += bool(same)
        sc_ok += bool(same and (scores[i, :oc[i]] == ov[i, :oc[i]]).all())
    emit(tag, n=n, d=d, b=b, k=k, min_score=min_score, checked=nq, id_parity=id_ok, score_bitexact=sc_ok,
         counts=counts[:4].tolist(), ocounts=oc[:4].tolist(), slots0=slots[0, :6].tolist(),
         oslots0=os_[0, :6].tolist(), scores0=scores[0, :4].tolist(), oscores0=ov[0, :4].tolist(), device_ms=ms,
         wall_ms=wall * 1e3, scan_ms=st["last_scan_ms"], fallback=st["fallback_queries"], kprime=st["last_kprime"])
    ix.close()


def stage_time(n=1_000_000, d=768, b=256, k=32, iters=5, tag="time"):
    import numpy as np
    import torch
    from runbookai_b200 import Index, synth
    ix = Index(d, capacity_hint=n)
    g = torch.Generator(device="cuda").manual_seed(7)
    chunk = 1 << 18
    for r0 in range(0, n, chunk):
        m = min(chunk, n - r0)
        t = torch.randn(m, d, device="cuda", generator=g, dtype=torch.float32).to(torch.bfloat16)
        torch.cuda.synchronize()
        ix.append_bf16_device(t.data_ptr(), m)
    q = synth.random_queries(b, d, 8).astype(np.float32)
    res = []
    for it in range(iters):
        slots, scores, counts, ms = ix.search(q, k, None)
        st = ix.stats()
        res.append((ms, st["last_scan_ms"]))
    best_scan = min(r[1] for r in res[1:])
    emit(tag, n=n, d=d, b=b, k=k, total_ms=[round(r[0], 3) for r in res], scan_ms=[round(r[1], 3) for r in res],
         qps_scan=b / (best_scan * 1e-3), tflops=2.0 * b * n * d / (best_scan * 1e-3) / 1e12,
         gbps=2.0 * n * d / (best_scan * 1e-3) / 1e9, fallback=ix.stats()["fallback_queries"],
         ring=ix.stats()["last_ring_stages"],
         counts=counts[:4].tolist())
    ix.close()


def stage_sweep(b=1024):
    """scan time vs N at fixed B (fixed per-launch cost of the scan kernel)."""
    import numpy as np
    import torch
    from runbookai_b200 import Index, synth
    d, k = 768, 32
    q = synth.random_queries(b, d, 8).astype(np.float32)
    g = torch.Generator(device="cuda").manual_seed(7)
    ix = Index(d, capacity_hint=4_000_000)
    have = 0
    for n in (65536, 262144, 524288, 1048576, 2097152, 4194304):
        while have < n:
            m = min(1 << 18, n - have)
            t = torch.randn(m, d, device="cuda", generator=g, dtype=torch.float32).to(torch.bfloat16)
            torch.cuda.synchronize()
            ix.append_bf16_device(t.data_ptr(), m)
            have += m
        res = []
        for it in range(12):
            ix.search(q, k, None)
            res.append(ix.stats()["last_scan_ms"])
        emit("sweep", n=n, b=b, scan_ms_min=min(res[2:]), scan_ms_med=float(np.median(res[2:])),
             total_ms=ix.stats()["last_total_ms"])
    ix.close()


STAGES = {
    "gemm": lambda: stage_gemm(),
    "gemm2": lambda: stage_gemm(n=3000, d=768, b=130, tag="gemm2"),
    "gemm3": lambda: stage_gemm(n=700, d=100, b=3, tag="gemm3"),
    "search1": lambda: stage_search(),
    "search2": lambda: stage_search(n=50000, d=768, b=200, k=32, planted=40, tag="search2"),
    "search3": lambda: stage_search(n=200000, d=768, b=64, k=32, planted=0, min_score=None, tag="search3"),
    "search4": lambda: stage_search(n=300000, d=1024, b=300, k=64, planted=0, min_score=None, tag="search4"),
    "time1": lambda: stage_time(),
    "time2": lambda: stage_time(n=2_000_000, b=1024, tag="time2"),
    "time3": lambda: stage_time(n=1_000_000, b=1, k=10, tag="time3"),
    "sweep": lambda: stage_sweep(),
    "sweep256": lambda: stage_sweep(256),
    "sweep1": lambda: stage_sweep(1),
}

if __name__ == "__main__":
    if len(sys.argv) > 1:
        for s in sys.argv[1:]:
            STAGES[s]()
        sys.exit(0)
    for name in STAGES:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, name], timeout=300, capture_output=True, text=True)
            tail = (r.stdout + r.stderr)[-1500:]
            if r.returncode != 0:
                emit(name + "_FAILED", rc=r.returncode, tail=tail)
            else:
                print(r.stdout, end="")
        except subprocess.TimeoutExpired:
            emit(name + "_TIMEOUT", secs=time.time() - t0)
