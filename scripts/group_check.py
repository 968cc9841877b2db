"""rbk_group on every GPU of the box, ONE host process: parity against the oracle and end-to-end throughput of
`rbk_group_search_f32` (host queries in, host results out; per-GPU scans + one ncclAllGather + merge inside the call).

    python scripts/group_check.py [rows] [dim] [B] [k_fetch]
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    import oracle
    from runbookai_b200 import Group, synth
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    k = int(sys.argv[4]) if len(sys.argv) > 4 else 32
    G = torch.cuda.device_count()
    corpus = synth.random_corpus(n, d, 51)
    for g in range(1, G):
        corpus[g * 4096 + 1] = corpus[3]                  # exact ties spread over the devices
    q = synth.random_queries(B, d, 52)
    t0 = time.perf_counter()
    grp = Group(d, list(range(G)), capacity_hint=n)
    t_create = time.perf_counter() - t0
    t0 = time.perf_counter()
    for r0 in range(0, n, 1 << 18):
        grp.append_bf16(corpus[r0:r0 + (1 << 18)])
    t_load = time.perf_counter() - t0
    for _ in range(3):
        s, v, c, ms = grp.search(q, k, None)
    steps = 20
    t0 = time.perf_counter()
    for _ in range(steps):
        s, v, c, ms = grp.search(q, k, None)
    dt = (time.perf_counter() - t0) / steps
    nq = 64
    es, ev, ec = oracle.search_batch_verify(corpus, q[:nq].astype(np.float64), k, None)
    ok = bool((c[:nq] == ec).all() and (s[:nq] == es).all() and np.array_equal(v[:nq], ev, equal_nan=True))
    s5, v5, c5, _ = grp.search(q[:nq], k, 0.5)
    es5, ev5, ec5 = oracle.search_batch_verify(corpus, q[:nq].astype(np.float64), k, 0.5)
    ok5 = bool((c5 == ec5).all() and (s5 == es5).all() and np.array_equal(v5, ev5, equal_nan=True))
    st = grp.stats()
    print(json.dumps({"devices": G, "rows": n, "dim": d, "batch": B, "k_fetch": k, "parity": ok, "parity_min_score_0.5": ok5,
                      "e2e_ms_per_search": dt * 1e3, "e2e_qps": B / dt, "device_ms_last": ms,
                      "create_s": round(t_create, 2), "load_s": round(t_load, 2), "redone_batches": st["redone_batches"],
                      "fallback_queries": st["fallback_queries"]}), flush=True)
    grp.close()
    sys.exit(0 if ok and ok5 else 1)


if __name__ == "__main__":
    main()
