"""Development aid: with a library built with -DRBK_EPI_PROFILE, one search prints the cycle breakdown of
one epilogue thread of three units (first, second, last) of query block 0."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from runbookai_b200 import Index, synth

b = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sizes = [int(x) for x in sys.argv[2:]] or [65536, 1048576]
d, k = 768, 32
q = synth.random_queries(b, d, 8).astype(np.float32)
g = torch.Generator(device="cuda").manual_seed(7)
ix = Index(d, capacity_hint=max(sizes))
have = 0
for n in sizes:
    while have < n:
        m = min(1 << 18, n - have)
        t = torch.randn(m, d, device="cuda", generator=g, dtype=torch.float32).to(torch.bfloat16)
        torch.cuda.synchronize()
        ix.append_bf16_device(t.data_ptr(), m)
        have += m
    for it in range(3):
        print(f"--- n={n} b={b} run {it}", flush=True)
        ix.search(q, k, None)
        torch.cuda.synchronize()
        print("scan_ms", ix.stats()["last_scan_ms"], flush=True)
ix.close()
