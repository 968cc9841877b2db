"""torchrun parity check of the sharded path (NCCL all-gather + merge kernel) against the oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 \
        scripts/dist_check.py
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    from runbookai_b200 import Index, synth
    from runbookai_b200.sharded import ShardedSearcher, shard_bounds
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, d, b, k = 400_003, 768, 300, 32
    corpus = synth.random_corpus(n, d, 5)           # same on every rank
    per = -(-n // world)
    for g in range(1, world):                        # exact ties straddling every shard boundary
        corpus[g * per + 1] = corpus[g * per - 2]
    q = synth.random_queries(b, d, 6)
    synth.plant_neighbours(corpus, q, 8, 7)
    # 150 exact duplicates of one row, all in shard 0: more ties than any candidate margin holds -> that shard's
    # proof fails for query 0, the dirty flag travels through the all-gather and EVERY rank re-answers the batch
    dup = np.random.default_rng(8).choice(per - 10, 150, replace=False)
    corpus[dup] = synth.f32_to_bf16_bits(q[0] * 0.5)
    lo, hi = shard_bounds(n, world, rank)
    ix = Index(d, device=local, capacity_hint=hi - lo)
    ix.set_slot_base(lo)
    ix.append_bf16(corpus[lo:hi])
    sh = ShardedSearcher(ix)               # one stream for the engine, NCCL and the copies
    ok = True
    for ms in (None, 0.5):
        s, v, c = sh.search(torch.from_numpy(q), k, ms, dev)
        s, v, c = s.numpy(), v.numpy(), c.numpy()
        if rank == 0:
            import oracle
            nq = 64
            es, ev, ec = oracle.search_batch_verify(corpus, q[:nq].astype(np.float64), k, ms)
            good = (c[:nq] == ec).all() and all(
                (s[i, :ec[i]] == es[i, :ec[i]]).all() and (v[i, :ec[i]] == ev[i, :ec[i]]).all() for i in range(nq))
            ok = ok and bool(good)
            print(json.dumps({"world": world, "min_score": ms, "parity": bool(good), "counts": c[:4].tolist(),
                              "fallback": ix.stats()["fallback_queries"], "redone_batches": sh.redone_batches}),
                  flush=True)
            ok = ok and sh.redone_batches >= 1     # the dirty-flag path was really taken
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    ix.close()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
