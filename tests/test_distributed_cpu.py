"""CPU suite, part 3: the N>1 orchestration (row sharding, all-gather, merge order) with
world_size 2 over gloo.  The device index and the CUDA merge kernel are replaced by the
oracle-backed stand-ins; the sharding/collective/merge-order logic is the real one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, b, k, ret):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from common import OracleIndex
    from runbookai_b200 import synth
    from runbookai_b200.sharded import ShardedSearcher, merge_topk_host, shard_bounds
    corpus = synth.random_corpus(n, d, 21)
    corpus[n // 2 + 3] = corpus[5]          # a tie that straddles the shard boundary
    corpus[n // 2 + 9] = corpus[7]
    queries = synth.random_queries(b, d, 22)
    lo, hi = shard_bounds(n, world, rank)
    ix = OracleIndex(d)
    ix.set_slot_base(lo)
    ix.append_bf16(corpus[lo:hi])

    def local_search(q, k_fetch, min_score):
        s, v, c, _ = ix.search(q.numpy().astype(np.float64), k_fetch, min_score)
        return torch.from_numpy(s), torch.from_numpy(v), torch.from_numpy(c)

    sh = ShardedSearcher(ix, local_search=local_search, merge=merge_topk_host)
    s, v, c = sh.search_device(torch.from_numpy(queries), k, None)
    ret[rank] = (s.numpy().copy(), v.numpy().copy(), c.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sharded_search_equals_single_index(oracle_mod):
    from runbookai_b200 import synth
    n, d, b, k = 2001, 24, 6, 12
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), n, d, b, k, ret), nprocs=2, join=True)
    corpus = synth.random_corpus(n, d, 21)
    corpus[n // 2 + 3] = corpus[5]
    corpus[n // 2 + 9] = corpus[7]
    queries = synth.random_queries(b, d, 22).astype(np.float64)
    for rank in (0, 1):
        s, v, c = ret[rank]
        for i in range(b):
            es, ev = oracle_mod.search(corpus, queries[i], k, None)
            assert c[i] == len(es)
            assert s[i, :c[i]].tolist() == es.tolist()
            assert v[i, :c[i]].tolist() == ev.tolist()


def test_shard_bounds_cover_and_keep_order():
    from runbookai_b200.sharded import shard_bounds
    for n in (0, 1, 7, 8, 1000, 10_000_000):
        for w in (1, 2, 4, 8):
            prev = 0
            for r in range(w):
                lo, hi = shard_bounds(n, w, r)
                assert lo == prev and lo <= hi
                prev = hi
            assert prev == n
