"""Row-sharded search across the GPUs of one box.

One process per GPU (torchrun): rank r owns the contiguous slot block
[r*ceil(N/G), (r+1)*ceil(N/G)) as its own `Index` with slot_base set, queries are
replicated, every rank answers them against its shard (exact fp64 scores, so per-shard
lists are already final within the shard), then ONE all-gather of the per-rank
(slots, scores, counts) over NCCL/NVLink and a merge kernel give every rank the global
answer.  Contiguous blocks keep slot order, so the (score desc, slot asc) tie-break
survives the merge.  torch is plumbing here: tensors for device buffers, the process
group for the collective.
"""
from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist

from . import _native


def shard_bounds(n_total: int, world: int, rank: int) -> tuple[int, int]:
    per = -(-n_total // world)
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def merge_topk_host(slots, scores, counts, k_fetch: int):
    """Reference-order merge on host tensors (CPU/gloo path of the tests): [G,B,k] -> [B,k]."""
    G, B, _ = slots.shape
    out_s = torch.full((B, k_fetch), -1, dtype=torch.int64)
    out_v = torch.full((B, k_fetch), float("nan"), dtype=torch.float64)
    out_c = torch.zeros((B,), dtype=torch.int32)
    for b in range(B):
        ent = [(-float(scores[g, b, i]), int(slots[g, b, i])) for g in range(G) for i in range(int(counts[g, b]))]
        ent.sort()
        ent = ent[:k_fetch]
        for i, (ns, sl) in enumerate(ent):
            out_s[b, i], out_v[b, i] = sl, -ns
        out_c[b] = len(ent)
    return out_s, out_v, out_c


class ShardedSearcher:
    """search(queries) over a row-sharded corpus; every rank returns the global result."""

    def __init__(self, index, group=None, local_search: Callable | None = None, merge: Callable | None = None):
        self.index = index
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._local_search = local_search
        self._merge = merge
        self._bufs: dict = {}

    def _buffers(self, B: int, k: int, device):
        key = (B, k, str(device))
        if key not in self._bufs:
            G = self.world
            self._bufs = {key: dict(
                ls=torch.empty((B, k), dtype=torch.int64, device=device),
                lv=torch.empty((B, k), dtype=torch.float64, device=device),
                lc=torch.empty((B,), dtype=torch.int32, device=device),
                gs=torch.empty((G, B, k), dtype=torch.int64, device=device),
                gv=torch.empty((G, B, k), dtype=torch.float64, device=device),
                gc=torch.empty((G, B), dtype=torch.int32, device=device),
                os=torch.empty((B, k), dtype=torch.int64, device=device),
                ov=torch.empty((B, k), dtype=torch.float64, device=device),
                oc=torch.empty((B,), dtype=torch.int32, device=device))}
        return self._bufs[key]

    def search_device(self, q_dev: torch.Tensor, k_fetch: int, min_score: float | None):
        """q_dev: float32 [B, d] on this rank's GPU.  Returns device (slots, scores, counts)."""
        B = q_dev.shape[0]
        buf = self._buffers(B, k_fetch, q_dev.device)
        if self._local_search is not None:      # CPU test hook
            ls, lv, lc = self._local_search(q_dev, k_fetch, min_score)
            buf["ls"].copy_(ls), buf["lv"].copy_(lv), buf["lc"].copy_(lc)
        else:
            self.index.search_device(q_dev.data_ptr(), B, k_fetch, min_score, buf["ls"].data_ptr(),
                                     buf["lv"].data_ptr(), buf["lc"].data_ptr())
        if self.world == 1:
            return buf["ls"], buf["lv"], buf["lc"]
        # the single exchange step of the path: per-rank top-k lists, <= B*k*20 bytes per rank
        G = self.world   # outputs viewed as the concatenation along dim 0 (what gloo insists on)
        dist.all_gather_into_tensor(buf["gs"].view(G * B, k_fetch), buf["ls"], group=self.group)
        dist.all_gather_into_tensor(buf["gv"].view(G * B, k_fetch), buf["lv"], group=self.group)
        dist.all_gather_into_tensor(buf["gc"].view(G * B), buf["lc"], group=self.group)
        if self._merge is not None:
            s, v, c = self._merge(buf["gs"], buf["gv"], buf["gc"], k_fetch)
            buf["os"].copy_(s), buf["ov"].copy_(v), buf["oc"].copy_(c)
        else:
            stream = torch.cuda.current_stream(q_dev.device).cuda_stream
            _native.merge_topk_device(q_dev.device.index or 0, stream, self.world, B, k_fetch,
                                      buf["gs"].data_ptr(), buf["gv"].data_ptr(), buf["gc"].data_ptr(),
                                      buf["os"].data_ptr(), buf["ov"].data_ptr(), buf["oc"].data_ptr())
        return buf["os"], buf["ov"], buf["oc"]

    def search(self, queries_host: torch.Tensor, k_fetch: int, min_score: float | None, device):
        """Host float32 queries (pinned or not) -> host results; H2D and D2H inside."""
        q = queries_host.to(device, non_blocking=True)
        s, v, c = self.search_device(q, k_fetch, min_score)
        return s.cpu(), v.cpu(), c.cpu()
