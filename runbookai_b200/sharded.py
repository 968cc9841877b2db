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
        """One PACKED block per rank (slots i64 | scores f64 | counts i32, include/rbk_knn.h) so the
        exchange is a single all-gather; the local search writes straight into this rank's block."""
        key = (B, k, str(device))
        if key not in self._bufs:
            G = self.world
            nk = B * k
            blk = _native.packed_block_bytes(B, k) if device.type == "cuda" else nk * 16 + -(-B * 4 // 16) * 16
            local = torch.empty((blk,), dtype=torch.uint8, device=device)
            allb = torch.empty((G * blk,), dtype=torch.uint8, device=device)

            def views(buf, off):
                return (buf[off:off + nk * 8].view(torch.int64).view(B, k),
                        buf[off + nk * 8:off + nk * 16].view(torch.float64).view(B, k),
                        buf[off + nk * 16:off + nk * 16 + B * 4].view(torch.int32))
            self._bufs = {key: dict(blk=blk, local=local, all=allb, l=views(local, 0),
                                    g=[views(allb, g * blk) for g in range(G)],
                                    os=torch.empty((B, k), dtype=torch.int64, device=device),
                                    ov=torch.empty((B, k), dtype=torch.float64, device=device),
                                    oc=torch.empty((B,), dtype=torch.int32, device=device))}
        return self._bufs[key]

    def search_device(self, q_dev: torch.Tensor, k_fetch: int, min_score: float | None):
        """q_dev: float32 [B, d] on this rank's GPU.  Returns device (slots, scores, counts)."""
        B = q_dev.shape[0]
        buf = self._buffers(B, k_fetch, q_dev.device)
        ls, lv, lc = buf["l"]
        if self._local_search is not None:      # CPU test hook
            s, v, c = self._local_search(q_dev, k_fetch, min_score)
            ls.copy_(s), lv.copy_(v), lc.copy_(c)
        else:
            self.index.search_device(q_dev.data_ptr(), B, k_fetch, min_score, ls.data_ptr(), lv.data_ptr(),
                                     lc.data_ptr())
        if self.world == 1:
            return ls, lv, lc
        # the single exchange step of the path: one all-gather of <= B*(k*16+4) bytes per rank
        dist.all_gather_into_tensor(buf["all"], buf["local"], group=self.group)
        if self._merge is not None:             # CPU test hook
            gs = torch.stack([g[0] for g in buf["g"]])
            gv = torch.stack([g[1] for g in buf["g"]])
            gc = torch.stack([g[2] for g in buf["g"]])
            s, v, c = self._merge(gs, gv, gc, k_fetch)
            buf["os"].copy_(s), buf["ov"].copy_(v), buf["oc"].copy_(c)
        else:
            stream = torch.cuda.current_stream(q_dev.device).cuda_stream
            _native.merge_topk_packed_device(q_dev.device.index or 0, stream, self.world, B, k_fetch,
                                             buf["all"].data_ptr(), buf["os"].data_ptr(), buf["ov"].data_ptr(),
                                             buf["oc"].data_ptr())
        return buf["os"], buf["ov"], buf["oc"]

    def search(self, queries_host: torch.Tensor, k_fetch: int, min_score: float | None, device):
        """Host float32 queries (pinned or not) -> host results; H2D and D2H inside."""
        q = queries_host.to(device, non_blocking=True)
        s, v, c = self.search_device(q, k_fetch, min_score)
        return s.cpu(), v.cpu(), c.cpu()
