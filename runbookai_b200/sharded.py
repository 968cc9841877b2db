"""Row-sharded search across the GPUs of one box.

One process per GPU (torchrun): rank r owns the contiguous slot block
[r*ceil(N/G), (r+1)*ceil(N/G)) as its own `Index` with slot_base set, queries are
replicated, every rank answers them against its shard (exact fp64 scores, so per-shard
lists are already final within the shard), then ONE all-gather of the per-rank packed
block (slots | scores | counts | exactness flags) over NCCL/NVLink and a merge kernel give
every rank the global answer.  Contiguous blocks keep slot order, so the (score desc,
slot asc) tie-break survives the merge.

A step is ONE enqueue and ONE host synchronisation: the local search is queued without a
host round trip (`rbk_index_search_device_async`), the all-gather and the merge are queued
behind it on the same stream, and the "not proven exact" flags travel inside the packed
block, are OR-ed by the merge and are checked once, after the merge, together with the
results.  Every rank sees the same merged flags, so the (rare) decision to re-answer a batch
through the synchronous path is taken identically everywhere without another exchange.
torch is plumbing here: tensors for device buffers, the process group for the collective.
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
    """search(queries) over a row-sharded corpus; every rank returns the global result.

    All device work (H2D of the queries, the engine's kernels, the NCCL all-gather, the merge, D2H of the
    results) is issued on ONE stream, `self.stream`, which is also made the index's stream: stream order is the
    only synchronisation between the steps.  (With the index on a private stream a non-blocking H2D of the
    queries on torch's stream could still be in flight when the scan starts reading them.)
    """

    def __init__(self, index, group=None, local_search: Callable | None = None, merge: Callable | None = None,
                 stream: "torch.cuda.Stream | None" = None):
        self.index = index
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._local_search = local_search
        self._merge = merge
        self._bufs: dict = {}
        self._dirty_seen = 0
        self.redone_batches = 0
        self.stream = None
        if local_search is None:   # a real device index
            dev = torch.device("cuda", index.device)
            self.stream = stream if stream is not None else torch.cuda.Stream(dev)
            index.set_stream(self.stream.cuda_stream)

    # ------------------------------------------------------------------ buffers
    def _buffers(self, B: int, k: int, device):
        """One PACKED block per rank (slots i64 | scores f64 | counts i32 | flags i32, include/rbk_knn.h) so the
        exchange is a single all-gather; the local search writes straight into this rank's block."""
        key = (B, k, str(device))
        if key not in self._bufs:
            G = self.world
            nk = B * k
            cuda = device.type == "cuda"
            pad = -(-B * 4 // 16) * 16
            blk = _native.packed_block_bytes(B, k) if cuda else nk * 16 + 2 * pad
            off_f = _native.packed_flags_offset(B, k) if cuda else nk * 16 + pad
            local = torch.zeros((blk,), dtype=torch.uint8, device=device)
            allb = torch.zeros((G * blk,), dtype=torch.uint8, device=device)

            def views(buf, off):
                return (buf[off:off + nk * 8].view(torch.int64).view(B, k),
                        buf[off + nk * 8:off + nk * 16].view(torch.float64).view(B, k),
                        buf[off + nk * 16:off + nk * 16 + B * 4].view(torch.int32),
                        buf[off + off_f:off + off_f + B * 4].view(torch.int32))
            out = torch.zeros((nk * 16 + (2 * B + 1) * 4,), dtype=torch.uint8, device=device)   # os | ov | oc | of[B+1]
            b = dict(blk=blk, local=local, all=allb, l=views(local, 0), g=[views(allb, g * blk) for g in range(G)],
                     out=out, os=out[:nk * 8].view(torch.int64).view(B, k),
                     ov=out[nk * 8:nk * 16].view(torch.float64).view(B, k),
                     oc=out[nk * 16:nk * 16 + B * 4].view(torch.int32),
                     of=out[nk * 16 + B * 4:].view(torch.int32))
            if cuda:
                b["h_out"] = torch.zeros_like(out, device="cpu").pin_memory()
            self._bufs = {key: b}
            self._dirty_seen = 0          # fresh (zeroed) buffers: the merge kernel's running count restarts
        return self._bufs[key]

    # ------------------------------------------------------------------ device-resident path
    def search_device_async(self, q_dev: torch.Tensor, k_fetch: int, min_score: float | None):
        """Enqueue one sharded search on `self.stream`; no host synchronisation.  q_dev: float32 [B, d] on this
        rank's GPU (written on `self.stream`, or already complete).  Returns device (slots, scores, counts,
        flags[B+1]): flags[b] = 1 if query b is not proven exact on some shard, flags[B] = running count of such
        queries over all calls.  Callers that cannot tolerate an unproven answer use search_device()."""
        B = q_dev.shape[0]
        buf = self._buffers(B, k_fetch, q_dev.device)
        ls, lv, lc, lf = buf["l"]
        with torch.cuda.stream(self.stream):
            self.index.search_device_async(q_dev.data_ptr(), B, k_fetch, min_score, ls.data_ptr(), lv.data_ptr(),
                                           lc.data_ptr(), lf.data_ptr())
            self._exchange(buf, B, k_fetch)
        return buf["os"], buf["ov"], buf["oc"], buf["of"]

    def _exchange(self, buf, B: int, k_fetch: int):
        """The single exchange step of the path: one all-gather of <= B*(k*16+8) bytes per rank, then the merge
        (which also ORs the exactness flags).  With one rank the merge of one block just copies it to the
        output arrays and counts the dirty flags, so both shapes of the job share one code path."""
        src = buf["local"]
        if self.world > 1:
            dist.all_gather_into_tensor(buf["all"], buf["local"], group=self.group)
            src = buf["all"]
        _native.merge_topk_packed_device(self.index.device, self.stream.cuda_stream, self.world, B, k_fetch,
                                         src.data_ptr(), buf["os"].data_ptr(), buf["ov"].data_ptr(),
                                         buf["oc"].data_ptr(), buf["of"].data_ptr())

    def _finish(self, buf, q_dev, B: int, k_fetch: int, min_score):
        """After the one synchronisation of a step: did any shard fail to prove a query?  Then every rank (they all
        read the same merged flags) re-answers the batch through the synchronous call, which rescans with the
        widest margin and falls back to the exhaustive fp64 kernel, and the exchange is repeated."""
        dirty_total = int(buf["h_out"][-4:].view(torch.int32)[0])
        if dirty_total == self._dirty_seen:
            return False
        self._dirty_seen = dirty_total
        self.redone_batches += 1
        ls, lv, lc, lf = buf["l"]
        with torch.cuda.stream(self.stream):
            self.index.search_device(q_dev.data_ptr(), B, k_fetch, min_score, ls.data_ptr(), lv.data_ptr(),
                                     lc.data_ptr())
            lf.zero_()                      # exact by construction now
            self._exchange(buf, B, k_fetch)
            buf["h_out"].copy_(buf["out"], non_blocking=True)
        self.stream.synchronize()
        self._dirty_seen = int(buf["h_out"][-4:].view(torch.int32)[0])
        return True

    def search_device(self, q_dev: torch.Tensor, k_fetch: int, min_score: float | None):
        """q_dev: float32 [B, d] on this rank's GPU.  Returns device (slots, scores, counts), exact; the call has
        synchronised `self.stream` once."""
        B = q_dev.shape[0]
        if self._local_search is not None:
            return self._search_hooks(q_dev, k_fetch, min_score)
        s, v, c, f = self.search_device_async(q_dev, k_fetch, min_score)
        buf = self._buffers(B, k_fetch, q_dev.device)
        with torch.cuda.stream(self.stream):
            buf["h_out"][-(B + 1) * 4:].copy_(buf["out"][-(B + 1) * 4:], non_blocking=True)   # flags only
        self.stream.synchronize()           # the ONE host round trip of a step
        self._finish(buf, q_dev, B, k_fetch, min_score)
        return s, v, c

    def search(self, queries_host: torch.Tensor, k_fetch: int, min_score: float | None, device):
        """Host float32 queries (pinned or not) -> host results; H2D and D2H inside, one synchronisation."""
        if self._local_search is not None:
            s, v, c = self._search_hooks(queries_host.to(device), k_fetch, min_score)
            return s.cpu(), v.cpu(), c.cpu()
        B = queries_host.shape[0]
        with torch.cuda.stream(self.stream):
            q = queries_host.to(device, non_blocking=True)
        self.search_device_async(q, k_fetch, min_score)
        buf = self._buffers(B, k_fetch, q.device)
        with torch.cuda.stream(self.stream):
            buf["h_out"].copy_(buf["out"], non_blocking=True)   # results + flags in one D2H
        self.stream.synchronize()
        self._finish(buf, q, B, k_fetch, min_score)
        nk = B * k_fetch
        h = buf["h_out"]
        return (h[:nk * 8].view(torch.int64).view(B, k_fetch).clone(),
                h[nk * 8:nk * 16].view(torch.float64).view(B, k_fetch).clone(),
                h[nk * 16:nk * 16 + B * 4].view(torch.int32).clone())

    # ------------------------------------------------------------------ CPU test hooks (gloo, world_size 2)
    def _search_hooks(self, q_dev, k_fetch, min_score):
        B = q_dev.shape[0]
        buf = self._buffers(B, k_fetch, q_dev.device)
        ls, lv, lc, _ = buf["l"]
        s, v, c = self._local_search(q_dev, k_fetch, min_score)
        ls.copy_(s), lv.copy_(v), lc.copy_(c)
        if self.world == 1:
            return ls, lv, lc
        dist.all_gather_into_tensor(buf["all"], buf["local"], group=self.group)
        gs = torch.stack([g[0] for g in buf["g"]])
        gv = torch.stack([g[1] for g in buf["g"]])
        gc = torch.stack([g[2] for g in buf["g"]])
        merge = self._merge or merge_topk_host
        s, v, c = merge(gs, gv, gc, k_fetch)
        buf["os"].copy_(s), buf["ov"].copy_(v), buf["oc"].copy_(c)
        return buf["os"], buf["ov"], buf["oc"]
