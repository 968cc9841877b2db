"""ctypes driver for librbk_oracle.so (oracle/rbk_oracle.c).  Test infrastructure.

Every function cites the reference lines its C counterpart restates; see rbk_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    """Compile the C restatement with the contract flags (oracle/Makefile)."""
    so = _HERE / "librbk_oracle.so"
    src = _HERE / "rbk_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "librbk_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = C.CDLL(str(build()))
        i64, f64, vp = C.c_int64, C.c_double, C.c_void_p
        lib.rbk_oracle_cosine.restype = f64
        lib.rbk_oracle_cosine.argtypes = [vp, vp, i64]
        lib.rbk_oracle_scores_f64.argtypes = [vp, i64, i64, vp, vp]
        lib.rbk_oracle_scores_bf16.argtypes = [vp, i64, i64, vp, vp]
        for name in ("rbk_oracle_search_f64", "rbk_oracle_search_bf16"):
            fn = getattr(lib, name)
            fn.restype = i64
            fn.argtypes = [vp, i64, i64, vp, i64, vp, C.c_int, f64, i64, vp, vp]
        lib.rbk_oracle_search_batch_bf16_mt.restype = i64
        lib.rbk_oracle_search_batch_bf16_mt.argtypes = [vp, i64, i64, vp, i64, vp, C.c_int, f64, i64, C.c_int,
                                                        vp, vp, vp]
        lib.rbk_oracle_search_batch_bf16_verify.restype = i64
        lib.rbk_oracle_search_batch_bf16_verify.argtypes = [vp, i64, i64, vp, i64, vp, C.c_int, f64, i64, C.c_int, i64,
                                                            vp, vp, vp]
        lib.rbk_oracle_rrf.restype = i64
        lib.rbk_oracle_rrf.argtypes = [vp, i64, vp, i64, f64, f64, f64, i64, vp, vp]
        _LIB = lib
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def cosine(a, b) -> float:
    """embedder.ts:168-184.  Raises like the reference on a length mismatch (S2)."""
    a, b = _f64(a), _f64(b)
    if a.shape[0] != b.shape[0]:
        raise ValueError("Vectors must have the same length")
    return float(_lib().rbk_oracle_cosine(_p(a), _p(b), a.shape[0]))


def _corpus(corpus):
    """float64 [n,d] or bf16-as-uint16 [n,d]."""
    if corpus.dtype == np.uint16:
        return np.ascontiguousarray(corpus), True
    return _f64(corpus), False


def scores(corpus, query) -> np.ndarray:
    c, bf = _corpus(corpus)
    q = _f64(query)
    out = np.empty(c.shape[0], dtype=np.float64)
    fn = _lib().rbk_oracle_scores_bf16 if bf else _lib().rbk_oracle_scores_f64
    fn(_p(c), c.shape[0], c.shape[1], _p(q), _p(out))
    return out


def search(corpus, query, k_fetch: int, min_score: float | None = 0.5, live=None):
    """vector-store.ts:207-221: scan, `>= minScore`, stable sort desc, first k_fetch.

    min_score=None -> findMostSimilar semantics (embedder.ts:189-202, no threshold).
    Returns (slots int64[cnt], scores float64[cnt]).
    """
    c, bf = _corpus(corpus)
    q = _f64(query)
    lv = None if live is None else np.ascontiguousarray(live, dtype=np.uint8)
    out_s = np.empty(max(k_fetch, 1), dtype=np.int64)
    out_v = np.empty(max(k_fetch, 1), dtype=np.float64)
    fn = _lib().rbk_oracle_search_bf16 if bf else _lib().rbk_oracle_search_f64
    cnt = fn(_p(c), c.shape[0], c.shape[1], _p(q), q.shape[0], _p(lv), 0 if min_score is None else 1,
             0.0 if min_score is None else float(min_score), k_fetch, _p(out_s), _p(out_v))
    if cnt == -2:
        raise ValueError("Vectors must have the same length")
    return out_s[:cnt].copy(), out_v[:cnt].copy()


def find_most_similar(query, embeddings, top_k: int = 10):
    """embedder.ts:189-202."""
    return search(embeddings, query, top_k, None)


def search_batch_mt(corpus_bf16, queries, k_fetch: int, min_score: float | None = 0.5, live=None,
                    n_threads: int | None = None):
    """ref-allcores: rows split across threads, identical per-pair arithmetic."""
    c = np.ascontiguousarray(corpus_bf16, dtype=np.uint16)
    q = _f64(queries)
    assert q.ndim == 2 and q.shape[1] == c.shape[1]
    lv = None if live is None else np.ascontiguousarray(live, dtype=np.uint8)
    nq = q.shape[0]
    out_s = np.full((nq, k_fetch), -1, dtype=np.int64)
    out_v = np.full((nq, k_fetch), np.nan, dtype=np.float64)
    out_c = np.zeros(nq, dtype=np.int32)
    nt = n_threads or (os.cpu_count() or 1)
    _lib().rbk_oracle_search_batch_bf16_mt(_p(c), c.shape[0], c.shape[1], _p(q), nq, _p(lv),
                                           0 if min_score is None else 1,
                                           0.0 if min_score is None else float(min_score), k_fetch, nt,
                                           _p(out_s), _p(out_v), _p(out_c))
    return out_s, out_v, out_c


def host_threads() -> int:
    """Host threads this process may really use (a cgroup / affinity mask can be narrower than cpu_count)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def search_batch_verify(corpus_bf16, queries, k_fetch: int, min_score: float | None = 0.5, live=None,
                        n_threads: int | None = None, slot_base: int = 0):
    """Checker for large runs: identical per-pair arithmetic (hence identical results) to search_batch_mt,
    norms hoisted and 8 queries' dot chains advanced side by side (rbk_oracle.c, "verify" variant)."""
    c = np.ascontiguousarray(corpus_bf16, dtype=np.uint16)
    q = _f64(queries)
    assert q.ndim == 2 and q.shape[1] == c.shape[1]
    lv = None if live is None else np.ascontiguousarray(live, dtype=np.uint8)
    nq = q.shape[0]
    out_s = np.full((nq, k_fetch), -1, dtype=np.int64)
    out_v = np.full((nq, k_fetch), np.nan, dtype=np.float64)
    out_c = np.zeros(nq, dtype=np.int32)
    _lib().rbk_oracle_search_batch_bf16_verify(_p(c), c.shape[0], c.shape[1], _p(q), nq, _p(lv),
                                               0 if min_score is None else 1,
                                               0.0 if min_score is None else float(min_score), k_fetch,
                                               n_threads or host_threads(), slot_base, _p(out_s), _p(out_v), _p(out_c))
    return out_s, out_v, out_c


def merge_lists(parts, k_fetch: int):
    """Merge per-chunk / per-shard (slots, scores, counts) results by (score desc, slot asc) and cut: the first
    k_fetch of the stable descending sort over the whole corpus (vector-store.ts:218-221), since slots are
    globally unique and each part already holds its own first k_fetch."""
    nq = parts[0][0].shape[0]
    out_s = np.full((nq, k_fetch), -1, dtype=np.int64)
    out_v = np.full((nq, k_fetch), np.nan, dtype=np.float64)
    out_c = np.zeros(nq, dtype=np.int32)
    for b in range(nq):
        ent = [(-float(v[b, i]), int(s[b, i])) for s, v, c in parts for i in range(int(c[b]))]
        ent.sort()
        ent = ent[:k_fetch]
        out_c[b] = len(ent)
        for i, (nv, sl) in enumerate(ent):
            out_s[b, i], out_v[b, i] = sl, -nv
    return out_s, out_v, out_c


def search_chunked(read_rows, n_rows: int, queries, k_fetch: int, min_score: float | None = 0.5,
                   chunk_rows: int = 1 << 20, n_threads: int | None = None, slot_base: int = 0, live=None):
    """Oracle answer for a corpus too large to hold on the host at once: read_rows(first, n) -> uint16 [n, d]
    (e.g. Index.read_rows_bf16) is called chunk by chunk, each chunk is answered by search_batch_verify, the
    per-chunk lists are merged.  Returned slots are slot_base + row."""
    parts = []
    for r0 in range(0, n_rows, chunk_rows):
        m = min(chunk_rows, n_rows - r0)
        lv = None if live is None else live[r0:r0 + m]
        parts.append(search_batch_verify(read_rows(r0, m), queries, k_fetch, min_score, live=lv,
                                         n_threads=n_threads, slot_base=slot_base + r0))
    if not parts:
        nq = np.asarray(queries).shape[0]
        return (np.full((nq, k_fetch), -1, dtype=np.int64), np.full((nq, k_fetch), np.nan),
                np.zeros(nq, dtype=np.int32))
    return merge_lists(parts, k_fetch)


def rrf(fts_ids, vec_ids, top_k: int, rrf_k: float = 60.0, fts_w: float = 0.4, vec_w: float = 0.6):
    """hybrid-search.ts:106-151 on interned integer ids."""
    f = np.ascontiguousarray(fts_ids, dtype=np.int64)
    v = np.ascontiguousarray(vec_ids, dtype=np.int64)
    cap = max(len(f) + len(v), 1)
    out_i = np.empty(cap, dtype=np.int64)
    out_s = np.empty(cap, dtype=np.float64)
    cnt = _lib().rbk_oracle_rrf(_p(f), len(f), _p(v), len(v), rrf_k, fts_w, vec_w, top_k, _p(out_i), _p(out_s))
    return out_i[:cnt].copy(), out_s[:cnt].copy()
