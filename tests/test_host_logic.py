"""CPU suite, part 2: host logic and the C-ABI surface (no compute calls without a GPU)."""
import os
import ctypes
import json
import re
import sqlite3
from pathlib import Path

import numpy as np
import pytest

from common import HashEmbedder, OracleIndex
from conftest import HAS_CUDA, ROOT


def test_library_exports_every_declared_symbol(native):
    header = (ROOT / "include" / "rbk_knn.h").read_text()
    declared = set(re.findall(r"\b(rbk_[a-z0-9_]+)\s*\(", header))
    declared -= {"rbk_index", "rbk_status", "rbk_stats"}
    assert declared == set(native.SYMBOLS), declared ^ set(native.SYMBOLS)
    lib = ctypes.CDLL(str(native.LIB_PATH))
    for sym in sorted(declared):
        assert hasattr(lib, sym), sym
    assert lib.rbk_abi_version() == 2


@pytest.mark.skipif(HAS_CUDA, reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback(native):
    from runbookai_b200 import Index, RbkError
    with pytest.raises(RbkError) as e:
        Index(8)
    assert e.value.status == native.RBK_ECUDA
    assert "no CPU path" in str(e.value)


def test_synth_bf16_rounding_matches_torch():
    import torch
    from runbookai_b200 import synth
    x = np.random.default_rng(0).standard_normal(10000).astype(np.float32) * 3
    x[:4] = [0.0, -0.0, 1.0039062, 65504.0]
    ours = synth.f32_to_bf16_bits(x)
    theirs = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert (ours == theirs).all()
    assert (synth.bf16_round(x) == torch.from_numpy(x).to(torch.bfloat16).float().numpy()).all()


def test_blob_codec_roundtrip_in_reference_schema(tmp_path):
    from runbookai_b200 import vector_store as vs
    emb = np.random.default_rng(1).standard_normal(48)
    buf = vs.float_array_to_buffer(emb)
    assert len(buf) == 48 * 8 and (vs.buffer_to_float_array(buf) == emb).all()
    db = sqlite3.connect(tmp_path / "v.db")
    db.executescript(vs.SCHEMA)
    db.execute("INSERT INTO vector_embeddings (id, chunk_id, document_id, embedding, content, type) "
               "VALUES ('vec_a','a','d',?, 'c','runbook')", (buf,))
    got = db.execute("SELECT embedding FROM vector_embeddings").fetchone()[0]
    assert (np.frombuffer(got, dtype="<f8") == emb).all()


def test_rrf_host_equals_oracle(oracle_mod):
    from runbookai_b200.hybrid_search import reciprocal_rank_fusion
    from runbookai_b200.vector_store import RetrievedChunk
    rng = np.random.default_rng(3)
    mk = lambda i: RetrievedChunk(id=f"c{i}", documentId="d", title="t", content="x", type="runbook")  # noqa: E731
    for _ in range(20):
        f = rng.permutation(30)[: rng.integers(0, 20)].tolist()
        v = rng.permutation(30)[: rng.integers(0, 20)].tolist()
        got = reciprocal_rank_fusion([mk(i) for i in f], [mk(i) for i in v], 10)
        ids, sc = oracle_mod.rrf(f, v, 10)
        assert [g.id for g in got] == [f"c{i}" for i in ids]
        assert [g.score for g in got] == sc.tolist()


@pytest.fixture
def store(tmp_path):
    from runbookai_b200 import embedder
    from runbookai_b200.vector_store import VectorStore
    embedder.configure(HashEmbedder(64))
    s = VectorStore(str(tmp_path / "vectors.db"), index_factory=lambda dim, dev: OracleIndex(dim))
    yield s
    s.close()
    embedder.reset()


def _chunks(n, doc="doc1", typ="runbook", services=("api",), text="redis connection pool exhausted restart"):
    return [{"chunk": {"id": f"{doc}_{i}", "documentId": doc, "content": f"{text} step {i}",
                       "sectionTitle": f"Section {i}"},
             "documentTitle": f"Title {doc}", "type": typ, "services": list(services)} for i in range(n)]


def test_vector_store_search_quirks(store, oracle_mod):
    from oracle import pyref
    from runbookai_b200 import embedder
    store.add_chunks(_chunks(12, "doc1", "runbook", ("api",)))
    store.add_chunks(_chunks(12, "doc2", "postmortem", ("db",), text="redis connection pool exhausted failover"))
    assert store.get_count() == 24 and store.has_document("doc1") and not store.has_document("nope")
    q = "redis connection pool exhausted"
    res = store.search(q, {"topK": 5, "minScore": 0.3})
    # independent restatement of :201-221 on the same (bf16-rounded) vectors
    rows = store.db.execute("SELECT id, embedding FROM vector_embeddings").fetchall()
    ref = pyref.vector_scan(embedder.embed_text(q), [(r["id"], np.frombuffer(r["embedding"], "<f8").tolist())
                                                    for r in rows], top_k=5, min_score=0.3)
    assert [f"vec_{r.id}" for r in res] == [i for i, _ in ref][:5]
    assert [r.score for r in res] == [s for _, s in ref][:5]
    assert all(set(r.to_dict()) == {"id", "documentId", "title", "content", "type", "services", "score"} for r in res)
    # S4: falsy minScore/topK fall back to 0.5 / 10
    assert store.search(q, {"minScore": 0, "topK": 0}) == store.search(q, {})
    # S8: filters apply AFTER the 2*topK cut -> fewer than topK even though more matches exist
    only_pm = store.search(q, {"topK": 3, "minScore": 0.3, "typeFilter": ["postmortem"]})
    cut = {i for i, _ in pyref.vector_scan(embedder.embed_text(q), [(r["id"], np.frombuffer(r["embedding"], "<f8")
                                           .tolist()) for r in rows], top_k=3, min_score=0.3)}
    assert {f"vec_{r.id}" for r in only_pm} == {i for i in cut if "doc2" in i}
    assert store.search(q, {"topK": 3, "minScore": 0.3, "serviceFilter": ["nope"]}) == []


def test_vector_store_mutation_semantics(store):
    store.add_chunks(_chunks(4, "docA"))
    before = store.search("redis connection pool exhausted", {"minScore": 0.2})
    assert len(before) == 4
    store.delete_document("docA")
    assert store.get_count() == 0 and store.search("redis connection pool exhausted", {"minScore": 0.2}) == []
    store.add_chunks(_chunks(2, "docA"))              # re-added ids are appended (new slots)
    assert store._index.size() == 6 and store._index.count() == 2
    store.add_chunk(_chunks(1, "docA")[0]["chunk"], "Title docA", "runbook", ["api"])   # re-set keeps its slot
    assert store._index.size() == 6
    store.clear()
    assert store.get_count() == 0 and store._index.size() == 0


def test_re_embedded_document_is_two_engine_calls_and_equals_sequential_sets(store, tmp_path):
    """addChunks over existing ids (vector-store.ts:135-183): the mirror issues ONE overwrite_f64_batch for the
    re-sets and ONE append for the new ids, and ends in the same Map as set() item by item - also when an id occurs
    twice in the batch (last value wins, position of the first occurrence)."""
    from runbookai_b200 import embedder
    from runbookai_b200.vector_store import VectorStore
    store.add_chunks(_chunks(6, "docA"))
    ix = store._index
    slots_before = dict(store._slot_of)
    ix.calls = 0
    again = _chunks(6, "docA", text="kafka consumer lag rebalance") + _chunks(3, "docB")
    again.append(_chunks(1, "docB", text="kafka broker down")[0])        # docB_0 a second time, other text
    store.add_chunks(again)
    assert ix.calls == 1                                                # one batch call for the six re-sets
    assert {v: s for v, s in store._slot_of.items() if v in slots_before} == slots_before
    assert [store._slot_of[f"vec_docB_{i}"] for i in range(3)] == [6, 7, 8] and ix.size() == 9
    # the same sequence, one set at a time, in a second store
    (tmp_path / "seq").mkdir()
    other = VectorStore(str(tmp_path / "seq" / "vectors.db"), index_factory=lambda dim, dev: OracleIndex(dim))
    try:
        other.add_chunks(_chunks(6, "docA"))
        for c in again:
            other.add_chunk(c["chunk"], c["documentTitle"], c["type"], c["services"])
        assert other._slot_of == store._slot_of and other._ids == store._ids
        assert np.array_equal(other._index.rows, store._index.rows)
        q = "kafka broker down"
        assert other.search(q, {"minScore": 0.1}) == store.search(q, {"minScore": 0.1})
    finally:
        other.close()


def test_vector_store_reload_is_rowid_order_and_errors(tmp_path):
    from runbookai_b200 import embedder
    from runbookai_b200.vector_store import VectorStore
    embedder.configure(HashEmbedder(32))
    p = str(tmp_path / "v.db")
    s = VectorStore(p, index_factory=lambda d, dev: OracleIndex(d))
    s.add_chunks(_chunks(5, "d1"))
    s.add_chunk(_chunks(1, "d1")[0]["chunk"], "Title d1", "runbook", ["api"])   # REPLACE moves the row to the end
    s.close()
    s2 = VectorStore(p, index_factory=lambda d, dev: OracleIndex(d))
    assert s2._ids == [f"vec_d1_{i}" for i in (1, 2, 3, 4, 0)]                  # S9b after restart
    embedder.configure(HashEmbedder(16))
    with pytest.raises(ValueError, match="Vectors must have the same length"):
        s2.search("redis")                                                       # S2
    embedder.reset()
    with pytest.raises(RuntimeError, match="Embedder not configured"):
        s2.search("redis")
    s2.close()


def test_hybrid_retriever_modes_with_cpu_standin(tmp_path, oracle_mod):
    """hybrid-search.ts:54-151 end to end on the host: FTS5 leg (SQLite) + vector leg (stand-in index) + RRF."""
    from runbookai_b200 import embedder
    from runbookai_b200.fts_store import KnowledgeStore
    from runbookai_b200.hybrid_search import HybridRetriever, reciprocal_rank_fusion
    from runbookai_b200.vector_store import VectorStore
    embedder.configure(HashEmbedder(64))
    fts = KnowledgeStore(str(tmp_path / "knowledge.db"))
    docs = {"d1": ("runbook", "redis connection pool exhausted restart the pool"),
            "d2": ("postmortem", "postgres replication lag after failover"),
            "d3": ("runbook", "kubernetes pod crashloop out of memory")}
    for did, (typ, text) in docs.items():
        fts.upsert_document({"id": did, "type": typ, "title": did.upper(), "services": ["api"],
                             "chunks": [{"id": f"{did}_{i}", "content": f"{text} step {i}", "sectionTitle": f"S{i}"}
                                        for i in range(3)]})
    h = HybridRetriever({"storePath": str(tmp_path / "knowledge.db"), "vectorStorePath": str(tmp_path / "vectors.db")},
                        fts_store=fts)
    h.vector_store.close()
    h.vector_store = VectorStore(str(tmp_path / "vectors.db"), index_factory=lambda d, dev: OracleIndex(d))
    for did, (typ, text) in docs.items():
        h.vector_store.add_chunks([{"chunk": {"id": f"{did}_{i}", "documentId": did, "content": f"{text} step {i}",
                                              "sectionTitle": f"S{i}"}, "documentTitle": did.upper(), "type": typ,
                                    "services": ["api"]} for i in range(3)])
    q = "redis connection pool"
    f = h.search(q, {"mode": "fts", "topK": 4})
    assert f and all(r.documentId == "d1" for r in f) and f[0].sourceUrl is None
    v = h.search(q, {"mode": "vector", "topK": 4})
    assert v and v[0].documentId == "d1"
    hy = h.search(q, {"topK": 4})
    want = reciprocal_rank_fusion(fts.search(q, {"limit": 8}), h.vector_store.search(q, {"topK": 8}), 4)
    assert [(r.id, r.score) for r in hy] == [(r.id, r.score) for r in want]
    assert h.search("a b", {"mode": "fts"}) == []            # terms of length <= 2 are dropped (sqlite.ts:139-147)
    by_type = h.search_by_type(q, {"topK": 6})
    assert set(by_type) == {"runbooks", "postmortems", "architecture", "knownIssues"}
    h.close()
    embedder.reset()


def test_micro_batcher_coalesces_and_matches_individual_searches(store):
    """concurrent search() calls share one device pass and each caller still gets exactly
    what its own VectorStore.search() returns (own topK, minScore, filters)."""
    import threading
    from runbookai_b200.batcher import MicroBatcher
    store.add_chunks(_chunks(15, "doc1", "runbook", ("api",)))
    store.add_chunks(_chunks(15, "doc2", "postmortem", ("db",), text="redis connection pool exhausted failover"))
    store.add_chunks(_chunks(15, "doc3", "runbook", ("web",), text="kubernetes pod crashloop oom"))
    asks = [("redis connection pool exhausted", {"topK": 5, "minScore": 0.3}),
            ("pod crashloop oom", {"topK": 3, "minScore": 0.2, "typeFilter": ["runbook"]}),
            ("redis failover", {"topK": 8, "minScore": 0.25, "serviceFilter": ["db"]}),
            ("connection pool", {}),
            ("nothing matches this zzz", {"topK": 4})] * 6
    want = [store.search(q, o) for q, o in asks]
    mb = MicroBatcher(store, window_ms=50.0, max_batch=64)
    got = [None] * len(asks)

    def worker(i):
        got[i] = mb.search(*asks[i])
    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(asks))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    mb.close()
    assert got == want
    assert mb.served == len(asks) and mb.batches <= 3        # 30 concurrent calls -> a handful of passes


def test_shared_index_is_refcounted_per_db_path(tmp_path):
    """call sites that build and close a store per use (hook-handlers.ts:329-337) attach to ONE
    index per db path; the last close frees it; mutations through any instance are seen by all."""
    from runbookai_b200 import embedder
    from runbookai_b200.vector_store import VectorStore, create_vector_store, shared_index_count
    embedder.configure(HashEmbedder(64))
    made = []

    def factory(d, dev):
        made.append(OracleIndex(d))
        return made[-1]
    base = str(tmp_path)
    a = create_vector_store(base, index_factory=factory)
    a.add_chunks(_chunks(6, "doc1", "runbook", ("api",)))
    b = create_vector_store(base, index_factory=factory)           # attaches: no second index, no reload
    assert len(made) == 1 and shared_index_count() == 1 and b._index is a._index
    b.add_chunks(_chunks(4, "doc2", "postmortem", ("db",), text="postgres replication lag"))
    q = "postgres replication lag"
    assert [r.id for r in a.search(q, {"minScore": 0.2})] == [r.id for r in b.search(q, {"minScore": 0.2})] != []
    a.delete_document("doc2")                                       # seen by b at once
    assert all(r.documentId != "doc2" for r in b.search(q, {"minScore": 0.05}))
    a.close()
    a.close()                                                       # idempotent, does not drop b's reference
    assert shared_index_count() == 1 and b.search("redis connection pool", {"minScore": 0.2})
    b.close()
    assert shared_index_count() == 0
    c = create_vector_store(base, index_factory=factory)           # fresh load from SQLite, rowid order
    assert len(made) == 2 and c.get_count() == 6 and len(c._ids) == 6
    private = VectorStore(str(tmp_path / "vectors.db"), index_factory=factory)   # constructor default: private
    assert len(made) == 3 and private._index is not c._index and shared_index_count() == 1
    private.close()
    c.close()
    embedder.reset()


def test_knowledge_retriever_sync_mirrors_documents_into_the_vector_store(tmp_path):
    """retriever/index.ts:44-126 with the sync hook of SURVEY §8f-2: documents land in the FTS store AND in the
    vector store; an updated document replaces its old vectors; search buckets chunks by type."""
    from runbookai_b200 import embedder
    from runbookai_b200.retriever import KnowledgeRetriever
    from runbookai_b200.vector_store import VectorStore
    embedder.configure(HashEmbedder(64))

    def doc(did, typ, text, n=3, services=("api",)):
        return {"id": did, "type": typ, "title": did.upper(), "services": list(services),
                "chunks": [{"id": f"{did}_{i}", "content": f"{text} step {i}", "sectionTitle": f"S{i}"} for i in range(n)]}
    batch = [[doc("d1", "runbook", "redis connection pool exhausted restart the pool"),
              doc("d2", "postmortem", "postgres replication lag after failover", services=("db",)),
              doc("d3", "known_issue", "kubernetes pod crashloop out of memory")]]
    vs = VectorStore(str(tmp_path / "vectors.db"), index_factory=lambda d, dev: OracleIndex(d))
    r = KnowledgeRetriever({"storePath": str(tmp_path / "knowledge.db"), "sources": [lambda since: batch[0]]},
                           vector_store=vs)
    assert r.sync() == {"added": 3, "updated": 0}
    assert r.get_document_count() == 3 and vs.get_count() == 9 and r.get_document_counts_by_type()["runbook"] == 1
    k = r.search("redis connection pool")
    assert set(k) == {"runbooks", "postmortems", "architecture", "knownIssues"}
    assert k["runbooks"] and all(c.documentId == "d1" for c in k["runbooks"])
    assert r.get_runbooks_for_service("api")["postmortems"] == []
    # an update with fewer chunks: the old vectors of d1 must be gone, not just overwritten
    batch[0] = [doc("d1", "runbook", "redis sentinel failover procedure", n=2)]
    assert r.sync() == {"added": 0, "updated": 1}
    assert vs.get_count() == 8 and not any(i == "vec_d1_2" for i in vs._ids if i)
    hits = vs.search("redis sentinel failover procedure", {"minScore": 0.2})
    assert hits and hits[0].documentId == "d1"
    r.close()
    embedder.reset()


def test_multi_device_index_equals_one_index(oracle_mod):
    """Single-process multi-GPU (multi_device.py) on CPU stand-ins: block-cyclic sharding over 3 'devices' gives
    the same (slots, fp64 scores, counts) as one index, through appends that straddle blocks, an overwrite,
    tombstones, ties across devices and a clear."""
    from runbookai_b200 import synth
    from runbookai_b200.multi_device import MultiDeviceIndex
    rng = np.random.default_rng(5)
    d = 48
    one = OracleIndex(d)
    multi = MultiDeviceIndex(d, devices=[0, 1, 2], block=64, index_factory=lambda dim, dev: OracleIndex(dim))
    q = synth.random_queries(4, d, 3).astype(np.float64)                 # bf16-exact values
    first_slots = []
    for n in (1, 63, 64, 130, 257, 5):                                  # runs that start and end mid-block
        rows = synth.bf16_bits_to_f32(synth.random_corpus(n, d, 100 + n)).astype(np.float64)
        first_slots.append((one.append_f64(rows), multi.append_f64(rows)))
    assert all(a == b for a, b in first_slots) and multi.size() == one.size() == 520
    dup = q[0] * 0.5                                                     # exact ties on different devices
    for s in (3, 70, 140, 300, 511):
        one.overwrite_f64(s, dup)
        multi.overwrite_f64(s, dup)
    dead = rng.choice(520, 40, replace=False)
    one.tombstone(dead)
    multi.tombstone(dead)
    assert multi.count() == one.count() == 480
    for k, ms in ((8, None), (16, 0.1), (32, 0.5)):
        a, b = one.search(q, k, ms), multi.search(q, k, ms)
        assert (a[2] == b[2]).all() and (a[0] == b[0]).all()
        assert np.array_equal(a[1], b[1], equal_nan=True)
    with pytest.raises(Exception):
        multi.search(np.zeros((1, d + 1)), 4, None)
    multi.clear()
    assert multi.size() == 0 and multi.search(q, 4, None)[2].tolist() == [0, 0, 0, 0]
    multi.close()


def test_large_limits_do_not_throw_and_match_the_reference_cut(store, tmp_path):
    """Reference call sites pass limit: 50 (infra-context.ts:229) and limit: 1000 (knowledge-context.ts:150): topK beyond
    the scan's candidate lists (2*topK > 112) must return what vector-store.ts:201-279 returns, not raise - through
    VectorStore.search, the micro-batcher and KnowledgeRetriever (which doubles topK again inside HybridRetriever)."""
    from oracle import pyref
    from runbookai_b200 import embedder
    from runbookai_b200.batcher import MicroBatcher
    from runbookai_b200.retriever import KnowledgeRetriever
    for d_i in range(8):
        store.add_chunks(_chunks(40, f"doc{d_i}", "runbook", ("api",), text=f"redis connection pool exhausted variant {d_i}"))
    q = "redis connection pool exhausted"
    rows = store.db.execute("SELECT id, embedding FROM vector_embeddings").fetchall()
    table = [(r["id"], np.frombuffer(r["embedding"], "<f8").tolist()) for r in rows]
    for top_k in (57, 100, 1000):
        res = store.search(q, {"topK": top_k, "minScore": 0.2})
        ref = pyref.vector_scan(embedder.embed_text(q), table, top_k=top_k, min_score=0.2)[:top_k]
        assert [f"vec_{r.id}" for r in res] == [i for i, _ in ref] and [r.score for r in res] == [s for _, s in ref]
    mb = MicroBatcher(store, window_ms=20.0)
    f_small, f_big = mb.submit(q, {"topK": 5, "minScore": 0.2}), mb.submit(q, {"topK": 300, "minScore": 0.2})
    assert f_big.result(timeout=30) == store.search(q, {"topK": 300, "minScore": 0.2})
    assert f_small.result(timeout=30) == store.search(q, {"topK": 5, "minScore": 0.2})
    mb.close()
    r = KnowledgeRetriever({"storePath": str(tmp_path / "kr" / "knowledge.db"), "sources": []}, vector_store=store)
    for limit in (50, 1000):
        k = r.search(q, {"limit": limit})
        assert set(k) == {"runbooks", "postmortems", "architecture", "knownIssues"} and len(k["runbooks"]) > 28
    r.store.close()


def test_mismatched_length_embedding_throws_only_while_it_is_in_the_map(store):
    """embedder.ts:169-171 throws while a wrong-length vector is stored; deleteDocument of that doc, or re-setting the
    id with a correct vector, makes search work again (the flag used to stick until clear())."""
    from runbookai_b200 import embedder
    from runbookai_b200._native import DimensionError
    store.add_chunks(_chunks(3, "docA"))
    q = "redis connection pool exhausted"
    assert len(store.search(q, {"minScore": 0.2})) == 3
    embedder.configure(HashEmbedder(48))                       # the embedder now returns another dimension
    store.add_chunk({"id": "docB_0", "documentId": "docB", "content": "redis pool", "sectionTitle": "S"}, "B", "runbook", [])
    store.add_chunk(_chunks(1, "docA")[0]["chunk"], "Title docA", "runbook", ["api"])   # re-set docA_0 with a bad length
    embedder.configure(HashEmbedder(64))
    with pytest.raises(DimensionError, match="Vectors must have the same length"):
        store.search(q, {"minScore": 0.2})
    store.delete_document("docB")                              # one offender gone, the re-set docA_0 still offends
    with pytest.raises(DimensionError):
        store.search(q, {"minScore": 0.2})
    store.add_chunk(_chunks(1, "docA")[0]["chunk"], "Title docA", "runbook", ["api"])   # re-set correctly: healed
    res = store.search(q, {"minScore": 0.2})
    assert sorted(r.id for r in res) == ["docA_0", "docA_1", "docA_2"]
    assert store._index.count() == 3                           # the stale vector of docA_0 was tombstoned, not kept


def test_failed_load_does_not_leave_a_half_built_shared_index(tmp_path):
    """If the bulk load of a shared store fails (e.g. cudaMalloc mid-append), the registry entry, the device index and
    the connection are dropped; the next opener loads from scratch instead of attaching to a half-filled index."""
    from runbookai_b200 import embedder
    from runbookai_b200.vector_store import create_vector_store, shared_index_count
    embedder.configure(HashEmbedder(64))
    base = str(tmp_path)
    s = create_vector_store(base, index_factory=lambda d, dev: OracleIndex(d))
    s.add_chunks(_chunks(5, "docA"))
    s.close()
    assert shared_index_count() == 0

    class Exploding(OracleIndex):
        def append_f64(self, rows):
            raise MemoryError("cudaMalloc failed")
    with pytest.raises(MemoryError):
        create_vector_store(base, index_factory=lambda d, dev: Exploding(d))
    assert shared_index_count() == 0
    s2 = create_vector_store(base, index_factory=lambda d, dev: OracleIndex(d))
    assert len(s2.search("redis connection pool exhausted", {"minScore": 0.2})) == 5
    s2.close()
    embedder.reset()


def test_micro_batcher_close_fails_queued_requests_and_rejects_new_ones(store):
    from runbookai_b200.batcher import MicroBatcher
    store.add_chunks(_chunks(3, "docA"))
    mb = MicroBatcher(store, window_ms=1.0)
    assert len(mb.search("redis connection pool exhausted", {"minScore": 0.2})) == 3
    mb.close()
    with pytest.raises(RuntimeError, match="batcher closed"):
        mb.submit("anything")
    # a request that slipped in behind the stop sentinel is failed, not left pending
    from concurrent.futures import Future
    fut = Future()
    mb._q.put(("late", {}, fut))
    mb.close()
    with pytest.raises(RuntimeError, match="batcher closed"):
        fut.result(timeout=1)


def test_reload_sidecar_is_used_when_current_and_ignored_when_stale(tmp_path):
    """`<db>.rbk` (one contiguous f64 matrix + id table, stamped with the table's row count and highest
    rowid) replaces the per-BLOB reload of vector-store.ts:56-66; any mutation by code that does not know about it
    (the reference itself: INSERT OR REPLACE / DELETE) makes it stale and the BLOBs are read again."""
    import sqlite3
    from runbookai_b200 import embedder
    from runbookai_b200.vector_store import VectorStore
    embedder.configure(HashEmbedder(64))
    path = str(tmp_path / "vectors.db")
    mk = lambda: VectorStore(path, index_factory=lambda d, dev: OracleIndex(d))
    s = mk()
    s.add_chunks(_chunks(6, "docA"))
    s.add_chunks(_chunks(4, "docB", text="kubernetes pod crashloop oom"))
    s.delete_document("docA")
    s.add_chunks(_chunks(3, "docA"))                       # re-added ids get new rowids: reload order = rowid order
    q = "redis connection pool exhausted"
    want = s.search(q, {"minScore": 0.1})
    s.close()
    assert os.path.exists(path + ".rbk")
    s2 = mk()
    assert s2.loaded_from_sidecar and s2.search(q, {"minScore": 0.1}) == want and s2._index.size() == 7
    assert [i for i in s2._ids] == [r[0] for r in s2.db.execute("SELECT id FROM vector_embeddings ORDER BY rowid")]
    s2.close()
    # the reference (or any foreign writer) changes the table: count / max rowid move -> sidecar ignored, then rewritten
    db = sqlite3.connect(path)
    db.execute("DELETE FROM vector_embeddings WHERE id = 'vec_docB_0'")
    db.commit()
    db.close()
    s3 = mk()
    assert not s3.loaded_from_sidecar and s3._index.size() == 6
    s3.close()
    s4 = mk()
    assert s4.loaded_from_sidecar and s4._index.size() == 6
    # SQLite hands the rowids of deleted LAST rows out again: deleting the newest document and adding one of the same
    # size leaves the row count and the highest rowid unchanged.  (a) through this class: the generation counter in
    # the db header moves
    fp = s4._fingerprint()
    s4.delete_document("docA")
    s4.add_chunks(_chunks(3, "docA", text="completely different words about dns"))
    fp2 = s4._fingerprint()
    assert fp2[:2] == fp[:2] and fp2[2] == fp[2] + 2 and fp2[3] != fp[3]
    want4 = s4.search("completely different words about dns", {"minScore": 0.1})
    s4.close()                                             # stale by generation -> rewritten
    s4b = mk()
    assert s4b.loaded_from_sidecar and s4b.search("completely different words about dns", {"minScore": 0.1}) == want4
    s4b.close()
    # (b) a foreign writer that knows nothing about the counter does the same: the tail hash catches it
    db = sqlite3.connect(path)
    blob = db.execute("SELECT embedding FROM vector_embeddings WHERE id = 'vec_docB_1'").fetchone()[0]
    row = db.execute("SELECT * FROM vector_embeddings ORDER BY rowid DESC LIMIT 1").fetchone()
    db.execute("DELETE FROM vector_embeddings WHERE id = ?", (row[0],))
    db.execute("INSERT INTO vector_embeddings (id, chunk_id, document_id, embedding, content, title, type, services) "
               "VALUES (?, ?, ?, ?, ?, ?, ?, ?)", (row[0], row[1], row[2], blob, row[4], row[5], row[6], row[7]))
    db.commit()
    db.close()
    s4c = mk()
    assert s4c._fingerprint()[:3] == fp2[:3] and not s4c.loaded_from_sidecar
    s4c.close()
    os.environ["RUNBOOK_KNN_SIDECAR"] = "0"
    try:
        s5 = mk()
        assert not s5.loaded_from_sidecar and s5.search(q, {"minScore": 0.1})
        s5.close()
    finally:
        del os.environ["RUNBOOK_KNN_SIDECAR"]
    embedder.reset()
