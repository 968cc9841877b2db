"""GPU suite (-m gpu), part 3: the §8(f) rows on the REAL device index - HybridRetriever + RRF (f-1), the
createRetriever variant with its sync hook, the per-dbPath shared index and the reload sidecar (f-2) - checked against
the pure-Python restatements of vector-store.ts:201-221 and hybrid-search.ts:106-151 (oracle/pyref.py).

The documents are the three markdown fixtures of the reference's MCP server test
(src/mcp/__tests__/server.test.ts:23-94: API Troubleshooting Guide, Database Connection Issues, Redis Memory Spike
Issue), chunked by their `##` sections the way the reference's markdown source does.  The embedder is the offline
hash embedder (the reference calls OpenAI); vectors are arbitrary float64, so the index runs with its f64 sidecar."""
import numpy as np
import pytest

from common import HashEmbedder

pytestmark = pytest.mark.gpu


class RawHashEmbedder(HashEmbedder):      # arbitrary float64 vectors, like real embeddings
    def embed_text(self, text):
        words = [w for w in text.lower().replace("\n", " ").split() if w]
        return np.sum([self._word(w) for w in words], axis=0).tolist() if words else [0.0] * self.dim


DOCS = [
    {"id": "api-troubleshooting", "type": "runbook", "title": "API Troubleshooting Guide", "services": ["api", "gateway"],
     "chunks": [{"id": "api-troubleshooting_0", "sectionTitle": "Symptoms", "content": "HTTP 500 errors High latency"},
                {"id": "api-troubleshooting_1", "sectionTitle": "Steps",
                 "content": "Check logs Verify database connections Check memory usage"}]},
    {"id": "database-issues", "type": "runbook", "title": "Database Connection Issues", "services": ["database", "postgres"],
     "chunks": [{"id": "database-issues_0", "sectionTitle": "Root Causes",
                 "content": "Connection pool exhaustion Network issues High query load"},
                {"id": "database-issues_1", "sectionTitle": "Resolution",
                 "content": "Check connection pool metrics Review slow queries"}]},
    {"id": "known-issue", "type": "known_issue", "title": "Redis Memory Spike Issue", "services": ["cache", "redis"],
     "chunks": [{"id": "known-issue_0", "sectionTitle": "Redis Memory Spike",
                 "content": "Known issue with Redis memory management."},
                {"id": "known-issue_1", "sectionTitle": "Workaround", "content": "Restart Redis pods during low traffic."}]},
]


@pytest.fixture
def rb(native):
    import torch
    assert torch.cuda.is_available(), "run -m gpu on a GPU box"
    from runbookai_b200 import embedder
    embedder.configure(RawHashEmbedder(96))
    yield
    embedder.reset()


def vector_table(vs):
    rows = vs.db.execute("SELECT id, embedding FROM vector_embeddings").fetchall()
    return [(r["id"], np.frombuffer(r["embedding"], "<f8").tolist()) for r in rows]


def test_create_retriever_sync_vector_and_hybrid_modes_on_gpu(rb, tmp_path):
    from oracle import pyref
    from runbookai_b200 import embedder
    from runbookai_b200._native import Index
    from runbookai_b200.retriever import create_retriever
    from runbookai_b200.vector_store import shared_index_count
    base = str(tmp_path / ".runbook")
    r = create_retriever(base, sources=[lambda since: DOCS])
    assert r.sync() == {"added": 3, "updated": 0}
    vs = r.vector_store
    assert isinstance(vs._index, Index) and vs.get_count() == 6 and shared_index_count() == 1
    table = vector_table(vs)
    for q in ("database connection pool exhaustion", "redis memory spike restart pods", "HTTP 500 errors high latency"):
        for top_k, ms in ((3, 0.1), (5, 0.3), (2, None)):
            got = r._hybrid.search(q, {"mode": "vector", "topK": top_k, **({"minScore": ms} if ms else {})})
            # hybrid-search.ts:78-83 passes no minScore: the vector leg runs with the store's default 0.5
            ref = pyref.vector_scan(embedder.embed_text(q), table, top_k=top_k, min_score=None)[:top_k]
            assert [f"vec_{c.id}" for c in got] == [i for i, _ in ref] and [c.score for c in got] == [s for _, s in ref]
        # hybrid = RRF over (FTS top 2k, vector top 2k), keyed by chunk id (hybrid-search.ts:85-100, 106-151)
        fts = r.store.search(q, {"limit": 8})
        vec = vs.search(q, {"topK": 8})
        want = pyref.rrf([c.id for c in fts], [c.id for c in vec], 4)
        got = r._hybrid.search(q, {"topK": 4})
        assert [(c.id, c.score) for c in got] == want
    k = r.search("redis memory spike")
    assert [c.documentId for c in k["knownIssues"]][:1] == ["known-issue"] and k["postmortems"] == []
    assert all(c.type == "runbook" for c in r.get_runbooks_for_service("database")["runbooks"])
    # limit: 50 and 1000 (infra-context.ts:229, knowledge-context.ts:150) go through the large-k path, on the device
    assert sum(len(v) for v in r.search("connection pool", {"limit": 1000}).values()) >= 1
    # an updated document replaces its vectors (sync hook of SURVEY 8f-2)
    DOCS2 = [dict(DOCS[1], chunks=[{"id": "database-issues_0", "sectionTitle": "Root Causes",
                                    "content": "Replication lag after failover"}])]
    r.config["sources"] = [lambda since: DOCS2]
    assert r.sync() == {"added": 0, "updated": 1} and vs.get_count() == 5
    assert vs.search("replication lag failover", {"minScore": 0.3})[0].id == "database-issues_0"
    r.close()
    assert shared_index_count() == 0


def test_shared_device_index_single_upload_and_sidecar_reload_on_gpu(rb, tmp_path):
    """Two stores on one db path share ONE device index (one upload); after the last close the next opener reloads
    from the sidecar and answers identically."""
    from runbookai_b200.vector_store import create_vector_store, shared_index_count
    base = str(tmp_path)
    a = create_vector_store(base)
    for t_i, topic in enumerate(["redis connection pool exhausted", "kubernetes pod crashloop oom",
                                 "postgres replication lag"]):
        a.add_chunks([{"chunk": {"id": f"d{t_i}_{i}", "documentId": f"d{t_i}", "content": f"{topic} step {i}",
                                 "sectionTitle": f"S{i}"}, "documentTitle": topic.title(), "type": "runbook",
                       "services": ["api"]} for i in range(50)])
    b = create_vector_store(base)                          # attaches: same index object, nothing re-uploaded
    assert b._index is a._index and shared_index_count() == 1
    launches = a._index.stats()["kernel_launches"]
    q = "redis pool exhausted"
    want = a.search(q, {"topK": 7, "minScore": 0.2})
    assert b.search(q, {"topK": 7, "minScore": 0.2}) == want
    a.delete_document("d1")                                # visible through the other instance at once
    assert all(c.documentId != "d1" for c in b.search("pod crashloop oom", {"minScore": 0.1}))
    assert a._index.stats()["kernel_launches"] > launches
    a.close()
    assert shared_index_count() == 1                       # b still holds it
    want = b.search(q, {"topK": 7, "minScore": 0.2})
    b.close()
    assert shared_index_count() == 0
    c = create_vector_store(base)
    assert c.loaded_from_sidecar and c._index.size() == 100 and c.search(q, {"topK": 7, "minScore": 0.2}) == want
    c.close()
