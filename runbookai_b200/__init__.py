"""runbookai_b200 — B200-native kNN engine behind RunbookAI's knowledge-base vector search.

Host-side mirror of the reference interface for this path (the reference is TypeScript;
Node is absent from the build image, so the mirror is Python over the same C ABI the
N-API addon binds — see INTEGRATION.md):

    VectorStore / create_vector_store   <- src/knowledge/store/vector-store.ts
    HybridRetriever / reciprocal_rank_fusion <- src/knowledge/retriever/hybrid-search.ts
    cosine_similarity / find_most_similar    <- src/knowledge/indexer/embedder.ts:168-202

Importing the package loads runbookai_b200/lib/librbk_knn.so and fails loudly if it has
not been built.
"""
from ._native import DimensionError, Index, RbkError  # noqa: F401

__all__ = ["Index", "RbkError", "DimensionError"]
