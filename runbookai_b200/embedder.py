"""Host mirror of src/knowledge/indexer/embedder.ts for the parts on the kNN path.

* cosine_similarity / find_most_similar (embedder.ts:168-202) run on the GPU through the
  C ABI (a throw-away index) — there is deliberately no CPU arithmetic here.
* The text->vector step (embedText/embedTexts, embedder.ts:57-163) is an HTTPS call to
  OpenAI in the reference and is OUT of the kernel scope.  It is kept
  as a pluggable `Embedder`; `configure()` / `is_embedder_configured()` keep the
  reference's module-level configuration semantics (embedder.ts:27-44).
"""
from __future__ import annotations

import hashlib
import json
import os
import urllib.request
from typing import Protocol, Sequence

import numpy as np

from ._native import DimensionError, Index, RBK_EDIM

OPENAI_API_BASE = "https://api.openai.com/v1"
DEFAULT_MODEL = "text-embedding-3-small"
DEFAULT_BATCH_SIZE = 100
DEFAULT_DIMENSIONS = 1536


class Embedder(Protocol):
    def embed_text(self, text: str) -> list[float]: ...
    def embed_texts(self, texts: Sequence[str]) -> list[list[float]]: ...


class OpenAIEmbedder:
    """embedder.ts:57-163: POST /v1/embeddings, md5-keyed in-memory cache, batches of 100."""

    def __init__(self, api_key: str, model: str = DEFAULT_MODEL, batch_size: int = DEFAULT_BATCH_SIZE,
                 dimensions: int = DEFAULT_DIMENSIONS):
        self.api_key, self.model, self.batch_size, self.dimensions = api_key, model, batch_size, dimensions
        self._cache: dict[str, list[float]] = {}

    def _key(self, text: str) -> str:
        return f"{self.model}:{hashlib.md5(text.encode('utf-8')).hexdigest()}"

    def _post(self, inputs):
        body = json.dumps({"model": self.model, "input": inputs, "dimensions": self.dimensions}).encode()
        req = urllib.request.Request(f"{OPENAI_API_BASE}/embeddings", data=body, method="POST", headers={
            "Content-Type": "application/json", "Authorization": f"Bearer {self.api_key}"})
        with urllib.request.urlopen(req) as resp:  # noqa: S310 (fixed https URL)
            data = json.loads(resp.read())
        return [d["embedding"] for d in sorted(data["data"], key=lambda d: d["index"])]

    def embed_text(self, text: str) -> list[float]:
        k = self._key(text)
        if k not in self._cache:
            self._cache[k] = self._post(text)[0]
        return self._cache[k]

    def embed_texts(self, texts: Sequence[str]) -> list[list[float]]:
        out: list[list[float] | None] = [self._cache.get(self._key(t)) for t in texts]
        todo = [i for i, e in enumerate(out) if e is None]
        for b0 in range(0, len(todo), self.batch_size):
            idx = todo[b0:b0 + self.batch_size]
            for i, e in zip(idx, self._post([texts[i] for i in idx])):
                self._cache[self._key(texts[i])] = e
                out[i] = e
        return out  # type: ignore[return-value]


_config: Embedder | None = None


def configure(api_key_or_embedder, **options) -> None:
    """embedder.ts:27-34.  Accepts an API key (OpenAI) or any Embedder object."""
    global _config
    _config = OpenAIEmbedder(api_key_or_embedder, **options) if isinstance(api_key_or_embedder, str) \
        else api_key_or_embedder


def reset() -> None:
    global _config
    _config = None


def is_embedder_configured() -> bool:
    """embedder.ts:42-44."""
    return _config is not None or bool(os.environ.get("OPENAI_API_KEY"))


def get_embedder() -> Embedder:
    global _config
    if _config is None:
        key = os.environ.get("OPENAI_API_KEY")
        if not key:
            raise RuntimeError("OpenAI API key not configured. Set OPENAI_API_KEY environment variable.")
        _config = OpenAIEmbedder(key)
    return _config


def embed_text(text: str) -> list[float]:
    return get_embedder().embed_text(text)


def embed_texts(texts: Sequence[str]) -> list[list[float]]:
    return get_embedder().embed_texts(texts)


def find_most_similar(query_embedding, embeddings, top_k: int = 10, device: int = 0):
    """embedder.ts:189-202 on the GPU: no threshold, stable descending order, first top_k.

    embeddings: sequence of {"id":..., "embedding":[...]} (or (id, vector) pairs).
    """
    items = [(e["id"], e["embedding"]) if isinstance(e, dict) else (e[0], e[1]) for e in embeddings]
    q = np.asarray(query_embedding, dtype=np.float64)
    if any(len(v) != q.shape[0] for _, v in items):
        raise DimensionError(RBK_EDIM, "Vectors must have the same length")
    if not items:
        return []
    out = []
    with Index(q.shape[0], device=device, capacity_hint=len(items), keep_f64=True) as ix:
        ix.append_f64(np.asarray([v for _, v in items], dtype=np.float64))
        for k0 in range(0, top_k, 112):  # the ABI serves at most 112 per call; page by exclusion
            k = min(112, top_k - k0)
            slots, scores, counts, _ = ix.search(q, k, None)
            hits = [(int(s), float(v)) for s, v in zip(slots[0, :counts[0]], scores[0, :counts[0]])]
            out.extend({"id": items[s][0], "score": v} for s, v in hits)
            if counts[0] < k:
                break
            ix.tombstone([s for s, _ in hits])
    return out[:top_k]


def cosine_similarity(a, b, device: int = 0) -> float:
    """embedder.ts:168-184 (one pair), evaluated by the engine's exact fp64 re-rank kernel."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape or a.ndim != 1:
        raise DimensionError(RBK_EDIM, "Vectors must have the same length")
    r = find_most_similar(a, [("b", b)], 1, device)
    return r[0]["score"] if r else float("nan")
