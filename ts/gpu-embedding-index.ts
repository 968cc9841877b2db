/**
 * GpuEmbeddingIndex — device-resident replacement for the `Map<string, number[]>` that
 * `VectorStore` keeps in RAM (src/knowledge/store/vector-store.ts:26), plus the one operation the
 * reference performs on it in its hot loop (:207-221): "all ids whose cosine with the query is
 * >= minScore, best first, stable, first 2*topK".
 *
 * NOT COMPILED in the build image (no Node/tsc).  It is deliberately tiny: the Map's ordered-key
 * semantics (insertion slot, re-set keeps the slot, delete frees the key but never the slot) on top
 * of the N-API addon (napi/rbk_napi.cc -> include/rbk_knn.h).  The tested mirror of the same
 * bookkeeping is runbookai_b200/vector_store.py (`_set`, `delete_document`, `_load_embeddings`).
 * INTEGRATION.md shows the few lines of vector-store.ts that change to use it.
 */
// eslint-disable-next-line @typescript-eslint/no-var-requires
const { RbkIndex } = require('../native/build/Release/rbk_knn.node');

export interface ScoredId {
  id: string;
  score: number;
}

export class GpuEmbeddingIndex {
  private index: any | null = null;
  private idOfSlot: (string | null)[] = [];
  private slotOfId = new Map<string, number>();
  private dim = 0;
  private mixedLengths = false;

  constructor(private device = Number(process.env.RUNBOOK_KNN_DEVICE ?? 0)) {}

  get size(): number {
    return this.slotOfId.size;
  }

  /** Bulk load at construction: rows are the f64-LE BLOBs exactly as SQLite returns them. */
  loadBlobs(rows: Array<{ id: string; embedding: Buffer }>): void {
    if (rows.length === 0) return;
    this.dim = rows[0].embedding.length / 8;
    const usable = rows.filter((r) => r.embedding.length === this.dim * 8);
    this.mixedLengths = usable.length !== rows.length;
    this.index = new RbkIndex(this.dim, this.device, usable.length);
    const packed = new Float64Array(usable.length * this.dim);
    usable.forEach((r, i) => {
      packed.set(new Float64Array(r.embedding.buffer, r.embedding.byteOffset, this.dim), i * this.dim);
    });
    const first = Number(this.index.appendF64(packed));
    usable.forEach((r, i) => this.remember(r.id, first + i));
  }

  /** Map.set: an existing key keeps its place in iteration order, a new key goes last. */
  set(id: string, embedding: number[]): void {
    if (!this.index) {
      this.dim = embedding.length;
      this.index = new RbkIndex(this.dim, this.device, 0);
    }
    if (embedding.length !== this.dim) {
      this.mixedLengths = true; // the reference would store it and throw on the next search
      return;
    }
    const row = Float64Array.from(embedding);
    const slot = this.slotOfId.get(id);
    if (slot !== undefined) this.index.overwriteF64(slot, row);
    else this.remember(id, Number(this.index.appendF64(row)));
  }

  /** Map.delete for a batch of keys. */
  deleteMany(ids: string[]): void {
    const slots: bigint[] = [];
    for (const id of ids) {
      const slot = this.slotOfId.get(id);
      if (slot === undefined) continue;
      this.slotOfId.delete(id);
      this.idOfSlot[slot] = null;
      slots.push(BigInt(slot));
    }
    if (slots.length > 0 && this.index) this.index.tombstone(BigInt64Array.from(slots));
  }

  clear(): void {
    this.idOfSlot = [];
    this.slotOfId.clear();
    this.mixedLengths = false;
    this.index?.clear();
  }

  /** The scan + threshold + stable sort + cut, for one query embedding. */
  async best(query: number[], limit: number, minScore: number): Promise<ScoredId[]> {
    if (!this.index || this.slotOfId.size === 0) return [];
    if (this.mixedLengths || query.length !== this.dim) throw new Error('Vectors must have the same length');
    const { slots, scores, counts } = await this.index.search(Float64Array.from(query), 1, limit, minScore);
    const out: ScoredId[] = [];
    for (let i = 0; i < counts[0]; i++) out.push({ id: this.idOfSlot[Number(slots[i])]!, score: scores[i] });
    return out;
  }

  private remember(id: string, slot: number): void {
    this.slotOfId.set(id, slot);
    this.idOfSlot[slot] = id;
  }
}
