/**
 * GPU-backed drop-in for src/knowledge/store/vector-store.ts.
 *
 * NOT COMPILED in the build image (no Node/tsc): this is the class a maintainer drops next to
 * the original.  Same schema, same public surface, same quirks; the only change is the hot
 * loop of search() (vector-store.ts:207-221), which becomes one call into the N-API addon
 * (napi/rbk_napi.cc -> include/rbk_knn.h).  The tested mirror of exactly this logic is
 * runbookai_b200/vector_store.py.
 */
import Database from 'better-sqlite3';
import { embedText, embedTexts, isEmbedderConfigured } from '../indexer/embedder';
import type { KnowledgeChunk, KnowledgeType, RetrievedChunk } from '../types';
// eslint-disable-next-line @typescript-eslint/no-var-requires
const { RbkIndex } = require('../../../native/build/Release/rbk_knn.node');

export class VectorStore {
  private db: Database.Database;
  private index: any | null = null;
  private ids: (string | null)[] = []; // slot -> id (null = deleted)
  private slotOf = new Map<string, number>(); // the reference's Map keys, with their insertion slot
  private ragged = false;

  constructor(dbPath: string, private device = Number(process.env.RUNBOOK_KNN_DEVICE ?? 0)) {
    this.db = new Database(dbPath);
    this.initSchema(); // identical SQL to vector-store.ts:34-51
    this.loadEmbeddings();
  }

  private initSchema(): void {
    this.db.exec(`
      CREATE TABLE IF NOT EXISTS vector_embeddings (
        id TEXT PRIMARY KEY, chunk_id TEXT NOT NULL, document_id TEXT NOT NULL, embedding BLOB NOT NULL,
        content TEXT NOT NULL, title TEXT, type TEXT NOT NULL, services TEXT,
        created_at TEXT DEFAULT CURRENT_TIMESTAMP);
      CREATE INDEX IF NOT EXISTS idx_vector_document_id ON vector_embeddings(document_id);
      CREATE INDEX IF NOT EXISTS idx_vector_type ON vector_embeddings(type);`);
  }

  /** vector-store.ts:56-66 — rows in rowid order go to the device; BLOBs are already f64 LE. */
  private loadEmbeddings(): void {
    const rows = this.db.prepare('SELECT id, embedding FROM vector_embeddings').all() as Array<{
      id: string;
      embedding: Buffer;
    }>;
    if (rows.length === 0) return;
    const dim = rows[0].embedding.length / 8;
    const good = rows.filter((r) => r.embedding.length === dim * 8);
    this.ragged = good.length !== rows.length;
    this.index = new RbkIndex(dim, this.device, good.length);
    const packed = new Float64Array(good.length * dim);
    good.forEach((r, i) =>
      packed.set(new Float64Array(r.embedding.buffer, r.embedding.byteOffset, dim), i * dim)
    );
    this.index.appendF64(packed);
    good.forEach((r) => {
      this.slotOf.set(r.id, this.ids.length);
      this.ids.push(r.id);
    });
  }

  /** `this.embeddings.set(id, e)`: an existing key keeps its position, a new one is appended. */
  private set(id: string, embedding: number[]): void {
    if (!this.index) this.index = new RbkIndex(embedding.length, this.device, 0);
    const slot = this.slotOf.get(id);
    if (slot !== undefined) this.index.overwriteF64(slot, Float64Array.from(embedding));
    else {
      this.slotOf.set(id, this.index.appendF64(Float64Array.from(embedding)));
      this.ids.push(id);
    }
  }

  async search(
    query: string,
    options: { topK?: number; typeFilter?: KnowledgeType[]; serviceFilter?: string[]; minScore?: number } = {}
  ): Promise<RetrievedChunk[]> {
    if (!isEmbedderConfigured()) throw new Error('Embedder not configured. Set OPENAI_API_KEY.');
    const topK = options.topK || 10;
    const minScore = options.minScore || 0.5;
    const queryEmbedding = await embedText(query);
    if (!this.index || this.ids.length === 0) return [];
    if (this.ragged) throw new Error('Vectors must have the same length');
    // vector-store.ts:207-221 — scan + `>= minScore` + stable sort + slice(0, 2*topK): one GPU call.
    // A wrong query length rejects with the reference's message (RBK_EDIM).
    const { slots, scores, counts } = await this.index.search(
      Float64Array.from(queryEmbedding), 1, topK * 2, minScore);
    const n = counts[0];
    const topIds = Array.from(slots.subarray(0, n), (s: bigint) => this.ids[Number(s)]!);
    if (topIds.length === 0) return [];
    // vector-store.ts:227-279 unchanged from here on
    let sql = `SELECT id, chunk_id, document_id, content, title, type, services FROM vector_embeddings
               WHERE id IN (${topIds.map(() => '?').join(',')})`;
    const params: (string | number)[] = [...topIds];
    if (options.typeFilter && options.typeFilter.length > 0) {
      sql += ` AND type IN (${options.typeFilter.map(() => '?').join(',')})`;
      params.push(...options.typeFilter);
    }
    const rows = this.db.prepare(sql).all(...params) as any[];
    const scoreMap = new Map(topIds.map((id, i) => [id, scores[i]]));
    const results: RetrievedChunk[] = [];
    for (const row of rows) {
      const services = JSON.parse(row.services || '[]') as string[];
      if (options.serviceFilter && options.serviceFilter.length > 0) {
        if (!options.serviceFilter.some((s) => services.includes(s))) continue;
      }
      results.push({
        id: row.chunk_id, documentId: row.document_id, title: row.title || '', content: row.content,
        type: row.type as KnowledgeType, services, score: scoreMap.get(row.id) || 0,
      });
    }
    results.sort((a, b) => b.score - a.score);
    return results.slice(0, topK);
  }

  async addChunk(chunk: KnowledgeChunk, documentTitle: string, type: KnowledgeType, services: string[]) {
    if (!isEmbedderConfigured()) throw new Error('Embedder not configured. Set OPENAI_API_KEY.');
    const embedding = await embedText([documentTitle, chunk.sectionTitle, chunk.content].filter(Boolean).join('\n\n'));
    const id = `vec_${chunk.id}`;
    const buf = Buffer.from(Float64Array.from(embedding).buffer);
    this.db.prepare(`INSERT OR REPLACE INTO vector_embeddings
      (id, chunk_id, document_id, embedding, content, title, type, services) VALUES (?, ?, ?, ?, ?, ?, ?, ?)`)
      .run(id, chunk.id, chunk.documentId, buf, chunk.content, chunk.sectionTitle || documentTitle, type,
           JSON.stringify(services));
    this.set(id, embedding);
  }

  async addChunks(chunks: Array<{ chunk: KnowledgeChunk; documentTitle: string; type: KnowledgeType; services: string[] }>) {
    if (!isEmbedderConfigured()) throw new Error('Embedder not configured. Set OPENAI_API_KEY.');
    const embeddings = await embedTexts(chunks.map((c) =>
      [c.documentTitle, c.chunk.sectionTitle, c.chunk.content].filter(Boolean).join('\n\n')));
    const stmt = this.db.prepare(`INSERT OR REPLACE INTO vector_embeddings
      (id, chunk_id, document_id, embedding, content, title, type, services) VALUES (?, ?, ?, ?, ?, ?, ?, ?)`);
    this.db.transaction(() => {
      chunks.forEach(({ chunk, documentTitle, type, services }, i) => {
        stmt.run(`vec_${chunk.id}`, chunk.id, chunk.documentId, Buffer.from(Float64Array.from(embeddings[i]).buffer),
                 chunk.content, chunk.sectionTitle || documentTitle, type, JSON.stringify(services));
      });
    })();
    chunks.forEach(({ chunk }, i) => this.set(`vec_${chunk.id}`, embeddings[i]));
  }

  deleteDocument(documentId: string): void {
    const rows = this.db.prepare('SELECT id FROM vector_embeddings WHERE document_id = ?').all(documentId) as Array<{ id: string }>;
    const slots: bigint[] = [];
    for (const row of rows) {
      const s = this.slotOf.get(row.id);
      if (s !== undefined) {
        this.slotOf.delete(row.id);
        this.ids[s] = null;
        slots.push(BigInt(s));
      }
    }
    if (slots.length && this.index) this.index.tombstone(BigInt64Array.from(slots));
    this.db.prepare('DELETE FROM vector_embeddings WHERE document_id = ?').run(documentId);
  }

  getCount(): number {
    return (this.db.prepare('SELECT COUNT(*) as count FROM vector_embeddings').get() as { count: number }).count;
  }

  hasDocument(documentId: string): boolean {
    return (this.db.prepare('SELECT COUNT(*) as count FROM vector_embeddings WHERE document_id = ?')
      .get(documentId) as { count: number }).count > 0;
  }

  clear(): void {
    this.db.exec('DELETE FROM vector_embeddings');
    this.ids = [];
    this.slotOf.clear();
    this.ragged = false;
    this.index?.clear();
  }

  close(): void {
    this.db.close();
    this.index = null; // the addon's finalizer calls rbk_index_destroy
  }
}

export function createVectorStore(baseDir: string = '.runbook'): VectorStore {
  return new VectorStore(`${baseDir}/vectors.db`);
}
