/*
 * rbk_knn.h — C ABI of the B200-native kNN engine behind RunbookAI's VectorStore.
 *
 * The reference has no FFI for this path; the seam is the TypeScript class
 * `VectorStore` (src/knowledge/store/vector-store.ts:24-333), whose hot loop is
 *     for (const [id, embedding] of this.embeddings)            (:210-215)
 *         score = cosineSimilarity(queryEmbedding, embedding)   (embedder.ts:168-184)
 *         if (score >= minScore) scored.push({id, score})
 *     scored.sort((a,b) => b.score - a.score)                   (:218)
 *     scored.slice(0, topK * 2)                                 (:221)
 * Each entry point below names the reference lines it replaces.  A Node N-API addon
 * (napi/rbk_napi.cc) or any other FFI binds exactly these symbols; INTEGRATION.md shows
 * the binding.  Plain C types only: no C++/torch/CUDA types cross this boundary
 * (CUDA streams and device pointers travel as void*).
 *
 * Conventions
 *   - "slot" = dense insertion index of a row = the reference's Map insertion position
 *     (SURVEY.md §8c S6/S9b).  The host side keeps the slot <-> "vec_<chunkId>" table.
 *   - Every function returns rbk_status; on failure rbk_last_error() holds the message
 *     the binding turns into `new Error(msg)`.  Nothing throws or aborts.
 *   - Inputs are borrowed for the duration of the call; outputs are caller-allocated.
 *   - An index is bound to ONE GPU.  A corpus sharded over several GPUs is either a
 *     group - rbk_group_*: one process, one handle, NCCL all-gather inside the search call - or one
 *     index per GPU in one process per GPU plus rbk_merge_topk_packed_device() after the
 *     caller's own exchange of the per-shard blocks.
 *   - Thread safety: calls on the same index are serialised by an internal mutex;
 *     different indexes are independent.
 *   - There is NO CPU fallback: without a CUDA device rbk_index_create fails with
 *     RBK_ECUDA.
 */
#ifndef RBK_KNN_H
#define RBK_KNN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RBK_ABI_VERSION 2
/* Largest k_fetch a single search accepts (the scan keeps k_fetch + margin <= 128). */
#define RBK_MAX_K_FETCH 112

typedef struct rbk_index rbk_index;

typedef enum {
  RBK_OK = 0,
  RBK_EINVAL = 1, /* bad argument */
  RBK_ENOMEM = 2, /* host or device allocation failed */
  RBK_ECUDA = 3,  /* CUDA runtime/driver error, or no device */
  RBK_ENCCL = 4,  /* NCCL missing or failing (rbk_group_* with more than one GPU) */
  RBK_EDIM = 5    /* "Vectors must have the same length" (embedder.ts:169-171) */
} rbk_status;

int rbk_abi_version(void);
/* Message of the last failure on this thread (valid until the next failing call). */
const char* rbk_last_error(void);

/* ---- lifetime: `new VectorStore(dbPath)` / `close()` (vector-store.ts:28-32, :330) ---- */
/* dim: embedding length (any d >= 1).  device: CUDA ordinal.  capacity_hint: rows to
 * pre-allocate (0 = default); the index grows on demand. */
rbk_status rbk_index_create(int32_t dim, int32_t device, int64_t capacity_hint, rbk_index** out);
/* flags: RBK_INDEX_KEEP_F64 keeps, next to the bf16 rows the scan reads, the ORIGINAL values of every appended
 * row as float64 (8*dim bytes per row; f32/bf16 inputs are widened exactly).  The exact re-rank then uses them:
 * results are the reference's fp64 cosine bit for bit for ARBITRARY float64 embeddings (the SQLite BLOBs of
 * vector-store.ts:71-88), not only for bf16-representable ones.  The scan's error bound grows by the largest
 * angle between a row and its bf16 rounding, so more queries may take the exhaustive path on near-tied data. */
#define RBK_INDEX_KEEP_F64 1u
rbk_status rbk_index_create_ex(int32_t dim, int32_t device, int64_t capacity_hint, uint32_t flags, rbk_index** out);
void rbk_index_destroy(rbk_index* idx); /* NULL is a no-op */

/* Run all device work of this index on the given cudaStream_t (NULL = the index's own
 * stream).  Lets a host framework time the engine with events on its current stream. */
rbk_status rbk_index_set_stream(rbk_index* idx, void* cuda_stream);
/* Global slot of local row 0 (row-sharded corpora).  Default 0. */
rbk_status rbk_index_set_slot_base(rbk_index* idx, int64_t slot_base);

/* ---- mutation: loadEmbeddings / addChunk(s) / deleteDocument / clear
 *      (vector-store.ts:56-66, :93-183, :285-297, :322-325) ---- */
/* rows: n_rows x dim, row-major.  f64 is the SQLite BLOB layout (little-endian float64,
 * vector-store.ts:71-88).  Values are stored as bf16 (round-to-nearest-even); the index
 * is exact for inputs representable in bf16.  first_slot_out (nullable)
 * receives the LOCAL slot of the first appended row.  `rows` (host memory, pageable or
 * page-locked, or device memory for the *_device variants) has been consumed when the
 * call returns; the conversion itself may still be running on the index's stream. */
rbk_status rbk_index_append_f64(rbk_index* idx, const double* rows, int64_t n_rows, int64_t* first_slot_out);
rbk_status rbk_index_append_f32(rbk_index* idx, const float* rows, int64_t n_rows, int64_t* first_slot_out);
rbk_status rbk_index_append_bf16(rbk_index* idx, const uint16_t* rows, int64_t n_rows, int64_t* first_slot_out);
/* Same, rows already in device memory on the index's GPU (bulk load without a PCIe hop). */
rbk_status rbk_index_append_bf16_device(rbk_index* idx, const void* dev_rows, int64_t n_rows,
                                        int64_t* first_slot_out);
/* f64 rows (the BLOB layout) already on the device, e.g. a sidecar file read with GPUDirect or staged by the host
 * framework in large pinned chunks. */
rbk_status rbk_index_append_f64_device(rbk_index* idx, const void* dev_rows, int64_t n_rows, int64_t* first_slot_out);
/* `this.embeddings.set(id, e)` on an existing id keeps its Map position (S9b). */
rbk_status rbk_index_overwrite_f64(rbk_index* idx, int64_t local_slot, const double* row);
/* The same for n rows at once - a re-embedded document (addChunks over existing ids, vector-store.ts:135-183):
 * rows[i] (n x dim, f64) replaces local_slots[i].  One call, one host round trip for the whole batch.  A slot that
 * is tombstoned stays dead and makes the call return RBK_EINVAL after the LIVE slots of the batch have been
 * written (the host mirror never overwrites a deleted id, so this is a caller bug, not a data path).  A slot named
 * more than once takes its LAST row, as Map.set twice would. */
rbk_status rbk_index_overwrite_f64_batch(rbk_index* idx, const int64_t* local_slots, int64_t n, const double* rows);
/* `this.embeddings.delete(id)`: the rows stop matching; slots are not reused. */
rbk_status rbk_index_tombstone(rbk_index* idx, const int64_t* local_slots, int64_t n);
rbk_status rbk_index_clear(rbk_index* idx);
int64_t rbk_index_count(const rbk_index* idx); /* live rows  */
int64_t rbk_index_size(const rbk_index* idx);  /* slots used, tombstones included */
int32_t rbk_index_dim(const rbk_index* idx);
/* Copy stored rows back (bf16 bits), for tests and for reload sidecars. */
rbk_status rbk_index_read_rows_bf16(rbk_index* idx, int64_t first_local_slot, int64_t n_rows, uint16_t* out);

/* ---- search: the scan + sort + cut of VectorStore.search (vector-store.ts:207-221)
 *      and findMostSimilar (embedder.ts:189-202), batched over B queries ---- */
/*
 * queries: B x query_dim row-major, HOST memory.  query_dim != dim -> RBK_EDIM, message
 * "Vectors must have the same length".  For each query b the call returns the first
 * out_counts[b] <= k_fetch entries of: all live rows with cosine >= min_score (fp64,
 * inclusive; pass -INFINITY for "no threshold"), ordered by score descending, ties by
 * ascending slot.  Scores are the reference's fp64 cosine, bit for bit, for the stored
 * (bf16-exact) rows.  out_slots are GLOBAL (slot_base + local).  Unused tail entries of
 * row b are slot -1 / score NaN.  kernel_ms_out (nullable): device time of the call.
 */
rbk_status rbk_index_search_f64(rbk_index* idx, const double* queries, int32_t B, int32_t query_dim,
                                int32_t k_fetch, double min_score, int64_t* out_slots, double* out_scores,
                                int32_t* out_counts, float* kernel_ms_out);
rbk_status rbk_index_search_f32(rbk_index* idx, const float* queries, int32_t B, int32_t query_dim,
                                int32_t k_fetch, double min_score, int64_t* out_slots, double* out_scores,
                                int32_t* out_counts, float* kernel_ms_out);
/* Device-resident variant: queries (f32, B x dim) and all outputs are device pointers on
 * the index's GPU; work is enqueued on the index stream and the call returns after the
 * exactness check of the batch (it synchronises the stream once). */
rbk_status rbk_index_search_device(rbk_index* idx, const void* dev_queries_f32, int32_t B, int32_t k_fetch,
                                   double min_score, void* dev_out_slots_i64, void* dev_out_scores_f64,
                                   void* dev_out_counts_i32);

/* More hits than RBK_MAX_K_FETCH (callers of the reference pass limit: 1000, knowledge-context.ts:150): the exact
 * fp64 cosine of EVERY row, out_scores[b * size() + slot], NaN for tombstoned / zero rows (which the reference's
 * `>= minScore` drops too, S3).  The host applies the threshold, the stable sort and the cut literally
 * (vector-store.ts:212-221).  One fp64 pass over the corpus per query: the large-k path, not the hot path. */
rbk_status rbk_index_exact_scores_f64(rbk_index* idx, const double* queries, int32_t B, int32_t query_dim,
                                      double* out_scores);

/* Enqueue-only variant: nothing is synchronised, the call returns as soon as the kernels are queued on the index
 * stream, so batches pipeline back to back and an exchange step (all-gather + rbk_merge_topk_packed_device) can be
 * queued behind it without a host round trip in between.  dev_out_flags_i32[B]: 0 = the answer of query b is
 * proven exact (the normal case), 1 = not proven (more near-ties around the k_fetch-th hit than the candidate
 * margin holds).  The caller checks the flags when it eventually synchronises - for a sharded corpus AFTER the
 * merge, whose out_flags OR the shards' flags - and re-answers a batch that has any dirty query with the
 * synchronous rbk_index_search_device (which rescans with a wide margin / falls back to the exhaustive kernel).
 * All four outputs may point into a packed block (below). */
rbk_status rbk_index_search_device_async(rbk_index* idx, const void* dev_queries_f32, int32_t B, int32_t k_fetch,
                                         double min_score, void* dev_out_slots_i64, void* dev_out_scores_f64,
                                         void* dev_out_counts_i32, void* dev_out_flags_i32);

/* Merge G per-shard result lists (layout [G][B][k_fetch], each sorted as above, device
 * memory, e.g. the output of an all-gather) into [B][k_fetch] by (score desc, slot asc).
 * Enqueued on cuda_stream; no synchronisation. */
rbk_status rbk_merge_topk_device(int32_t device, void* cuda_stream, int32_t G, int32_t B, int32_t k_fetch,
                                 const void* dev_slots_i64, const void* dev_scores_f64, const void* dev_counts_i32,
                                 void* dev_out_slots_i64, void* dev_out_scores_f64, void* dev_out_counts_i32);

/* Same merge for G PACKED per-shard blocks laid end to end (what ONE all-gather of each rank's block
 * produces).  Block layout, rbk_packed_block_bytes(B, k_fetch) bytes: slots i64[B*k_fetch] | scores
 * f64[B*k_fetch] | counts i32[B] (padded to 16 bytes) | flags i32[B] (padded to 16 bytes; starts at
 * rbk_packed_flags_offset).  rbk_index_search_device(_async) can write straight into a block: pass block,
 * block + B*k_fetch*8, block + B*k_fetch*16 (and block + rbk_packed_flags_offset for the flags).
 * dev_out_flags_i32 (nullable): i32[B+1]; [b] = OR of the shards' flags of query b, [B] += number of dirty
 * queries of this call (a running count the caller zeroes, so a pipelined loop checks once at the end). */
int64_t rbk_packed_block_bytes(int32_t B, int32_t k_fetch);
int64_t rbk_packed_flags_offset(int32_t B, int32_t k_fetch);
rbk_status rbk_merge_topk_packed_device(int32_t device, void* cuda_stream, int32_t G, int32_t B, int32_t k_fetch,
                                        const void* dev_blocks, void* dev_out_slots_i64, void* dev_out_scores_f64,
                                        void* dev_out_counts_i32, void* dev_out_flags_i32);

/* ---- one corpus over several GPUs behind ONE handle (rbk_group.cu) ----
 * What a single host process (RunbookAI is one Node process) uses to shard `this.embeddings` over the GPUs of a
 * box: one index per device, rows dealt out block-cyclically (4096-row blocks, so the corpus may grow at sync
 * time), every search = H2D of the queries to every GPU, the fused scan on every GPU, ONE ncclAllGather of the packed
 * per-GPU blocks over NVLink, a merge kernel on device_ids[0], one D2H and ONE host synchronisation - all inside
 * rbk_group_search_*.  Slots are GLOBAL insertion indices, exactly as for a single index; results are identical
 * to a single index holding the same rows (ids and fp64 scores bit for bit, ties by ascending slot).
 * NCCL is looked up at run time (dlopen "libnccl.so.2"); a group of more than one GPU fails with RBK_ENCCL if it
 * is missing, a one-GPU group never touches it.  Mutation and search semantics, limits (k_fetch <= RBK_MAX_K_FETCH)
 * and error conventions are those of the rbk_index_* call of the same name. */
typedef struct rbk_group rbk_group;
rbk_status rbk_group_create(int32_t dim, const int32_t* device_ids, int32_t n_devices, int64_t capacity_hint,
                            uint32_t flags /* RBK_INDEX_KEEP_F64 */, rbk_group** out);
void rbk_group_destroy(rbk_group* grp); /* NULL is a no-op */
rbk_status rbk_group_append_f64(rbk_group* grp, const double* rows, int64_t n_rows, int64_t* first_slot_out);
rbk_status rbk_group_append_f32(rbk_group* grp, const float* rows, int64_t n_rows, int64_t* first_slot_out);
rbk_status rbk_group_append_bf16(rbk_group* grp, const uint16_t* rows, int64_t n_rows, int64_t* first_slot_out);
rbk_status rbk_group_overwrite_f64_batch(rbk_group* grp, const int64_t* slots, int64_t n, const double* rows);
rbk_status rbk_group_tombstone(rbk_group* grp, const int64_t* slots, int64_t n);
rbk_status rbk_group_clear(rbk_group* grp);
int64_t rbk_group_count(const rbk_group* grp); /* live rows */
int64_t rbk_group_size(const rbk_group* grp);  /* slots used, tombstones included */
int32_t rbk_group_devices(const rbk_group* grp);
rbk_index* rbk_group_member(rbk_group* grp, int32_t i); /* the i-th device's index (stats, tests); owned by the group */
int64_t rbk_group_redone_batches(const rbk_group* grp); /* batches re-answered because a shard's proof failed */
rbk_status rbk_group_search_f32(rbk_group* grp, const float* queries, int32_t B, int32_t query_dim, int32_t k_fetch,
                                double min_score, int64_t* out_slots, double* out_scores, int32_t* out_counts,
                                float* device_ms_out);
rbk_status rbk_group_search_f64(rbk_group* grp, const double* queries, int32_t B, int32_t query_dim, int32_t k_fetch,
                                double min_score, int64_t* out_slots, double* out_scores, int32_t* out_counts,
                                float* device_ms_out);

/* ---- introspection ---- */
typedef struct {
  int64_t searches;         /* search calls */
  int64_t queries;          /* queries answered */
  int64_t fallback_queries; /* queries re-answered by the exhaustive fp64 kernel (after the wide retry) */
  int64_t scan_launches;    /* launches of the fused scan kernel */
  int64_t kernel_launches;  /* all kernel launches made by this index */
  float last_scan_ms;       /* device time of the scan kernel(s) of the last synchronous search (async: of the last finished scan) */
  float last_total_ms;      /* device time of the whole last search */
  int32_t last_kprime;      /* candidates kept per query by the last scan */
  int32_t sm_count;
  int32_t last_ring_stages; /* smem ring depth of the last scan kernel (pair kernel: 7, or 6 if the smem base is unaligned) */
  int32_t retry_batches;    /* batches scanned a second time with the widest candidate margin after a failed proof */
  double scan_ms_total;     /* device time of all scan kernels that have FINISHED so far (CUDA events around every launch) */
  int64_t scans_timed;      /* number of scan launches folded into scan_ms_total */
  int64_t graph_replays;    /* small-batch host searches served by replaying the captured CUDA graph (prep + scan + finalize + copies) */
} rbk_stats;
/* Never blocks: folds in the scans that have finished and returns. */
rbk_status rbk_index_stats(rbk_index* idx, rbk_stats* out);

/* Debug/validation aid (tests only): run the scan on `queries` (host f32, B x dim) and
 * return the approximate scores of every (query,row) pair, B x size() floats (host).
 * NaN marks tombstoned/zero rows. */
rbk_status rbk_index_debug_scores_f32(rbk_index* idx, const float* queries, int32_t B, float* out_scores);

#ifdef __cplusplus
}
#endif
#endif /* RBK_KNN_H */
