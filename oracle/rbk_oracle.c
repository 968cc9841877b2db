/*
 * rbk_oracle.c — CPU restatement of RunbookAI's knowledge-base vector search.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (runbookai_b200/) never links, imports or calls anything in this directory.
 *
 * PARITY UNPINNED: the reference (TypeScript, Node >= 20) cannot run in the build
 * container and ships no test, golden vector or fixture for this path (SURVEY.md §4,
 * §8c).  This file therefore restates the algorithm line by line from the sources
 * cited below; it is cross-checked against an independent pure-Python restatement
 * (oracle/pyref.py) but not against outputs of the reference itself.
 *
 * Build flags are part of the contract: -O2 -ffp-contract=off, no
 * -ffast-math, no vectorised reassociation, no BLAS.  JS Numbers are IEEE-754 binary64
 * with correctly rounded + * / sqrt and no fused multiply-add, which is what C double
 * arithmetic gives under these flags.
 *
 * Reference sources restated (paths relative to the reference checkout):
 *   cosineSimilarity      src/knowledge/indexer/embedder.ts:168-184
 *   findMostSimilar       src/knowledge/indexer/embedder.ts:189-202
 *   VectorStore.search    src/knowledge/store/vector-store.ts:188-221 (scan, sort, 2*topK cut)
 *   reciprocalRankFusion  src/knowledge/retriever/hybrid-search.ts:106-151
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RBK_ORACLE_EDIM (-2)

/* bf16 -> binary64 is exact: bf16 is the top 16 bits of a binary32. */
static inline double bf16_to_f64(uint16_t h) {
  uint32_t u = ((uint32_t)h) << 16;
  float f;
  memcpy(&f, &u, sizeof f);
  return (double)f;
}

/* embedder.ts:168-184.  Three separate accumulators, index order, multiply then add,
 * sqrt twice then multiply, one divide.  Zero vector -> 0/0 = NaN (S3). */
double rbk_oracle_cosine(const double *a, const double *b, int64_t d) {
  double dot = 0.0, na = 0.0, nb = 0.0;
  for (int64_t i = 0; i < d; i++) {
    dot += a[i] * b[i];
    na += a[i] * a[i];
    nb += b[i] * b[i];
  }
  return dot / (sqrt(na) * sqrt(nb));
}

static double cosine_row_bf16(const double *q, const uint16_t *row, int64_t d) {
  double dot = 0.0, na = 0.0, nb = 0.0;
  for (int64_t i = 0; i < d; i++) {
    double c = bf16_to_f64(row[i]);
    dot += q[i] * c;
    na += q[i] * q[i];
    nb += c * c;
  }
  return dot / (sqrt(na) * sqrt(nb));
}

/* All scores of one query (used by tests that need the full score vector). */
void rbk_oracle_scores_f64(const double *corpus, int64_t n, int64_t d, const double *q, double *out) {
  for (int64_t r = 0; r < n; r++) out[r] = rbk_oracle_cosine(q, corpus + r * d, d);
}
void rbk_oracle_scores_bf16(const uint16_t *corpus, int64_t n, int64_t d, const double *q, double *out) {
  for (int64_t r = 0; r < n; r++) out[r] = cosine_row_bf16(q, corpus + r * d, d);
}

typedef struct {
  double score;
  int64_t slot;
} hit_t;

/* Stable merge sort, descending by score: the observable behaviour of
 * `scored.sort((a, b) => b.score - a.score)` (vector-store.ts:218) under V8's stable
 * TimSort.  NaN never reaches it when a threshold is active (S3). */
static void merge_sort_desc(hit_t *a, hit_t *tmp, int64_t n) {
  if (n < 2) return;
  int64_t h = n / 2;
  merge_sort_desc(a, tmp, h);
  merge_sort_desc(a + h, tmp, n - h);
  int64_t i = 0, j = h, k = 0;
  while (i < h && j < n) {
    /* take from the right run only when strictly greater: ties keep insertion order */
    if (a[j].score > a[i].score) tmp[k++] = a[j++];
    else tmp[k++] = a[i++];
  }
  while (i < h) tmp[k++] = a[i++];
  while (j < n) tmp[k++] = a[j++];
  memcpy(a, tmp, (size_t)n * sizeof(hit_t));
}

/*
 * Literal restatement of vector-store.ts:207-221: scan every live row in insertion
 * (slot) order, keep score >= min_score (S5; NaN fails the compare), stable sort
 * descending (S6), return the first k_fetch (= 2*topK in the reference, S7).
 * live == NULL means every row is live.  use_threshold == 0 gives findMostSimilar
 * (embedder.ts:189-202: no threshold; NaN rows are dropped).
 * Returns the number of hits written, or RBK_ORACLE_EDIM.
 */
static int64_t search_fullsort(const void *corpus, int is_bf16, int64_t n, int64_t d, const double *q,
                               int64_t qd, const uint8_t *live, int use_threshold, double min_score,
                               int64_t k_fetch, int64_t *out_slots, double *out_scores) {
  if (qd != d) return RBK_ORACLE_EDIM; /* embedder.ts:169-171 */
  hit_t *scored = (hit_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(hit_t));
  hit_t *tmp = (hit_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(hit_t));
  int64_t m = 0;
  for (int64_t r = 0; r < n; r++) {
    if (live && !live[r]) continue;
    double s = is_bf16 ? cosine_row_bf16(q, (const uint16_t *)corpus + r * d, d)
                       : rbk_oracle_cosine(q, (const double *)corpus + r * d, d);
    if (use_threshold ? (s >= min_score) : (s == s)) {
      scored[m].score = s;
      scored[m].slot = r;
      m++;
    }
  }
  merge_sort_desc(scored, tmp, m);
  int64_t cnt = m < k_fetch ? m : k_fetch;
  for (int64_t i = 0; i < cnt; i++) {
    out_slots[i] = scored[i].slot;
    out_scores[i] = scored[i].score;
  }
  free(scored);
  free(tmp);
  return cnt;
}

int64_t rbk_oracle_search_f64(const double *corpus, int64_t n, int64_t d, const double *q, int64_t qd,
                              const uint8_t *live, int use_threshold, double min_score, int64_t k_fetch,
                              int64_t *out_slots, double *out_scores) {
  return search_fullsort(corpus, 0, n, d, q, qd, live, use_threshold, min_score, k_fetch, out_slots, out_scores);
}
int64_t rbk_oracle_search_bf16(const uint16_t *corpus, int64_t n, int64_t d, const double *q, int64_t qd,
                               const uint8_t *live, int use_threshold, double min_score, int64_t k_fetch,
                               int64_t *out_slots, double *out_scores) {
  return search_fullsort(corpus, 1, n, d, q, qd, live, use_threshold, min_score, k_fetch, out_slots, out_scores);
}

/* ---- "ref-allcores": same per-(query,row) arithmetic, rows split over threads, each
 * thread keeps its best k_fetch in (score desc, slot asc) order, then a merge.  The
 * result equals stable-sort-then-slice because (score desc, slot asc) is a total order
 * on hits and the first k_fetch of a stable descending sort are exactly its k_fetch
 * smallest elements.  Used as the fair same-box CPU baseline. ---- */
static inline int hit_before(const hit_t *a, const hit_t *b) {
  return a->score > b->score || (a->score == b->score && a->slot < b->slot);
}

static int64_t topk_insert(hit_t *best, int64_t cnt, int64_t k, hit_t h) {
  if (cnt == k) {
    if (!hit_before(&h, &best[k - 1])) return cnt;
    cnt--;
  }
  int64_t i = cnt;
  while (i > 0 && hit_before(&h, &best[i - 1])) {
    best[i] = best[i - 1];
    i--;
  }
  best[i] = h;
  return cnt + 1;
}

typedef struct {
  const uint16_t *corpus;
  int64_t r0, r1, d, slot_base;
  const double *queries;
  int64_t nq;
  const uint8_t *live;
  int use_threshold;
  double min_score;
  int64_t k;
  hit_t *best;  /* nq * k */
  int64_t *cnt; /* nq */
} mt_job_t;

static void *mt_worker(void *arg) {
  mt_job_t *j = (mt_job_t *)arg;
  for (int64_t b = 0; b < j->nq; b++) j->cnt[b] = 0;
  for (int64_t r = j->r0; r < j->r1; r++) {
    if (j->live && !j->live[r]) continue;
    const uint16_t *row = j->corpus + r * j->d;
    for (int64_t b = 0; b < j->nq; b++) {
      double s = cosine_row_bf16(j->queries + b * j->d, row, j->d);
      if (j->use_threshold ? (s >= j->min_score) : (s == s)) {
        hit_t h = {s, r + j->slot_base};
        j->cnt[b] = topk_insert(j->best + b * j->k, j->cnt[b], j->k, h);
      }
    }
  }
  return NULL;
}

/* Batched, multi-threaded search over a bf16 corpus.  out_* are [nq][k_fetch]. */
int64_t rbk_oracle_search_batch_bf16_mt(const uint16_t *corpus, int64_t n, int64_t d, const double *queries,
                                        int64_t nq, const uint8_t *live, int use_threshold, double min_score,
                                        int64_t k_fetch, int n_threads, int64_t *out_slots, double *out_scores,
                                        int32_t *out_counts) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n && n > 0) n_threads = (int)n;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
  mt_job_t *jobs = (mt_job_t *)malloc(sizeof(mt_job_t) * (size_t)n_threads);
  int64_t per = (n + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; t++) {
    mt_job_t *j = &jobs[t];
    j->corpus = corpus;
    j->r0 = t * per < n ? t * per : n;
    j->r1 = (t + 1) * per < n ? (t + 1) * per : n;
    j->d = d;
    j->slot_base = 0;
    j->queries = queries;
    j->nq = nq;
    j->live = live;
    j->use_threshold = use_threshold;
    j->min_score = min_score;
    j->k = k_fetch;
    j->best = (hit_t *)malloc(sizeof(hit_t) * (size_t)(nq * k_fetch > 0 ? nq * k_fetch : 1));
    j->cnt = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nq > 0 ? nq : 1));
    pthread_create(&th[t], NULL, mt_worker, j);
  }
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  for (int64_t b = 0; b < nq; b++) {
    hit_t *best = (hit_t *)malloc(sizeof(hit_t) * (size_t)(k_fetch > 0 ? k_fetch : 1));
    int64_t cnt = 0;
    for (int t = 0; t < n_threads; t++)
      for (int64_t i = 0; i < jobs[t].cnt[b]; i++) cnt = topk_insert(best, cnt, k_fetch, jobs[t].best[b * k_fetch + i]);
    for (int64_t i = 0; i < cnt; i++) {
      out_slots[b * k_fetch + i] = best[i].slot;
      out_scores[b * k_fetch + i] = best[i].score;
    }
    out_counts[b] = (int32_t)cnt;
    free(best);
  }
  for (int t = 0; t < n_threads; t++) {
    free(jobs[t].best);
    free(jobs[t].cnt);
  }
  free(jobs);
  free(th);
  return 0;
}

/* ---- "verify" variant of the batched search: the SAME per-(query,row) arithmetic, arranged so a CPU can
 * check millions of rows in seconds.  Two rearrangements, neither of which changes a single rounding:
 *   1. normA (query) and normB (row) are each one deterministic sequential chain (embedder.ts:179-180) whose
 *      value does not depend on the other vector, so they are computed once per query / once per row instead
 *      of once per pair;
 *   2. the dot chains of 8 different queries against one row are independent of each other, so they are
 *      advanced side by side (GCC vector extension, 2 lanes x 4 registers): every lane still does
 *      dot = dot + q[i]*c[i] for i = 0..d-1, multiply then add, no contraction (-ffp-contract=off).
 * tests/test_oracle.py checks it bit for bit against the literal loop above.  It is used only as the CHECKER
 * of large GPU runs (tests, bench.py's parity leg) - never as the timed CPU baseline, which stays literal. */
typedef double v2d __attribute__((vector_size(16)));

typedef struct {
  const uint16_t *corpus;
  int64_t r0, r1, d, slot_base;
  const double *qt;   /* [nq8/8][d][8] transposed query blocks, zero padded */
  const double *qn;   /* [nq] sqrt(normA) */
  int64_t nq;
  const uint8_t *live;
  int use_threshold;
  double min_score;
  int64_t k;
  hit_t *best;
  int64_t *cnt;
} vf_job_t;

static void *vf_worker(void *arg) {
  vf_job_t *j = (vf_job_t *)arg;
  const int64_t d = j->d;
  double *c = (double *)malloc(sizeof(double) * (size_t)(d > 0 ? d : 1));
  for (int64_t b = 0; b < j->nq; b++) j->cnt[b] = 0;
  for (int64_t r = j->r0; r < j->r1; r++) {
    if (j->live && !j->live[r]) continue;
    const uint16_t *row = j->corpus + r * d;
    double nb = 0.0;
    for (int64_t i = 0; i < d; i++) {
      c[i] = bf16_to_f64(row[i]);
      nb += c[i] * c[i];
    }
    const double sb = sqrt(nb);
    for (int64_t b0 = 0; b0 < j->nq; b0 += 8) {
      const double *qt = j->qt + (b0 / 8) * d * 8;
      v2d a0 = {0.0, 0.0}, a1 = a0, a2 = a0, a3 = a0;
      for (int64_t i = 0; i < d; i++) {
        const v2d cv = {c[i], c[i]};
        const v2d *q4 = (const v2d *)(qt + i * 8);
        a0 = a0 + q4[0] * cv;
        a1 = a1 + q4[1] * cv;
        a2 = a2 + q4[2] * cv;
        a3 = a3 + q4[3] * cv;
      }
      const double dots[8] = {a0[0], a0[1], a1[0], a1[1], a2[0], a2[1], a3[0], a3[1]};
      for (int64_t e = 0; e < 8 && b0 + e < j->nq; e++) {
        const int64_t b = b0 + e;
        const double s = dots[e] / (j->qn[b] * sb);   /* embedder.ts:183 */
        if (j->use_threshold ? (s >= j->min_score) : (s == s)) {
          hit_t h = {s, r + j->slot_base};
          j->cnt[b] = topk_insert(j->best + b * j->k, j->cnt[b], j->k, h);
        }
      }
    }
  }
  free(c);
  return NULL;
}

/* Same contract as rbk_oracle_search_batch_bf16_mt; slot_base is added to every returned slot. */
int64_t rbk_oracle_search_batch_bf16_verify(const uint16_t *corpus, int64_t n, int64_t d, const double *queries,
                                            int64_t nq, const uint8_t *live, int use_threshold, double min_score,
                                            int64_t k_fetch, int n_threads, int64_t slot_base, int64_t *out_slots,
                                            double *out_scores, int32_t *out_counts) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n && n > 0) n_threads = (int)n;
  const int64_t nblk = (nq + 7) / 8;
  double *qt = (double *)aligned_alloc(64, sizeof(double) * (size_t)((nblk * d * 8 > 0 ? nblk * d * 8 : 8)));
  double *qn = (double *)malloc(sizeof(double) * (size_t)(nq > 0 ? nq : 1));
  memset(qt, 0, sizeof(double) * (size_t)(nblk * d * 8));
  for (int64_t b = 0; b < nq; b++) {
    double na = 0.0;
    for (int64_t i = 0; i < d; i++) {
      const double x = queries[b * d + i];
      na += x * x;
      qt[(b / 8) * d * 8 + i * 8 + (b % 8)] = x;
    }
    qn[b] = sqrt(na);
  }
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
  vf_job_t *jobs = (vf_job_t *)malloc(sizeof(vf_job_t) * (size_t)n_threads);
  int64_t per = (n + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; t++) {
    vf_job_t *j = &jobs[t];
    j->corpus = corpus;
    j->r0 = t * per < n ? t * per : n;
    j->r1 = (t + 1) * per < n ? (t + 1) * per : n;
    j->d = d;
    j->slot_base = slot_base;
    j->qt = qt;
    j->qn = qn;
    j->nq = nq;
    j->live = live;
    j->use_threshold = use_threshold;
    j->min_score = min_score;
    j->k = k_fetch;
    j->best = (hit_t *)malloc(sizeof(hit_t) * (size_t)(nq * k_fetch > 0 ? nq * k_fetch : 1));
    j->cnt = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nq > 0 ? nq : 1));
    pthread_create(&th[t], NULL, vf_worker, j);
  }
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  hit_t *best = (hit_t *)malloc(sizeof(hit_t) * (size_t)(k_fetch > 0 ? k_fetch : 1));
  for (int64_t b = 0; b < nq; b++) {
    int64_t cnt = 0;
    for (int t = 0; t < n_threads; t++)
      for (int64_t i = 0; i < jobs[t].cnt[b]; i++) cnt = topk_insert(best, cnt, k_fetch, jobs[t].best[b * k_fetch + i]);
    for (int64_t i = 0; i < cnt; i++) {
      out_slots[b * k_fetch + i] = best[i].slot;
      out_scores[b * k_fetch + i] = best[i].score;
    }
    out_counts[b] = (int32_t)cnt;
  }
  free(best);
  for (int t = 0; t < n_threads; t++) {
    free(jobs[t].best);
    free(jobs[t].cnt);
  }
  free(jobs);
  free(th);
  free(qt);
  free(qn);
  return 0;
}

/*
 * hybrid-search.ts:106-151 (S12).  Inputs are the chunk ids of the two ranked lists as
 * small integers (the caller interns the id strings).  score(id) = sum w/(k + i + 1),
 * FTS list first, then vector list; Map insertion order = first appearance; stable
 * descending sort; first top_k.  Returns the number of fused entries written.
 */
int64_t rbk_oracle_rrf(const int64_t *fts_ids, int64_t n_fts, const int64_t *vec_ids, int64_t n_vec, double rrf_k,
                       double fts_w, double vec_w, int64_t top_k, int64_t *out_ids, double *out_scores) {
  int64_t cap = n_fts + n_vec;
  hit_t *ent = (hit_t *)malloc(sizeof(hit_t) * (size_t)(cap > 0 ? cap : 1));
  hit_t *tmp = (hit_t *)malloc(sizeof(hit_t) * (size_t)(cap > 0 ? cap : 1));
  int64_t m = 0;
  for (int pass = 0; pass < 2; pass++) {
    const int64_t *ids = pass == 0 ? fts_ids : vec_ids;
    int64_t cnt = pass == 0 ? n_fts : n_vec;
    double w = pass == 0 ? fts_w : vec_w;
    for (int64_t i = 0; i < cnt; i++) {
      double rrf = w * (1.0 / (rrf_k + (double)i + 1.0));
      int64_t e = 0;
      for (; e < m; e++)
        if (ent[e].slot == ids[i]) break;
      if (e < m) ent[e].score += rrf;
      else {
        ent[m].slot = ids[i];
        ent[m].score = rrf;
        m++;
      }
    }
  }
  merge_sort_desc(ent, tmp, m);
  int64_t cnt = m < top_k ? m : top_k;
  for (int64_t i = 0; i < cnt; i++) {
    out_ids[i] = ent[i].slot;
    out_scores[i] = ent[i].score;
  }
  free(ent);
  free(tmp);
  return cnt;
}

int rbk_oracle_abi(void) { return 1; }
