/* liblrag -- C ABI of the B200-native hybrid-retrieval engine (sm_100a only).
 *
 * The reference (Fan-Luo/Legal-RAG) is pure Python and has no FFI of its own: its hot
 * path bottoms out in three third-party objects.  Each entry point below replaces one
 * of those call sites (cited per function; paths are relative to the reference root).
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain C symbols, no torch / C++ types in any signature;
 *   - every tensor argument is a DEVICE pointer (torch.Tensor.data_ptr()); row-major;
 *   - work is enqueued on the caller's stream and the call returns without
 *     synchronising;
 *   - no allocation inside: the caller passes a workspace of at least
 *     <fn>_workspace_bytes(...) bytes, 256-byte aligned;
 *   - return 0 on success, a negative LRAG_E* code otherwise; lrag_last_error() gives
 *     the thread-local message;
 *   - one process per GPU.  After lrag_init the scoring entry points keep no state of their
 *     own between calls and may be called concurrently from any host thread on distinct
 *     streams with distinct workspaces.  Process-wide state exists in three places and is
 *     documented where it is declared: the device binding of lrag_init, the launch profiler
 *     (lrag_prof_*, benchmarks only, not thread-safe) and the BM25 item-size knob
 *     (lrag_bm25_set_item_slabs, set once before serving).  Kernel attributes are cached per
 *     device, so a second device bound in the same process launches correctly;
 *   - no CPU fallback: lrag_init fails on anything that is not compute capability 10.x.
 *
 * Result ordering everywhere: (score descending, id ascending); rows shorter than k are
 * padded with (LRAG_PAD_SCORE, -1).
 */
#ifndef LRAG_H_
#define LRAG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LRAG_VERSION 100

#define LRAG_OK 0
#define LRAG_EINVAL (-1)   /* bad argument                                   */
#define LRAG_ENOSPC (-2)   /* workspace too small                            */
#define LRAG_ECUDA (-3)    /* CUDA runtime / driver error                    */
#define LRAG_EARCH (-4)    /* device is not sm_100 (no fallback path exists) */

#define LRAG_PAD_SCORE (-3.4028234663852886e38f) /* faiss pads with lowest float */
#define LRAG_MAX_K 1024
#define LRAG_BM25_MAX_QUERY_TERMS 128 /* tokens per query, repeats included */

/* opaque stream handle: a cudaStream_t (torch.cuda.current_stream().cuda_stream) */
typedef void* lrag_stream_t;

int lrag_version(void);
/* Binds to `device`, checks the architecture, caches SM count and the driver's
 * tensor-map encoder.  Must be called once per device before any other call. */
int lrag_init(int device);
const char* lrag_last_error(void);
/* number of SMs the persistent kernels size their grids with (148 on B200) */
int lrag_sm_count(void);

/* Launch profiler (bench.py's roofline leg): after lrag_prof_enable(capacity) every launch of a
 * dominant kernel is bracketed by a CUDA event pair on the launching stream.  lrag_prof_collect
 * synchronises on them, writes up to max_n durations (ms) and kernel tags (0 dense scan, 1 BM25
 * scan, 2 MaxSim, 3 fusion) and returns how many it wrote (negative = error); the ring restarts.
 * lrag_prof_enable(0) switches it off.  Single-threaded use only. */
int lrag_prof_enable(int capacity);
int lrag_prof_collect(float* ms, int* tag, int max_n);
/* cumulative number of kernels this library has launched in the process (bench.py's gpu_launches) */
long long lrag_launch_count(void);

/* ---------------------------------------------------------------------------------
 * Dense channel.  Replaces faiss `index.search(q_vec, k)` at
 * legalrag/retrieval/dense_retriever.py:42 and legalrag/retrieval/vector_store.py:169
 * (IndexFlat, METRIC_INNER_PRODUCT): D, I = topk_k(Q . X^T).
 *   X [N, d] bf16, Q [nq, d] bf16, d % 8 == 0 (16-byte rows for TMA), 1 <= k <= LRAG_MAX_K.
 *   out_id = id_base + row index in X  (id_base: first global doc id of this shard).
 * One fused pass: tcgen05 bf16 contraction with TMA-staged corpus tiles, per-query
 * running-threshold candidate filter in the TMEM epilogue, then an exact select.
 * The [nq, N] score matrix is never written. */
size_t lrag_dense_topk_workspace_bytes(int64_t N, int d, int nq, int k);
int lrag_dense_topk_bf16(const void* X, int64_t N, int d, const void* Q, int nq, int k,
                         int64_t id_base, float* out_score, int64_t* out_id, void* ws,
                         size_t ws_bytes, lrag_stream_t stream);
/* The same scan confined to `max_ctas` SMs (0 = all of them), for a stage that shares the machine with another
 * persistent kernel (see lrag_sm_reserve).  Results do not depend on max_ctas; the workspace does. */
size_t lrag_dense_topk_workspace_bytes_part(int64_t N, int d, int nq, int k, int max_ctas);
int lrag_dense_topk_bf16_part(const void* X, int64_t N, int d, const void* Q, int nq, int k,
                              int64_t id_base, int max_ctas, float* out_score, int64_t* out_id, void* ws,
                              size_t ws_bytes, lrag_stream_t stream);
/* CUDA-core cross-check of the same contract (fp32 FMA, materialises [nq, N] scores in
 * the workspace).  Test instrument for the tcgen05 path; small shapes only. */
size_t lrag_dense_topk_ref_workspace_bytes(int64_t N, int d, int nq, int k);
int lrag_dense_topk_bf16_ref(const void* X, int64_t N, int d, const void* Q, int nq, int k,
                             int64_t id_base, float* out_score, int64_t* out_id, void* ws,
                             size_t ws_bytes, lrag_stream_t stream);

/* ---------------------------------------------------------------------------------
 * Exact row-wise top-k of a materialised fp32 score matrix S [nq, ld] (first N columns
 * of each row are live).  Replaces the `sorted(range(N), key=scores[i], reverse=True)[:k]`
 * of legalrag/retrieval/bm25_retriever.py:75 for small corpora and ranks MaxSim
 * candidate scores.  If `col_id` is non-NULL ([nq, N] int64) the returned id is
 * col_id[row, col] (entries with col_id < 0 are skipped; ids of any 64-bit width: they never
 * pass through the 32-bit tie field of the selection keys, equal scores are ordered by id
 * when the hits are written, and which of several EQUAL scores make the cut at rank k is
 * decided by column); else id_base + col.
 * With a workspace of lrag_topk_select_workspace_bytes() long rows are streamed once by a
 * one-wave grid (per-CTA candidate lists in the workspace, then a merge); without one
 * (ws NULL) every row is selected by a single CTA. */
size_t lrag_topk_select_workspace_bytes(int nq, int64_t N, int k);
int lrag_topk_select_f32(const float* S, int64_t ld, int nq, int64_t N, int k, int64_t id_base,
                         const int64_t* col_id, float* out_score, int64_t* out_id, void* ws,
                         size_t ws_bytes, lrag_stream_t stream);

/* k-way merge of per-shard candidate lists (the step after the NCCL all-gather; no
 * reference counterpart -- the reference is single-process).  score/id [nq, L];
 * id < 0 entries are padding.  Ids are full 64-bit values: a row whose ids span less than
 * 2^32 - 1 (any sharded corpus below 4 G documents) is ordered exactly (score desc, id asc);
 * a wider row is ordered the same way, except that which of several EQUAL scores make the
 * cut at rank k is decided by position in the input row.
 * lrag_topk_merge_shards reads the all-gathered buffers in place: `shards` blocks of
 * [nq, kin], shard g's block starting score_shard_stride / id_shard_stride ELEMENTS after
 * shard g - 1's (a packed all-gather of several channels passes each channel's offset into
 * the rank-0 block and the per-rank buffer length). */
int lrag_topk_merge(const float* score, const int64_t* id, int nq, int L, int k,
                    float* out_score, int64_t* out_id, lrag_stream_t stream);
int lrag_topk_merge_shards(const float* score, const int64_t* id, int64_t score_shard_stride,
                           int64_t id_shard_stride, int shards, int nq, int kin, int k,
                           float* out_score, int64_t* out_id, lrag_stream_t stream);

/* Gathered inner products: out_score[q, c] = <Q[q], X[rows[q, c]]> (rows outside [0, N) give -inf).
 * Replaces the re-embedding + cosine loop of the graph-expansion channel
 * (legalrag/retrieval/graph_retriever.py:177-186, `_cosine_sim` :20-22): neighbour chunks are index rows,
 * their unit-norm embeddings are already in the corpus.  rows [nq, C] int64, local to the shard. */
int lrag_dense_gather_scores_bf16(const void* X, int64_t N, int d, const void* Q, int nq,
                                  const int64_t* rows, int C, float* out_score, lrag_stream_t stream);

/* ---------------------------------------------------------------------------------
 * BM25 channel.  Replaces `bm25.get_scores(tokens)` + the full Python sort at
 * legalrag/retrieval/bm25_retriever.py:74-75 (rank_bm25.BM25Okapi).
 * Index = term-major CSR with doc ids ascending inside each term:
 *   indptr [V+1] int64, doc_id [nnz] int32 (LOCAL row in this shard), impact [nnz] fp32 (4-byte aligned; 128-byte
 *   alignment of both arrays keeps every warp-wide load on one line)
 *   = idf[t] * tf*(k1+1) / (tf + k1*(1-b+b*dl/avgdl))  with GLOBAL idf/avgdl.
 * Queries = CSR of term ids: q_indptr [nq+1] int64, q_term [*] int32 (repeats allowed and
 * scored once per occurrence, -1 / out-of-range = OOV).
 * Every document is a candidate: documents matching no term score 0 and are returned,
 * lowest id first, when fewer than k documents match (the reference's behaviour).
 * `max_query_terms` = the longest query's token count (host-side knowledge of q_indptr; at most
 * LRAG_BM25_MAX_QUERY_TERMS).  `nonneg` != 0 asserts every impact >= 0 (true whenever the corpus'
 * average idf is positive): doc slabs no query term touches are then skipped and zero-score
 * documents are filled in by id; with nonneg == 0 every document of every slab is ranked.
 * `impact_bound` >= max |impact[p]| over the index (compute it once when the index is built):
 * scores are accumulated in int32 fixed point with a per-query power-of-two scale chosen so that
 * impact_bound * (tokens of the query) stays below 2^30 -- at least 2^-22 absolute resolution for
 * an 8-token query with impacts up to 30 -- which makes the sums exact, order-independent and
 * reproducible; a bound that is too small lets a score overflow. */
size_t lrag_bm25_topk_workspace_bytes(int64_t N, int nq, int k, int64_t max_query_terms);
/* Tuning knob (process-wide; set it once, before sizing workspaces and not concurrently with launches): slabs per
 * work item (a slab is the doc range one CTA accumulates in shared memory: 22 528 docs for k <= 256).  Small items
 * keep the posting ranges all queries are working on inside L2; large items amortise the per-item state hand-off.
 * 0 restores the default (about 393 k docs per item, or the LRAG_BM25_ITEM_SLABS environment variable). */
int lrag_bm25_set_item_slabs(int slabs);
int lrag_bm25_topk(const int64_t* indptr, const int32_t* doc_id, const float* impact, int64_t V,
                   int64_t nnz, const int64_t* q_indptr, const int32_t* q_term, int nq, int64_t max_query_terms,
                   int64_t N, int k, int64_t id_base, int nonneg, float impact_bound, float* out_score,
                   int64_t* out_id, void* ws, size_t ws_bytes, lrag_stream_t stream);
/* The same with DENSE ROWS for the few terms that occur in most documents (under Zipf two or three terms carry most of
 * the posting volume): dense_term [n_dense] int32 ascending term ids, dense_rows [n_dense, dense_stride] fp32 with
 * dense_rows[r, d] = impact of term dense_term[r] in LOCAL doc d (0 where absent), dense_stride >= N a multiple of
 * 32, rows 128-byte aligned.  Such a term is walked by doc range instead of by (doc, impact) pairs: half the bytes at
 * density > 0.5, no doc ids, conflict-free accumulator updates.  Results are those of lrag_bm25_topk (a posting whose
 * impact is exactly 0 contributes nothing either way); the postings of those terms stay in the CSR (they define df). */
int lrag_bm25_topk_dense(const int64_t* indptr, const int32_t* doc_id, const float* impact, int64_t V, int64_t nnz,
                         const int32_t* dense_term, const float* dense_rows, int n_dense, int64_t dense_stride,
                         const int64_t* q_indptr, const int32_t* q_term, int nq, int64_t max_query_terms,
                         int64_t N, int k, int64_t id_base, int nonneg, float impact_bound, float* out_score,
                         int64_t* out_id, void* ws, size_t ws_bytes, lrag_stream_t stream);
/* The same confined to `max_sms` SMs (0 = the whole machine; the scan keeps as many CTAs resident per SM as fit, two)
 * whose CTAs add one to *start_counter (may be NULL) as they become resident: the BM25 side of an SM partition
 * (lrag_sm_reserve).  lrag_bm25_grid tells how many CTAs such a call launches.  Results do not depend on max_sms. */
int lrag_bm25_grid(int64_t N, int nq, int k, int64_t max_query_terms, int max_sms);
int lrag_bm25_topk_part(const int64_t* indptr, const int32_t* doc_id, const float* impact, int64_t V, int64_t nnz,
                        const int32_t* dense_term, const float* dense_rows, int n_dense, int64_t dense_stride,
                        const int64_t* q_indptr, const int32_t* q_term, int nq, int64_t max_query_terms,
                        int64_t N, int k, int64_t id_base, int nonneg, float impact_bound, int max_sms,
                        unsigned long long* start_counter, float* out_score, int64_t* out_id, void* ws,
                        size_t ws_bytes, lrag_stream_t stream);

/* ---------------------------------------------------------------------------------
 * SM partition for two persistent scans that run side by side (no reference counterpart: the reference runs its
 * channels one after the other on the CPU, legalrag/retrieval/hybrid_retriever.py:295-299).  The dense scan
 * (192 KB of shared memory per CTA) and the BM25 scan (2 x 113 KB per SM) cannot share an SM.  lrag_sm_reserve parks
 * one CTA holding 200 KB of shared memory on `ctas` SMs until *counter >= target (or `timeout_ms` passes): launch
 * it first, then lrag_bm25_topk_part on a second stream with max_sms = SMs - ctas and the same counter
 * (target = counter value before + lrag_bm25_grid(...)); its CTAs can only land on the other SMs.  A third stream
 * that waits for the reservation then runs lrag_dense_topk_bf16_part with max_ctas = ctas on exactly the SMs the
 * reservation gives back.  On a power-capped board the two stages side by side draw a steady load at one clock;
 * one after the other the low-power stage inherits the clock the high-power stage was throttled to. */
int lrag_sm_reserve(int ctas, const unsigned long long* counter, unsigned long long target, int timeout_ms,
                    lrag_stream_t stream);

/* ---------------------------------------------------------------------------------
 * ColBERT channel.  Replaces the scoring inside `Searcher.search(query, k)` at
 * legalrag/retrieval/colbert_retriever.py:152 (colbert_score: sum_i max_j <q_i, d_j>).
 *   D [Nd, Ld, dim] bf16 token store (dim == 128, Ld <= 256, Ld % 8 == 0), doclen [Nd]
 *   int32 or NULL (= Ld), Q [nq, Lq, dim] bf16 (Lq <= 32), cand [nq, C] int64 LOCAL rows
 *   (-1 = skip).  out_id = id_base + row.  Fused: per candidate one tcgen05 tile
 *   (doc tokens x query tokens) whose row-max / sum epilogue runs out of TMEM; the
 *   [Lq, Ld] similarity matrix is never written. */
size_t lrag_maxsim_rerank_workspace_bytes(int nq, int C, int k);
int lrag_maxsim_rerank_bf16(const void* D, const int32_t* doclen, int64_t Nd, int Ld, int dim,
                            const void* Q, int nq, int Lq, const int64_t* cand, int C, int k,
                            int64_t id_base, float* out_score, int64_t* out_id, void* ws,
                            size_t ws_bytes, lrag_stream_t stream);
/* candidate scores only, [nq, C] fp32 (-inf for skipped) -- the multi-GPU path scores the
 * candidates each shard owns and combines them with a max-reduce. */
int lrag_maxsim_scores_bf16(const void* D, const int32_t* doclen, int64_t Nd, int Ld, int dim,
                            const void* Q, int nq, int Lq, const int64_t* cand, int C,
                            float* out_score, lrag_stream_t stream);

/* Full-corpus MaxSim for a query batch (the reference's ColBERT channel is a first-stage retriever over
 * the whole corpus, colbert_retriever.py:152; PLAID approximates this scan): one tensor-core contraction
 * of all query-token rows against all doc-token rows with a MaxSim epilogue out of TMEM, the token store
 * is read from HBM once per batch.  Ld must be 32, 64, 128 or 256.  `scores` writes the [nq, Nd] matrix
 * (row stride ld_out >= Nd, -9999 * Lq for empty documents); `topk` ranks it (ids = id_base + row).
 * Workspace: lrag_maxsim_scan_workspace_bytes(Nd, nq, k) with k = 0 for the scores entry point. */
size_t lrag_maxsim_scan_workspace_bytes(int64_t Nd, int nq, int k);
int lrag_maxsim_scan_scores_bf16(const void* D, const int32_t* doclen, int64_t Nd, int Ld, int dim,
                                 const void* Q, int nq, int Lq, float* out_score, int64_t ld_out,
                                 void* ws, size_t ws_bytes, lrag_stream_t stream);
int lrag_maxsim_scan_topk_bf16(const void* D, const int32_t* doclen, int64_t Nd, int Ld, int dim,
                               const void* Q, int nq, int Lq, int k, int64_t id_base, float* out_score,
                               int64_t* out_id, void* ws, size_t ws_bytes, lrag_stream_t stream);

/* ---------------------------------------------------------------------------------
 * Fusion.  Replaces HybridRetriever._fuse (legalrag/retrieval/hybrid_retriever.py:389-551,
 * _minmax :24-30, _rrf_with_breakdown :33-56) plus the min_final_score filter (:309-310)
 * and the final slice (:384).  Inputs are the three per-channel lists, each [nq, kc],
 * already sorted (score desc), id == -1 = padding; a NULL channel is empty.
 * method: 0 weighted_sum, 1 rrf, 2 wrrf, 3 rrf_norm_blend.
 * out_breakdown (optional, [nq, k, 8] fp32): rrf_norm, weighted_sum, dense_norm,
 * bm25_norm, colbert_norm, contrib_dense, contrib_bm25, contrib_colbert. */
int lrag_fuse_topk(const float* s_dense, const int64_t* i_dense, const float* s_bm25,
                   const int64_t* i_bm25, const float* s_colb, const int64_t* i_colb, int nq,
                   int kc, int k, int method, double w_dense, double w_bm25, double w_colb,
                   int rrf_k, double alpha, double min_final, float* out_score, int64_t* out_id,
                   float* out_breakdown, lrag_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LRAG_H_ */
