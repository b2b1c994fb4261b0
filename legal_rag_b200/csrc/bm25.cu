// BM25 channel: Okapi scoring over term-major CSR postings + exact top-k (replaces
// `bm25.get_scores(tokens)` and the full Python sort at legalrag/retrieval/bm25_retriever.py:74-75
// of the reference; rank_bm25.BM25Okapi semantics, see oracle/bm25.py).
//
// HBM-bound integer/float streaming work, no tensor cores.  One CTA owns one (query, doc-range)
// pair and walks its range slab by slab (BM25_SLAB docs = 64 KB of fp32 accumulators in shared
// memory):
//   1. posting boundaries of every query term for the next BM25_BATCH slabs are found with one
//      parallel round of windowed binary searches (postings are doc-id sorted);
//   2. per slab: zero the accumulators, then term after term stream that term's postings for the
//      slab with coalesced loads and add `mult * impact` into acc[doc - slab0].  Doc ids are unique
//      inside a term, so a term's adds never collide and need no atomics; a barrier separates terms;
//   3. scan the slab, append every score that beats the query's running k-th best to a candidate
//      buffer in shared memory; on overflow a radix select over (buffer U slab) cuts back to k and
//      raises the threshold.
// Slabs in which no query term has a posting are skipped when impacts are known non-negative;
// documents that match nothing (score 0) are then added by the merge step, lowest id first, exactly
// as the reference's stable sort does.  bm25_merge_kernel merges the per-range lists.
#include "common.cuh"
#include "select.cuh"

namespace lrag {

constexpr int BM25_THREADS = 512;
constexpr int BM25_SLAB = 16384;
constexpr int BM25_BATCH = 16;
constexpr int BM25_MAXT = 128;
constexpr int BM25_UNROLL = 4;

struct Bm25Params {
  const int64_t* indptr; const int32_t* doc_id; const float* impact; int64_t V;
  const int64_t* q_indptr; const int32_t* q_term;
  int64_t N; int64_t docs_per_split;
  int nq, k, nonneg, nsplit, cap, P;
  uint64_t* out_keys;   // [nq, nsplit, k]
};

struct Bm25Shared {
  SelectShared sel;
  int64_t t_start[BM25_MAXT];     // first posting of the term
  int32_t t_len[BM25_MAXT];       // df
  int32_t t_cur[BM25_MAXT];       // postings before the current batch (relative)
  float t_mult[BM25_MAXT];        // occurrences of the term in the query
  int32_t raw[BM25_MAXT];
  int32_t owner[BM25_MAXT];
  int32_t bound[BM25_MAXT][BM25_BATCH + 1];
  int32_t slab_any[BM25_BATCH];
  int nt;
  int cand_cnt;
  unsigned long long thr_key;
};

__device__ __forceinline__ int lower_bound_doc(const int32_t* __restrict__ ids, int lo, int hi, int64_t target) {
  // first index in [lo, hi) whose doc id >= target
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    if (int64_t(__ldg(ids + mid)) < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// warp-aggregated append to the shared candidate buffer; returns false on overflow
__device__ __forceinline__ void cand_append(bool want, uint64_t key, uint64_t* cand, int cap, int* cnt) {
  const uint32_t m = __ballot_sync(0xffffffffu, want);
  if (m == 0) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == (__ffs(m) - 1)) base = atomicAdd(cnt, __popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (want) {
    const int pos = base + __popc(m & ((1u << lane) - 1));
    if (pos < cap) cand[pos] = key;
  }
}

struct Bm25Union {
  const uint64_t* cand; int ncand;
  const float* acc; int64_t slab0; int64_t range_end; unsigned long long thr_key; int nonneg;
  template <class F> __device__ void operator()(F&& f) const {
    for (int i = threadIdx.x; i < ncand; i += BM25_THREADS) f(cand[i]);
    for (int i = threadIdx.x; i < BM25_SLAB; i += BM25_THREADS) {
      const int64_t doc = slab0 + i;
      if (doc >= range_end) break;
      const uint64_t key = make_key(acc[i], uint32_t(doc));
      if (key > thr_key) f(key);
    }
  }
};

// One batch of one (slab, term) pair: postings [pos, min(pos + BATCH, hi)) of term t in slab j of the
// current slab group.  t < 0 = no more work in this group.
struct Bm25Cursor { int t, j, pos, hi; };
constexpr int BM25_BATCH_POSTINGS = BM25_THREADS * BM25_UNROLL;

__device__ __forceinline__ Bm25Cursor bm25_seek(const Bm25Shared& sh, int nt, int t, int j, int64_t b0, int64_t range_end,
                                                int nonneg) {
  for (; j < BM25_BATCH && b0 + int64_t(j) * BM25_SLAB < range_end; ++j, t = 0) {
    if (nonneg && !sh.slab_any[j]) continue;
    for (; t < nt; ++t)
      if (sh.bound[t][j + 1] > sh.bound[t][j]) return Bm25Cursor{t, j, sh.bound[t][j], sh.bound[t][j + 1]};
  }
  return Bm25Cursor{-1, BM25_BATCH, 0, 0};
}
__device__ __forceinline__ Bm25Cursor bm25_first(const Bm25Shared& sh, int nt, int64_t b0, int64_t range_end, int nonneg) {
  return bm25_seek(sh, nt, 0, 0, b0, range_end, nonneg);
}
__device__ __forceinline__ Bm25Cursor bm25_next(const Bm25Shared& sh, int nt, const Bm25Cursor& c, int64_t b0, int64_t range_end,
                                               int nonneg) {
  if (c.pos + BM25_BATCH_POSTINGS < c.hi) return Bm25Cursor{c.t, c.j, c.pos + BM25_BATCH_POSTINGS, c.hi};
  return bm25_seek(sh, nt, c.t + 1, c.j, b0, range_end, nonneg);
}
// issue this thread's loads of a batch (doc id < 0 marks an empty lane)
__device__ __forceinline__ void bm25_load(const Bm25Params& p, const Bm25Shared& sh, const Bm25Cursor& c, int tid,
                                          int (&d)[BM25_UNROLL], float (&v)[BM25_UNROLL]) {
  if (c.t < 0) {
#pragma unroll
    for (int u = 0; u < BM25_UNROLL; ++u) { d[u] = -1; v[u] = 0.f; }
    return;
  }
  const int32_t* __restrict__ ids = p.doc_id + sh.t_start[c.t];
  const float* __restrict__ imp = p.impact + sh.t_start[c.t];
#pragma unroll
  for (int u = 0; u < BM25_UNROLL; ++u) {
    const int ii = c.pos + tid + u * BM25_THREADS;
    const bool ok = ii < c.hi;
    d[u] = ok ? __ldg(ids + ii) : -1;
    v[u] = ok ? __ldg(imp + ii) : 0.f;
  }
}

struct Bm25Cands {
  const uint64_t* c; int n;
  template <class F> __device__ void operator()(F&& f) const {
    for (int i = threadIdx.x; i < n; i += BM25_THREADS) f(c[i]);
  }
};

__global__ void __launch_bounds__(BM25_THREADS, 2)
bm25_scan_kernel(const Bm25Params p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* acc = reinterpret_cast<float*>(smem_raw);                                  // [BM25_SLAB]
  uint64_t* cand = reinterpret_cast<uint64_t*>(smem_raw + BM25_SLAB * 4);           // [cap]
  Bm25Shared& sh = *reinterpret_cast<Bm25Shared*>(smem_raw + BM25_SLAB * 4 + size_t(p.cap) * 8);

  const int tid = threadIdx.x;
  const int q = blockIdx.x / p.nsplit;
  const int r = blockIdx.x % p.nsplit;
  const int64_t range_begin = int64_t(r) * p.docs_per_split;
  const int64_t range_end = min(p.N, range_begin + p.docs_per_split);
  const int cap = p.cap;

  // ---- query terms: drop OOV, merge repeats into a multiplicity (first-occurrence order) ----
  const int64_t qs = p.q_indptr[q];
  const int64_t qlen = p.q_indptr[q + 1] - qs;
  const int nraw = int(qlen < BM25_MAXT ? qlen : BM25_MAXT);
  if (tid < BM25_MAXT) {
    int t = -1;
    if (tid < nraw) { t = p.q_term[qs + tid]; if (t < 0 || int64_t(t) >= p.V) t = -1; }
    sh.raw[tid] = t;
  }
  if (tid == 0) {
    sh.cand_cnt = 0;
    sh.thr_key = p.nonneg ? ((uint64_t(ord32(0.0f)) << 32) | 0xffffffffull) : 0ull;
  }
  __syncthreads();
  if (tid < BM25_MAXT) {
    const int t = sh.raw[tid];
    int own = (t >= 0);
    for (int j = 0; j < tid && own; ++j) own = (sh.raw[j] != t);
    if (own && p.indptr[t + 1] == p.indptr[t]) own = 0;   // term without postings
    sh.owner[tid] = own;
  }
  __syncthreads();
  if (tid < BM25_MAXT && sh.owner[tid]) {
    const int t = sh.raw[tid];
    int slot = 0, mult = 0;
    for (int j = 0; j < tid; ++j) slot += sh.owner[j];
    for (int j = tid; j < nraw; ++j) mult += (sh.raw[j] == t);
    const int64_t s = p.indptr[t];
    sh.t_start[slot] = s;
    sh.t_len[slot] = int32_t(p.indptr[t + 1] - s);
    sh.t_mult[slot] = float(mult);
  }
  if (tid == 0) {
    int n = 0;
    for (int j = 0; j < nraw; ++j) n += sh.owner[j];
    sh.nt = n;
  }
  __syncthreads();
  const int nt = sh.nt;
  // position every term at the start of this CTA's doc range
  if (tid < nt) sh.t_cur[tid] = lower_bound_doc(p.doc_id + sh.t_start[tid], 0, sh.t_len[tid], range_begin);
  __syncthreads();

  {
    float4* a4 = reinterpret_cast<float4*>(acc);
#pragma unroll
    for (int i = 0; i < BM25_SLAB / 4 / BM25_THREADS; ++i) a4[tid + i * BM25_THREADS] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const int64_t batch_docs = int64_t(BM25_SLAB) * BM25_BATCH;
  for (int64_t b0 = range_begin; b0 < range_end; b0 += batch_docs) {
    // ---- 1. posting boundaries for the next BM25_BATCH slabs ----
    for (int w = tid; w < nt * (BM25_BATCH + 1); w += BM25_THREADS) {
      const int t = w / (BM25_BATCH + 1), j = w % (BM25_BATCH + 1);
      const int cur = sh.t_cur[t], len = sh.t_len[t];
      const int64_t target = b0 + int64_t(j) * BM25_SLAB;
      // ids are strictly increasing inside a term: the answer is at most (target - b0) past cur
      const int64_t reach = int64_t(cur) + (target - b0);
      const int hi = int(reach < len ? reach : int64_t(len));
      sh.bound[t][j] = (j == 0) ? cur : lower_bound_doc(p.doc_id + sh.t_start[t], cur, hi, target);
    }
    __syncthreads();
    if (tid < BM25_BATCH) {
      int any = 0;
      for (int t = 0; t < nt; ++t) any |= (sh.bound[t][tid + 1] > sh.bound[t][tid]);
      sh.slab_any[tid] = any;
    }
    if (tid < nt) sh.t_cur[tid] = sh.bound[tid][BM25_BATCH];
    __syncthreads();

    // ---- 2. per slab: pipelined term-at-a-time accumulation, scan, re-zero ----
    // The postings of a (slab, term) pair are consumed in batches of BM25_THREADS * BM25_UNROLL; the
    // loads of the NEXT batch (possibly of the next term or the next slab) are issued before the
    // current batch is added, so the HBM latency of one batch hides behind the shared-memory work
    // and the barrier of the previous one.
    Bm25Cursor cur = bm25_first(sh, nt, b0, range_end, p.nonneg);
    int d_cur[BM25_UNROLL]; float v_cur[BM25_UNROLL];
    bm25_load(p, sh, cur, tid, d_cur, v_cur);
    for (int j = 0; j < BM25_BATCH; ++j) {
      const int64_t slab0 = b0 + int64_t(j) * BM25_SLAB;
      if (slab0 >= range_end) break;
      if (p.nonneg && !sh.slab_any[j]) continue;
      const int sl0 = int(slab0);   // doc ids are int32
      while (cur.t >= 0 && cur.j == j) {
        const Bm25Cursor nxt = bm25_next(sh, nt, cur, b0, range_end, p.nonneg);
        int d_nxt[BM25_UNROLL]; float v_nxt[BM25_UNROLL];
        bm25_load(p, sh, nxt, tid, d_nxt, v_nxt);
        const float mult = sh.t_mult[cur.t];
        float a[BM25_UNROLL];
#pragma unroll
        for (int u = 0; u < BM25_UNROLL; ++u) a[u] = (d_cur[u] >= 0) ? acc[d_cur[u] - sl0] : 0.f;
#pragma unroll
        for (int u = 0; u < BM25_UNROLL; ++u) if (d_cur[u] >= 0) acc[d_cur[u] - sl0] = fmaf(mult, v_cur[u], a[u]);
        // same-term batches never touch the same doc: a barrier is only needed when the term changes
        if (nxt.t != cur.t || nxt.j != j) __syncthreads();
        cur = nxt;
#pragma unroll
        for (int u = 0; u < BM25_UNROLL; ++u) { d_cur[u] = d_nxt[u]; v_cur[u] = v_nxt[u]; }
      }
      // ---- 3. scan the slab for candidates (fast path: nothing in this thread's 32 docs beats thr) ----
      const int cnt_before = sh.cand_cnt;
      const unsigned long long thr_key = sh.thr_key;
      // initial thresholds: nothing yet (-inf), or "strictly positive" when zero scores are filled in later
      const float thr_s = !thr_key ? -INFINITY : (uint32_t(thr_key) == 0xffffffffu && key_score(thr_key) == 0.f) ? 1.4e-45f : key_score(thr_key);
      __syncthreads();   // everyone has read cand_cnt before anyone appends
      {
        const float4* a4 = reinterpret_cast<const float4*>(acc);
        float4 s4[BM25_SLAB / 4 / BM25_THREADS];
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < BM25_SLAB / 4 / BM25_THREADS; ++i) {
          s4[i] = a4[tid + i * BM25_THREADS];
          mx = fmaxf(mx, fmaxf(fmaxf(s4[i].x, s4[i].y), fmaxf(s4[i].z, s4[i].w)));
        }
        if (__any_sync(0xffffffffu, mx >= thr_s)) {
#pragma unroll
          for (int i = 0; i < BM25_SLAB / 4 / BM25_THREADS; ++i) {
            const int idx = (tid + i * BM25_THREADS) * 4;
            const float sv[4] = {s4[i].x, s4[i].y, s4[i].z, s4[i].w};
            const bool any4 = fmaxf(fmaxf(sv[0], sv[1]), fmaxf(sv[2], sv[3])) >= thr_s;
            if (!__any_sync(0xffffffffu, any4)) continue;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int64_t doc = slab0 + idx + e;
              bool want = (sv[e] >= thr_s) && (doc < range_end);
              uint64_t key = 0;
              if (want) { key = make_key(sv[e], uint32_t(doc)); want = key > thr_key; }
              cand_append(want, key, cand, cap, &sh.cand_cnt);
            }
          }
        }
      }
      __syncthreads();
      if (sh.cand_cnt > cap) {
        // ---- overflow: exact k-th best of (buffer U slab) becomes the new threshold ----
        Bm25Union uni{cand, cnt_before, acc, slab0, range_end, thr_key, p.nonneg};
        const unsigned long long pivot = block_select_pivot(uni, p.k, sh.sel);
        uint64_t keep[4];
        int nkeep = 0;
        for (int i = tid; i < cnt_before; i += BM25_THREADS) { keep[nkeep & 3] = cand[i]; ++nkeep; }
        __syncthreads();
        if (tid == 0) sh.cand_cnt = 0;
        __syncthreads();
        // cap <= 4 * BM25_THREADS, so every thread holds at most 4 old keys
        for (int i = 0; i < 4; ++i) {
          const bool want = (i < nkeep) && (keep[i] >= pivot);
          cand_append(want, want ? keep[i] : 0, cand, cap, &sh.cand_cnt);
        }
        for (int i = tid; i < BM25_SLAB; i += BM25_THREADS) {
          const int64_t doc = slab0 + i;
          uint64_t key = 0;
          bool want = doc < range_end;
          if (want) { key = make_key(acc[i], uint32_t(doc)); want = (key > thr_key) && (key >= pivot); }
          cand_append(want, key, cand, cap, &sh.cand_cnt);
        }
        __syncthreads();
        if (tid == 0 && pivot > sh.thr_key) sh.thr_key = pivot;
        __syncthreads();
      }
      // ---- re-zero the slab for the next one ----
      {
        float4* a4 = reinterpret_cast<float4*>(acc);
#pragma unroll
        for (int i = 0; i < BM25_SLAB / 4 / BM25_THREADS; ++i) a4[tid + i * BM25_THREADS] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();
    }
  }

  // ---- sorted top-k of the surviving candidates -> this (query, range)'s key list ----
  {
    __syncthreads();
    const int ncand = min(sh.cand_cnt, cap);
    Bm25Cands cands{cand, ncand};
    const unsigned long long pivot = block_select_pivot(cands, p.k, sh.sel);
    // winners -> acc (reused as the sort buffer), then sort
    uint64_t* sortbuf = reinterpret_cast<uint64_t*>(acc);
    const int P = p.P;
    for (int i = tid; i < P; i += BM25_THREADS) sortbuf[i] = 0;
    if (tid == 0) sh.sel.nsel = 0;
    __syncthreads();
    for (int i = tid; i < ncand; i += BM25_THREADS) {
      const uint64_t key = cand[i];
      if (key >= pivot) { const int pos = atomicAdd(&sh.sel.nsel, 1); if (pos < P) sortbuf[pos] = key; }
    }
    __syncthreads();
    block_sort_desc(sortbuf, P);
    uint64_t* out = p.out_keys + (size_t(q) * p.nsplit + r) * p.k;
    for (int i = tid; i < p.k; i += BM25_THREADS) out[i] = sortbuf[i];
  }
}

// Merge the per-range key lists of one query; when impacts are non-negative, documents that matched
// nothing score 0 and follow the matched ones in ascending id order (bm25_retriever.py:75 keeps them).
struct Bm25Lists {
  const uint64_t* keys; int n;
  template <class F> __device__ void operator()(F&& f) const {
    for (int i = threadIdx.x; i < n; i += SELECT_THREADS) { const uint64_t key = keys[i]; if (key) f(key); }
  }
};

__global__ void __launch_bounds__(SELECT_THREADS)
bm25_merge_kernel(const uint64_t* keys, int nsplit, int k, int P, int64_t N, int64_t id_base, int nonneg,
                  float* out_score, int64_t* out_id) {
  extern __shared__ uint8_t sm_raw[];
  __shared__ SelectShared ss;
  uint64_t* sel_key = reinterpret_cast<uint64_t*>(sm_raw);
  const int q = blockIdx.x;
  float* os = out_score + size_t(q) * k;
  int64_t* oi = out_id + size_t(q) * k;
  Bm25Lists lists{keys + size_t(q) * nsplit * k, nsplit * k};
  const int n = block_topk_sorted(lists, k, P, ss, sel_key, id_base, os, oi, static_cast<uint64_t*>(nullptr));
  if (!nonneg || n >= k) return;
  // zero-score fill: the first (k - n) local doc ids that are not among the n positive hits.
  // At most n of the ids [0, k) are taken, so scanning [0, k + n) is enough.
  const int64_t lim = N < int64_t(k) + n ? N : int64_t(k) + n;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int placed = 0;
    for (int64_t c0 = 0; c0 < lim && placed < k - n; c0 += 32) {
      const int64_t c = c0 + lane;
      bool free_id = c < lim;
      if (free_id) for (int i = 0; i < n; ++i) if (key_id(sel_key[i]) == uint32_t(c)) { free_id = false; break; }
      const uint32_t m = __ballot_sync(0xffffffffu, free_id);
      const int pos = n + placed + __popc(m & ((1u << lane) - 1));
      if (free_id && pos < k) { os[pos] = 0.0f; oi[pos] = id_base + c; }
      placed += __popc(m);
    }
  }
}

struct Bm25Plan { int nsplit, cap, P; int64_t docs_per_split; size_t smem, ws; };

static Bm25Plan bm25_plan(int64_t N, int nq, int k, int sms) {
  Bm25Plan pl;
  pl.P = next_pow2(k);
  pl.cap = 2 * pl.P < 1024 ? 1024 : 2 * pl.P;        // <= 2048 = 4 * BM25_THREADS
  const int64_t nslab = (N + BM25_SLAB - 1) / BM25_SLAB;
  int64_t want = (int64_t(4) * sms + nq - 1) / nq;   // enough CTAs to fill the machine twice over
  if (want < 1) want = 1;
  if (want > nslab) want = nslab > 0 ? nslab : 1;
  int64_t slabs_per = (nslab + want - 1) / want;
  if (slabs_per < 1) slabs_per = 1;
  pl.docs_per_split = slabs_per * BM25_SLAB;
  pl.nsplit = int((N + pl.docs_per_split - 1) / pl.docs_per_split);
  if (pl.nsplit < 1) pl.nsplit = 1;
  pl.smem = size_t(BM25_SLAB) * 4 + size_t(pl.cap) * 8 + sizeof(Bm25Shared);
  pl.ws = align_up(size_t(nq) * pl.nsplit * k * 8, 256);
  return pl;
}

}  // namespace lrag

using namespace lrag;

extern "C" size_t lrag_bm25_topk_workspace_bytes(int64_t N, int nq, int k, int64_t max_query_terms) {
  (void)max_query_terms;
  if (N < 0 || nq <= 0 || k <= 0) return 0;
  return bm25_plan(N, nq, k, sm_count()).ws;
}

extern "C" int lrag_bm25_topk(const int64_t* indptr, const int32_t* doc_id, const float* impact, int64_t V,
                              const int64_t* q_indptr, const int32_t* q_term, int nq, int64_t max_query_terms,
                              int64_t N, int k, int64_t id_base, int nonneg, float* out_score, int64_t* out_id,
                              void* ws, size_t ws_bytes, lrag_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LRAG_REQUIRE(initialised(), "lrag_init has not been called");
  LRAG_REQUIRE(nq > 0 && k > 0 && k <= LRAG_MAX_K, "bm25_topk: need nq > 0 and 1 <= k <= %d (nq=%d k=%d)", LRAG_MAX_K, nq, k);
  LRAG_REQUIRE(N >= 0 && N < (int64_t(1) << 31), "bm25_topk: N=%lld out of range for one shard", (long long)N);
  LRAG_REQUIRE(V >= 0 && V < (int64_t(1) << 31), "bm25_topk: V=%lld out of range", (long long)V);
  LRAG_REQUIRE(max_query_terms >= 0 && max_query_terms <= LRAG_BM25_MAX_QUERY_TERMS,
               "bm25_topk: a query has %lld terms; at most %d are supported", (long long)max_query_terms,
               LRAG_BM25_MAX_QUERY_TERMS);
  LRAG_REQUIRE(indptr && q_indptr && out_score && out_id, "bm25_topk: null pointer");
  const Bm25Plan pl = bm25_plan(N, nq, k, sm_count());
  if (ws_bytes < pl.ws || !ws) { set_error("bm25_topk: workspace %zu < required %zu", ws_bytes, pl.ws); return LRAG_ENOSPC; }
  Bm25Params p;
  p.indptr = indptr; p.doc_id = doc_id; p.impact = impact; p.V = V; p.q_indptr = q_indptr; p.q_term = q_term;
  p.N = N; p.docs_per_split = pl.docs_per_split; p.nq = nq; p.k = k; p.nonneg = nonneg ? 1 : 0;
  p.nsplit = pl.nsplit; p.cap = pl.cap; p.P = pl.P; p.out_keys = static_cast<uint64_t*>(ws);
  static bool attr_set = false;
  if (!attr_set) {
    LRAG_CHECK_CUDA(cudaFuncSetAttribute(bm25_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    attr_set = true;
  }
  prof_begin(stream, PROF_BM25_SCAN);
  bm25_scan_kernel<<<unsigned(int64_t(nq) * pl.nsplit), BM25_THREADS, pl.smem, stream>>>(p);
  prof_end(stream);
  LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
  bm25_merge_kernel<<<nq, SELECT_THREADS, select_smem_bytes(k), stream>>>(p.out_keys, pl.nsplit, k, pl.P, N, id_base,
                                                                          p.nonneg, out_score, out_id);
  LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
  return LRAG_OK;
}
