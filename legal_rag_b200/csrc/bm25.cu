// BM25 channel: Okapi scoring over term-major CSR postings + exact top-k (replaces
// `bm25.get_scores(tokens)` and the full Python sort at legalrag/retrieval/bm25_retriever.py:74-75
// of the reference; rank_bm25.BM25Okapi semantics, see oracle/bm25.py).
//
// HBM-bound integer/float streaming work, no tensor cores.  One CTA owns one (query, doc-range)
// pair and walks its range slab by slab (BM25_SLAB docs = 64 KB of fp32 accumulators in shared
// memory).  The CTA is warp-specialised so that posting traffic never waits for the arithmetic and
// the arithmetic never does bookkeeping:
//   * bounds warp   -- for the next group of slabs, finds the posting boundaries of every query term
//                      with one parallel round of windowed binary searches (postings are doc-id
//                      sorted); double-buffered, one group ahead of the copy warp;
//   * copy warp     -- walks the (slab, term) runs in order and streams them, <= BM25_CHUNK postings at
//                      a time, into a shared-memory ring with bulk async copies (cp.async.bulk, 16-byte
//                      aligned source windows) that complete on mbarriers.  Each ring stage carries a
//                      descriptor {count, skip, term multiplicity, slab base, flags}: the consumers
//                      are a plain interpreter of that stream;
//   * 16 consumer warps -- per stage add `mult * impact` into acc[doc - slab0] and keep a per-thread
//                      running max of what they wrote.  Doc ids are unique inside a term, so a
//                      term's adds never collide and need no atomics; a named barrier separates terms
//                      (flag TERM_END).  At SLAB_END one `bar.red.or` tells whether any thread wrote
//                      a score that reaches the query's running k-th best: only then is the slab
//                      scanned for candidates (appended to a shared buffer; an overflow triggers an
//                      exact radix select that raises the threshold).  The slab is re-zeroed.
// Slabs in which no query term has a posting are skipped when impacts are known non-negative;
// documents that match nothing (score 0) are then added by the merge step, lowest id first, exactly
// as the reference's stable sort does.  bm25_merge_kernel merges the per-range lists.
#include "common.cuh"
#include "select.cuh"

namespace lrag {

constexpr int BM25_CONSUMERS = 512;                  // 16 warps
constexpr int BM25_THREADS = BM25_CONSUMERS + 64;    // + copy warp (16) + bounds warp (17)
constexpr int BM25_SLAB = 16384;
constexpr int BM25_MAX_GROUP = 16;                   // slabs per bounds group (fewer when a query has many terms)
constexpr int BM25_BOUND_CAP = 17 * 32;              // ints per bounds buffer: (group + 1) * nt must fit
constexpr int BM25_MAXT = 128;
constexpr int BM25_CHUNK = 1024;                     // postings per ring stage
constexpr int BM25_STAGES = 3;                      // 64 KB slab + 24 KB ring + 8 KB candidates + state: two CTAs per SM
constexpr int BM25_RING_BYTES = BM25_STAGES * BM25_CHUNK * 8;
constexpr int BM25_BAR_CONSUMERS = 1;                // named barrier id of the consumer warps
constexpr int BM25_PER_THREAD = BM25_CHUNK / BM25_CONSUMERS;
constexpr int BM25_F_TERM_END = 1, BM25_F_SLAB_END = 2, BM25_F_END = 4;

struct Bm25Params {
  const int64_t* indptr; const int32_t* doc_id; const float* impact; int64_t V; int64_t nnz;
  const int64_t* q_indptr; const int32_t* q_term;
  int64_t N; int64_t docs_per_split;
  int nq, k, nonneg, nsplit, cap, P;
  uint64_t* out_keys;   // [nq, nsplit, k]
};

struct Bm25Shared {
  SelectShared sel;
  int4 sdesc[BM25_STAGES];        // {n, skip | flags << 8, mult (float bits), slab0}
  uint64_t full_bar[BM25_STAGES], empty_bar[BM25_STAGES];   // posting ring
  uint64_t bfull_bar[2], bempty_bar[2];                      // bounds buffers
  int64_t t_start[BM25_MAXT];     // first posting of the term
  int32_t t_len[BM25_MAXT];       // df
  int32_t t_cur[BM25_MAXT];       // bounds warp only: postings before the next group (relative)
  float t_mult[BM25_MAXT];        // occurrences of the term in the query
  int32_t raw[BM25_MAXT];
  int32_t owner[BM25_MAXT];
  int32_t bound[2][BM25_BOUND_CAP];     // [buffer][t * (gs + 1) + j]
  int32_t slab_any[2][BM25_MAX_GROUP];
  int nt, gs;
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

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// barrier over the consumer warps that also ORs a predicate across them
__device__ __forceinline__ bool consumers_bar_or(bool pred) {
  uint32_t out;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 q, %1, 0;\n\t"
      "bar.red.or.pred p, %2, %3, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(out)
      : "r"(uint32_t(pred)), "r"(BM25_BAR_CONSUMERS), "r"(BM25_CONSUMERS)
      : "memory");
  return out != 0;
}

// warp-aggregated append to the shared candidate buffer (entries past `cap` are dropped and counted)
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
  const float* acc; int64_t slab0; int64_t range_end; unsigned long long thr_key;
  template <class F> __device__ void operator()(F&& f) const {
    for (int i = threadIdx.x; i < ncand; i += BM25_CONSUMERS) f(cand[i]);
    for (int i = threadIdx.x; i < BM25_SLAB; i += BM25_CONSUMERS) {
      const int64_t doc = slab0 + i;
      if (doc >= range_end) break;
      const uint64_t key = make_key(acc[i], uint32_t(doc));
      if (key > thr_key) f(key);
    }
  }
};

struct Bm25Cands {
  const uint64_t* c; int n;
  template <class F> __device__ void operator()(F&& f) const {
    for (int i = threadIdx.x; i < n; i += BM25_CONSUMERS) f(c[i]);
  }
};

// One ring stage: postings [pos, pos + n) of term t (t < 0: none) for the slab starting at doc sl0.
struct Bm25Chunk { int t, pos, n, sl0; };

__global__ void __launch_bounds__(BM25_THREADS, 2)
bm25_scan_kernel(const Bm25Params p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* acc = reinterpret_cast<float*>(smem_raw);                                          // [BM25_SLAB]
  int32_t* ring_id = reinterpret_cast<int32_t*>(smem_raw + BM25_SLAB * 4);                  // [STAGES][CHUNK]
  float* ring_imp = reinterpret_cast<float*>(smem_raw + BM25_SLAB * 4 + BM25_RING_BYTES / 2);
  uint64_t* cand = reinterpret_cast<uint64_t*>(smem_raw + BM25_SLAB * 4 + BM25_RING_BYTES); // [cap]
  Bm25Shared& sh = *reinterpret_cast<Bm25Shared*>(smem_raw + BM25_SLAB * 4 + BM25_RING_BYTES + size_t(p.cap) * 8);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
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
    for (int s = 0; s < BM25_STAGES; ++s) { mbar_init(&sh.full_bar[s], 1); mbar_init(&sh.empty_bar[s], BM25_CONSUMERS / 32); }
    for (int s = 0; s < 2; ++s) { mbar_init(&sh.bfull_bar[s], 1); mbar_init(&sh.bempty_bar[s], 1); }
    fence_barrier_init();
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
    int gs = n > 0 ? BM25_BOUND_CAP / n - 1 : BM25_MAX_GROUP;
    sh.gs = gs < 1 ? 1 : (gs > BM25_MAX_GROUP ? BM25_MAX_GROUP : gs);
  }
  // zero the slab once; every slab end leaves it zeroed again
  for (int i = tid; i < BM25_SLAB / 4; i += BM25_THREADS) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const int nt = sh.nt, gs = sh.gs;
  const int64_t group_docs = int64_t(BM25_SLAB) * gs;
  const NamedBarrier cbar{BM25_BAR_CONSUMERS, BM25_CONSUMERS};

  if (warp == BM25_CONSUMERS / 32 + 1) {
    // ===================== bounds warp =====================
    for (int t = lane; t < nt; t += 32)
      sh.t_cur[t] = range_begin == 0 ? 0 : lower_bound_doc(p.doc_id + sh.t_start[t], 0, sh.t_len[t], range_begin);
    __syncwarp();
    uint32_t g = 0;
    for (int64_t b0 = range_begin; b0 < range_end; b0 += group_docs, ++g) {
      const uint32_t bb = g & 1;
      mbar_wait(&sh.bempty_bar[bb], ((g >> 1) & 1) ^ 1);
      int32_t* bound = sh.bound[bb];
      for (int w = lane; w < nt * (gs + 1); w += 32) {
        const int t = w / (gs + 1), j = w % (gs + 1);
        const int cur = sh.t_cur[t], len = sh.t_len[t];
        const int64_t target = b0 + int64_t(j) * BM25_SLAB;
        // ids are strictly increasing inside a term: the answer is at most (target - b0) past cur
        const int64_t reach = int64_t(cur) + (target - b0);
        const int hi = int(reach < len ? reach : int64_t(len));
        bound[w] = (j == 0) ? cur : lower_bound_doc(p.doc_id + sh.t_start[t], cur, hi, target);
      }
      __syncwarp();
      if (lane < gs) {
        int any = 0;
        for (int t = 0; t < nt; ++t) any |= (bound[t * (gs + 1) + lane + 1] > bound[t * (gs + 1) + lane]);
        sh.slab_any[bb][lane] = any;
      }
      __syncwarp();
      for (int t = lane; t < nt; t += 32) sh.t_cur[t] = bound[t * (gs + 1) + gs];
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.bfull_bar[bb]);
    }
  } else if (warp == BM25_CONSUMERS / 32) {
    // ===================== copy warp: postings -> shared-memory ring =====================
    // A chunk is emitted once its successor is known (that decides its TERM_END / SLAB_END flags).
    uint32_t c = 0;
    auto emit = [&](const Bm25Chunk& ch, int flags) {
      const uint32_t s = c % BM25_STAGES;
      mbar_wait(&sh.empty_bar[s], ((c / BM25_STAGES) & 1) ^ 1);
      ++c;
      int skip = 0, cnt4 = 0;
      float mult = 0.f;
      if (ch.n > 0) {
        const int64_t first = sh.t_start[ch.t] + ch.pos;
        skip = int(first & 3);
        const int64_t a0 = first - skip;                                   // multiple of 4 elements = 16 bytes
        const int cnt = ch.n + skip;
        cnt4 = cnt & ~3;
        int32_t* dst_id = ring_id + s * BM25_CHUNK;
        float* dst_imp = ring_imp + s * BM25_CHUNK;
        // the (at most 3) elements past the last whole 16-byte unit are moved by hand
        if (lane < cnt - cnt4) {
          dst_id[cnt4 + lane] = __ldg(p.doc_id + a0 + cnt4 + lane);
          dst_imp[cnt4 + lane] = __ldg(p.impact + a0 + cnt4 + lane);
        }
        mult = sh.t_mult[ch.t];
        __syncwarp();
        if (lane == 0) {
          sh.sdesc[s] = make_int4(ch.n, skip | (flags << 8), __float_as_int(mult), ch.sl0);
          mbar_arrive_expect_tx(&sh.full_bar[s], uint32_t(cnt4) * 8u);
          if (cnt4 > 0) {
            bulk_copy_g2s(dst_id, p.doc_id + a0, uint32_t(cnt4) * 4u, &sh.full_bar[s]);
            bulk_copy_g2s(dst_imp, p.impact + a0, uint32_t(cnt4) * 4u, &sh.full_bar[s]);
          }
        }
      } else if (lane == 0) {
        sh.sdesc[s] = make_int4(0, flags << 8, 0, ch.sl0);
        mbar_arrive(&sh.full_bar[s]);
      }
      __syncwarp();
    };
    Bm25Chunk pend{-1, 0, 0, 0};
    bool have = false;
    auto push = [&](const Bm25Chunk& ch) {
      if (have) {
        const bool slab_end = pend.sl0 != ch.sl0;
        emit(pend, (slab_end ? (BM25_F_SLAB_END | BM25_F_TERM_END) : 0) | ((slab_end || pend.t != ch.t) ? BM25_F_TERM_END : 0));
      }
      pend = ch;
      have = true;
    };
    uint32_t g = 0;
    for (int64_t b0 = range_begin; b0 < range_end; b0 += group_docs, ++g) {
      const uint32_t bb = g & 1;
      mbar_wait(&sh.bfull_bar[bb], (g >> 1) & 1);
      const int32_t* bound = sh.bound[bb];
      const int32_t* any = sh.slab_any[bb];
      for (int j = 0; j < gs; ++j) {
        const int64_t slab0 = b0 + int64_t(j) * BM25_SLAB;
        if (slab0 >= range_end) break;
        if (!any[j]) {
          if (!p.nonneg) push(Bm25Chunk{-1, 0, 0, int(slab0)});   // a slab of zero scores still has to be ranked
          continue;
        }
        for (int t = 0; t < nt; ++t) {
          const int lo = bound[t * (gs + 1) + j], hi = bound[t * (gs + 1) + j + 1];
          for (int pos = lo; pos < hi;) {
            const int skip = int((sh.t_start[t] + pos) & 3);
            const int n = min(hi - pos, BM25_CHUNK - skip);
            push(Bm25Chunk{t, pos, n, int(slab0)});
            pos += n;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.bempty_bar[bb]);   // `pend` holds copies, not references into bound[]
    }
    if (have) emit(pend, BM25_F_TERM_END | BM25_F_SLAB_END | BM25_F_END);
    else emit(Bm25Chunk{-1, 0, 0, 0}, BM25_F_END);
  } else {
    // ===================== consumers =====================
    float mx = -INFINITY;     // largest score this thread wrote into the current slab
    for (uint32_t c = 0;; ++c) {
      const uint32_t s = c % BM25_STAGES;
      mbar_wait(&sh.full_bar[s], (c / BM25_STAGES) & 1);
      const int4 de = sh.sdesc[s];
      const int n = de.x, skip = de.y & 0xff, flags = de.y >> 8, sl0 = de.w;
      const float mult = __int_as_float(de.z);
      {
        const int32_t* ids = ring_id + s * BM25_CHUNK + skip;
        const float* imp = ring_imp + s * BM25_CHUNK + skip;
        float* accr = acc - sl0;                   // accr[doc] == acc[doc - slab0]
        int d[BM25_PER_THREAD]; float v[BM25_PER_THREAD], a[BM25_PER_THREAD];
#pragma unroll
        for (int u = 0; u < BM25_PER_THREAD; ++u) {
          const int i = tid + u * BM25_CONSUMERS;
          const bool ok = i < n;
          d[u] = ok ? ids[i] : -1;
          v[u] = ok ? imp[i] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < BM25_PER_THREAD; ++u) a[u] = (d[u] >= 0) ? accr[d[u]] : 0.f;
#pragma unroll
        for (int u = 0; u < BM25_PER_THREAD; ++u)
          if (d[u] >= 0) { a[u] = fmaf(mult, v[u], a[u]); accr[d[u]] = a[u]; mx = fmaxf(mx, a[u]); }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.empty_bar[s]);     // this warp is done reading the stage
      if (flags & BM25_F_SLAB_END) {
        const int64_t slab0 = sl0;
        const int cnt_before = sh.cand_cnt;
        const unsigned long long thr_key = sh.thr_key;
        // initial thresholds: nothing yet (-inf), or "strictly positive" when zero scores are filled in later
        const float thr_s = !thr_key ? -INFINITY
                            : (uint32_t(thr_key) == 0xffffffffu && key_score(thr_key) == 0.f) ? 1.4e-45f : key_score(thr_key);
        // closes the slab's last term; with negative impacts an untouched doc (score 0) can be a hit too
        const bool hit = consumers_bar_or(!p.nonneg || mx >= thr_s);
        if (hit) {
          // ---- scan the slab for candidates ----
          const float4* a4 = reinterpret_cast<const float4*>(acc);
#pragma unroll 2
          for (int i = 0; i < BM25_SLAB / 4 / BM25_CONSUMERS; ++i) {
            const int idx = (tid + i * BM25_CONSUMERS) * 4;
            const float4 s4 = a4[tid + i * BM25_CONSUMERS];
            const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
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
          cbar();
          if (sh.cand_cnt > cap) {
            // ---- overflow: exact k-th best of (buffer U slab) becomes the new threshold ----
            Bm25Union uni{cand, cnt_before, acc, slab0, range_end, thr_key};
            const unsigned long long pivot = block_select_pivot(uni, p.k, sh.sel, cbar);
            uint64_t keep[4];     // cap <= 4 * BM25_CONSUMERS: at most 4 old keys per thread
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int idx = tid + i * BM25_CONSUMERS;
              keep[i] = idx < cnt_before ? cand[idx] : 0ull;
            }
            cbar();
            if (tid == 0) sh.cand_cnt = 0;
            cbar();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const bool want = keep[i] != 0ull && keep[i] >= pivot;
              cand_append(want, keep[i], cand, cap, &sh.cand_cnt);
            }
            for (int i = tid; i < BM25_SLAB; i += BM25_CONSUMERS) {
              const int64_t doc = slab0 + i;
              uint64_t key = 0;
              bool want = doc < range_end;
              if (want) { key = make_key(acc[i], uint32_t(doc)); want = (key > thr_key) && (key >= pivot); }
              cand_append(want, key, cand, cap, &sh.cand_cnt);
            }
            cbar();
            if (tid == 0 && pivot > sh.thr_key) sh.thr_key = pivot;
            cbar();
          }
        }
        // ---- re-zero the slab for the next one ----
        float4* z4 = reinterpret_cast<float4*>(acc);
#pragma unroll
        for (int i = 0; i < BM25_SLAB / 4 / BM25_CONSUMERS; ++i) z4[tid + i * BM25_CONSUMERS] = make_float4(0.f, 0.f, 0.f, 0.f);
        mx = -INFINITY;
        cbar();
      } else if (flags & BM25_F_TERM_END) {
        cbar();   // the next term may touch the docs this one did
      }
      if (flags & BM25_F_END) break;
    }

    // ---- sorted top-k of the surviving candidates -> this (query, range)'s key list ----
    cbar();
    const int ncand = min(sh.cand_cnt, cap);
    Bm25Cands cands{cand, ncand};
    const unsigned long long pivot = block_select_pivot(cands, p.k, sh.sel, cbar);
    uint64_t* sortbuf = reinterpret_cast<uint64_t*>(acc);      // the slab doubles as the sort buffer
    const int P = p.P;
    for (int i = tid; i < P; i += BM25_CONSUMERS) sortbuf[i] = 0;
    if (tid == 0) sh.sel.nsel = 0;
    cbar();
    for (int i = tid; i < ncand; i += BM25_CONSUMERS) {
      const uint64_t key = cand[i];
      if (key >= pivot) { const int pos = atomicAdd(&sh.sel.nsel, 1); if (pos < P) sortbuf[pos] = key; }
    }
    cbar();
    block_sort_desc(sortbuf, P, cbar);
    uint64_t* out = p.out_keys + (size_t(q) * p.nsplit + r) * p.k;
    for (int i = tid; i < p.k; i += BM25_CONSUMERS) out[i] = sortbuf[i];
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
  pl.cap = 2 * pl.P < 1024 ? 1024 : 2 * pl.P;        // <= 2048 = 4 * BM25_CONSUMERS
  const int64_t nslab = (N + BM25_SLAB - 1) / BM25_SLAB;
  int64_t want = (int64_t(4) * sms + nq - 1) / nq;   // enough CTAs to fill the machine twice over
  if (want < 1) want = 1;
  if (want > nslab) want = nslab > 0 ? nslab : 1;
  int64_t slabs_per = (nslab + want - 1) / want;
  if (slabs_per < 1) slabs_per = 1;
  pl.docs_per_split = slabs_per * BM25_SLAB;
  pl.nsplit = int((N + pl.docs_per_split - 1) / pl.docs_per_split);
  if (pl.nsplit < 1) pl.nsplit = 1;
  pl.smem = size_t(BM25_SLAB) * 4 + BM25_RING_BYTES + size_t(pl.cap) * 8 + sizeof(Bm25Shared);
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

extern "C" int lrag_bm25_topk(const int64_t* indptr, const int32_t* doc_id, const float* impact, int64_t V, int64_t nnz,
                              const int64_t* q_indptr, const int32_t* q_term, int nq, int64_t max_query_terms,
                              int64_t N, int k, int64_t id_base, int nonneg, float* out_score, int64_t* out_id,
                              void* ws, size_t ws_bytes, lrag_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LRAG_REQUIRE(initialised(), "lrag_init has not been called");
  LRAG_REQUIRE(nq > 0 && k > 0 && k <= LRAG_MAX_K, "bm25_topk: need nq > 0 and 1 <= k <= %d (nq=%d k=%d)", LRAG_MAX_K, nq, k);
  LRAG_REQUIRE(N >= 0 && N < (int64_t(1) << 31), "bm25_topk: N=%lld out of range for one shard", (long long)N);
  LRAG_REQUIRE(V >= 0 && V < (int64_t(1) << 31) && nnz >= 0, "bm25_topk: V=%lld nnz=%lld out of range", (long long)V, (long long)nnz);
  LRAG_REQUIRE(max_query_terms >= 0 && max_query_terms <= LRAG_BM25_MAX_QUERY_TERMS,
               "bm25_topk: a query has %lld terms; at most %d are supported", (long long)max_query_terms,
               LRAG_BM25_MAX_QUERY_TERMS);
  LRAG_REQUIRE(indptr && q_indptr && out_score && out_id, "bm25_topk: null pointer");
  LRAG_REQUIRE((reinterpret_cast<uintptr_t>(doc_id) & 15) == 0 && (reinterpret_cast<uintptr_t>(impact) & 15) == 0,
               "bm25_topk: doc_id and impact must be 16-byte aligned (bulk async copies)");
  const Bm25Plan pl = bm25_plan(N, nq, k, sm_count());
  if (ws_bytes < pl.ws || !ws) { set_error("bm25_topk: workspace %zu < required %zu", ws_bytes, pl.ws); return LRAG_ENOSPC; }
  Bm25Params p;
  p.indptr = indptr; p.doc_id = doc_id; p.impact = impact; p.V = V; p.nnz = nnz; p.q_indptr = q_indptr; p.q_term = q_term;
  p.N = N; p.docs_per_split = pl.docs_per_split; p.nq = nq; p.k = k; p.nonneg = nonneg ? 1 : 0;
  p.nsplit = pl.nsplit; p.cap = pl.cap; p.P = pl.P; p.out_keys = static_cast<uint64_t*>(ws);
  static bool attr_set = false;
  if (!attr_set) {
    LRAG_CHECK_CUDA(cudaFuncSetAttribute(bm25_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 118 * 1024));
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
