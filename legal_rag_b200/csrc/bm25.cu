// BM25 channel: Okapi scoring over term-major CSR postings + exact top-k (replaces
// `bm25.get_scores(tokens)` and the full Python sort at legalrag/retrieval/bm25_retriever.py:74-75
// of the reference; rank_bm25.BM25Okapi semantics, see oracle/bm25.py).
//
// Integer/float streaming work over posting lists, no tensor cores; HBM sees every posting range once,
// the rest is served from L2, so the binding resource is the SM side: load/store-unit cycles (one
// global load pair + one shared-memory atomic per posting) and issue slots.
//
// Work decomposition.  A *chain* is one (query, doc split) pair; it walks its split in *items* of
// `item_slabs` slabs (a slab = `slab` docs of int32 fixed-point accumulators in shared memory, as many
// as two resident CTAs per SM allow: 22 528 docs = 88 KB for k <= 256).  Items are
// numbered step-major -- item = step * n_chains + chain -- and handed out by an atomic counter to
// persistent CTAs, so at any moment the whole machine works on the same narrow doc range of every
// query: a posting list is fetched from HBM once per range and then served from L2 to all the other
// queries that share the term.  A chain's state between two of its items (term cursors, running
// threshold, surviving candidates) lives in the workspace; item (step, chain) waits on a flag that
// item (step - 1, chain) -- always a lower item number, so already claimed by a running CTA -- sets.
//
// The CTA is warp-specialised so that the arithmetic never does bookkeeping:
//   * bounds warp   -- claims the next item, loads the chain's term cursors and finds the posting
//                      boundaries of every query term at every slab edge with one parallel round of
//                      windowed binary searches (postings are doc-id sorted); double-buffered, one
//                      group of slabs ahead of the emit warp;
//   * emit warp     -- publishes one small *slab table* per slab in a shared-memory ring: per query term the run of
//                      postings that falls into the slab {first posting, count, multiplicity x scale} and the
//                      prefix of the runs' chunk counts (a chunk = 256 postings starting at a 128-byte boundary of
//                      the posting arrays), plus flags {end of slab | item begin | item end};
//   * 16 consumer warps -- each reads the table and takes the chunks c = warp, warp + 16, ... of the slab: every warp
//                      makes the same number of trips to memory (+-1) and nobody claims anything.  A chunk's
//                      (doc, impact) pairs are loaded straight from global memory / L2 into registers
//                      (coalesced 4-byte loads: no shared-memory staging, no per-stage
//                      hand-shake between warps) and added into acc[doc - slab0] with one
//                      shared-memory integer atomic per posting: scores are kept in fixed point (int32, a
//                      per-query power-of-two scale sized from `impact_bound`), so adds commute exactly, the
//                      chunks of a slab need no ordering between them and the result is independent
//                      of scheduling.  With non-negative impacts a doc's running score only grows: the add that
//                      brings it to the query's running k-th best sees that in the atomic's return value and
//                      notes the doc in a short "hot" list.  At the end of the slab the warps meet at a barrier and
//                      only the hot docs are looked at (the slab is scanned in full only while there is no
//                      threshold yet, when the list overflows, or when impacts can be negative; an overflow of the
//                      candidate buffer triggers an exact radix select that raises the threshold).  The slab is
//                      re-zeroed.  Item-begin / item-end tables load / store the chain's candidate state; the
//                      chain's last item sorts its top-k into the per-chain key list.
// Round 1 staged postings through a bulk-copy ring that all 16 warps consumed stage by stage: 44 warp
// instructions and 10 shared-memory wavefronts per 32 postings (ring write + ring read + atomic + the
// per-stage hand-shake).  The direct loads need ~12 instructions and ~6 load/store-unit cycles
// (tools/micro/bm25_stream.cu: 1.2-1.5 T postings/s against 0.44-0.71 T for the ring).
// Slabs in which no query term has a posting are skipped when impacts are known non-negative;
// documents that match nothing (score 0) are then added by the merge step, lowest id first, exactly
// as the reference's stable sort does.  bm25_merge_kernel merges the per-split lists of a query.
#include <algorithm>
#include <climits>
#include <cstdlib>

#include "common.cuh"
#include "select.cuh"

namespace lrag {

#ifndef LRAG_BM25_CONSUMERS
#define LRAG_BM25_CONSUMERS 512
#endif
constexpr int BM25_CONSUMERS = LRAG_BM25_CONSUMERS;  // consumer threads
constexpr int BM25_THREADS = BM25_CONSUMERS + 64;    // consumers, emit warp, bounds warp
constexpr int BM25_SLAB_STEP = 4 * BM25_CONSUMERS;   // slab sizes are multiples of one int4 per consumer thread
constexpr int BM25_SLAB_MAX = 24576 / BM25_SLAB_STEP * BM25_SLAB_STEP;   // <= 24 576 docs = 96 KB
constexpr int BM25_MAX_GROUP = 16;                   // slabs per bounds group (fewer when a query has many terms)
constexpr int BM25_BOUND_CAP = 17 * 32;              // ints per bounds buffer: (group + 1) * nt must fit
constexpr int BM25_MAXT = LRAG_BM25_MAX_QUERY_TERMS;
#ifndef LRAG_BM25_SEG
#define LRAG_BM25_SEG 256
#endif
constexpr int BM25_SEG = LRAG_BM25_SEG;              // postings per segment descriptor (8 per lane)
constexpr int BM25_SEG_PER_LANE = BM25_SEG / 32;
constexpr int BM25_TABS = 4;                         // slab tables in flight
constexpr int BM25_WARPS = BM25_CONSUMERS / 32;
constexpr int BM25_HOT = 256;                        // docs per slab whose running score reached the threshold (more: full scan)
constexpr int BM25_BAR_CONSUMERS = 1;                // named barrier id of the consumer warps
constexpr int BM25_KEEP = 2 * LRAG_MAX_K / BM25_CONSUMERS;   // candidate keys a thread may hold across a compaction
constexpr int BM25_DEFAULT_ITEM_DOCS = 32 * 12288;   // docs per work item (the item is a whole number of slabs)
constexpr int BM25_DENSE_FLAG = 1 << 30;             // in a slab table's run count: the run is a doc range of a dense row
enum : int { BM25_F_SLAB_END = 1, BM25_F_ITEM_BEGIN = 2, BM25_F_ITEM_END = 4, BM25_F_FINAL = 8, BM25_F_END = 16, BM25_F_SLAB_BEGIN = 32 };

#ifdef LRAG_BM25_TIMING
// Debug build only (tools/gpu_bm25_timing.py): warp-cycles the consumer warps spend per phase, summed over the grid.
__device__ unsigned long long g_bm25_t[16];    // table wait, slab init, chunk loop, slab end, item begin/end, tables, chunks, slabs
__device__ long long g_bm25_trace[1024 * 16 * 4];   // block 0, first 1024 slabs, per consumer warp: chunk loop begin / end, chunks, barrier passed
#define BM25_T(i, expr) do { const long long t_ = clock64(); expr; tacc[i] += clock64() - t_; } while (0)
#else
#define BM25_T(i, expr) do { expr; } while (0)
#endif

struct Bm25Ws {
  unsigned long long* counter;     // next item
  int* cur_flag;                   // [nc] steps whose final term cursors are published
  int* cand_flag;                  // [nc] steps whose candidate state is published
  int* q_nt;                       // [nq] distinct in-vocabulary terms with postings
  int64_t* tq_start;               // [nq, TS] first posting; -(r + 1) = dense row r
  int32_t* tq_len;                 // [nq, TS] df
  float* tq_mult;                  // [nq, TS] occurrences in the query x the query's fixed-point scale
  float* q_inv_scale;              // [nq] 1 / scale
  int32_t* ch_cur;                 // [nc, TS] postings consumed so far (relative to tq_start)
  unsigned long long* ch_thr;      // [nc]
  int* ch_cnt;                     // [nc]
  uint64_t* ch_cand;               // [nc, cap]
  uint64_t* out_keys;              // [nc, k] = [nq, S, k]
};

struct Bm25Params {
  const int64_t* indptr; const int32_t* doc_id; const float* impact; int64_t V; int64_t nnz;
  const int64_t* q_indptr; const int32_t* q_term;
  // Dense rows: for the few terms that occur in most documents, impact[d] for EVERY doc of the shard (0 where the term is
  // absent), n_dense rows of dense_stride floats (a multiple of 32).  A dense term's "posting list" is the doc range
  // itself: half the bytes of (doc, impact) pairs at density > 0.5, no doc ids to load, and consecutive lanes add into
  // consecutive accumulators (no bank conflicts).
  const int32_t* dense_term; const float* dense_rows; int n_dense; int64_t dense_stride;
  int64_t N; int64_t dps;          // docs per split (multiple of the item size)
  unsigned long long total_items;
  int nq, k, nonneg, S, cap, P, TS, item_slabs, steps, nc;
  int slab;                        // docs per slab (multiple of BM25_SLAB_STEP)
  float impact_bound;              // >= max |impact| over the index
  unsigned long long* start_counter;   // every CTA counts itself in when it starts (partition.cu), or null
  Bm25Ws ws;
};

struct Bm25Group {                 // bounds warp -> emit warp
  int kind;                        // 0 = slabs of an item, 1 = no more work
  int chain, step, first, last, final_step, nt, ns;
  int b0, range_end;
  float inv_scale;
  int64_t t_start[BM25_MAXT];
  float t_mult[BM25_MAXT];
  int32_t bound[BM25_BOUND_CAP];   // [t * (ns + 1) + j]
  int32_t last_t[BM25_MAX_GROUP];  // last term with postings in the slab, -1 = none
};

// One slab's work (<= 32 runs; a query with more terms takes several tables per slab) or a bare marker, written by
// the emit warp.  `tag` (lap + 1) is written last and publishes the table; every consumer warp adds one to `arrived`
// once it holds the table in registers, and the slot is rewritten when all have.
struct alignas(16) Bm25Tab {
  int flags, slab0, nruns, nchunks;
  int chain, step; float inv_scale; int tag;
  int arrived, pad[3];
  int first_lo[32], first_hi[32], cnt[32], cpre[32];     // per run: first posting, count, chunks of the runs before it
  float mult[32];                                         // occurrences in the query x the query's fixed-point scale
};

struct Bm25Shared {
  SelectShared sel;
  Bm25Tab tab[BM25_TABS];
  int hot[2][BM25_HOT];            // docs of the current slab whose running score reached the threshold (buffer = slab parity)
  int hot_cnt[2];
  int drained;                     // last slab whose hot docs warp 0 has read out of the accumulators
  uint64_t bfull_bar[2], bempty_bar[2];                      // group buffers
  Bm25Group grp[2];
  int64_t t_start[BM25_MAXT];      // bounds warp's view of the current item's query
  int32_t t_len[BM25_MAXT];
  int32_t t_cur[BM25_MAXT];
  int cand_cnt;
  unsigned long long thr_key;
  float inv_scale;                 // fixed point -> fp32 score of the current item's query
  int range_end;                   // docs at or past the end of the chain's split are not ranked
};
constexpr int BM25_SH_BYTES = (int(sizeof(Bm25Shared)) + 127) / 128 * 128;   // the slab follows the control block

__device__ __forceinline__ int lower_bound_doc(const int32_t* __restrict__ ids, int lo, int hi, int64_t target) {
  // first index in [lo, hi) whose doc id >= target
  while (lo < hi) {
    const int mid = lo + ((hi - lo) >> 1);
    if (int64_t(__ldg(ids + mid)) < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Streaming 4-byte loads of postings.  Every posting is used once per CTA but by many CTAs at about the same time, so it
// must stay in L2 and has no business in L1: ld.global.cg.  (Measured at the hybrid shape: `.nc.L1::no_allocate` reads
// every posting ~10 times from DRAM -- 27 GB against 2.2 GB, L2 hit rate 82 % against 97 % -- and is 9 % slower.)
__device__ __forceinline__ int ldg_stream_s32(const int32_t* p) {
  int v; asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}
__device__ __forceinline__ float ldg_stream_f32(const float* p) {
  float v; asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
}
// shared-memory words that another warp of the CTA publishes
__device__ __forceinline__ int lds_volatile(const int* p) {
  int v; asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory"); return v;
}
__device__ __forceinline__ void sts_volatile(int* p, int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// Bounded spin on a shared-memory word until it equals / reaches a value.  The waiter (the emit warp, tables ahead of the
// consumers) sleeps between polls: a poll every few hundred nanoseconds is plenty for a slot that frees once per slab (several
// microseconds), and polling loops were 11 % of the kernel's executed instructions before (ncu, round 2).
#ifndef LRAG_BM25_POLL_NS
#define LRAG_BM25_POLL_NS 256
#endif
__device__ __forceinline__ void spin_shared(const int* p, int want, bool at_least) {
  int v = lds_volatile(p);
  if (at_least ? v >= want : v == want) return;
  long long t0 = 0;
  for (int it = 0;; ++it) {
    __nanosleep(LRAG_BM25_POLL_NS);
    v = lds_volatile(p);
    if (at_least ? v >= want : v == want) return;
    if ((it & 1023) == 0) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 20000000000LL) {
        printf("lrag: bm25 shared-memory wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
        __trap();
      }
    }
  }
}

// mbarrier wait of the two bookkeeping warps: they wait long and often, so they sleep between polls and leave the
// issue slots to the consumers
__device__ __forceinline__ void mbar_wait_lazy(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_now(bar, parity)) return;
  long long t0 = 0;
  for (int it = 0;; ++it) {
    __nanosleep(2 * LRAG_BM25_POLL_NS);
    if (mbar_try_wait_now(bar, parity)) return;
    if ((it & 1023) == 0) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 20000000000LL) {
        printf("lrag: bm25 group-buffer wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
        __trap();
      }
    }
  }
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

// shared-memory accesses by 32-bit shared address (one address computation per access)
__device__ __forceinline__ int lds_s32(uint32_t a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ int atoms_add_s32(uint32_t a, int v) {
  int o;
  asm volatile("atom.shared.add.s32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory");
  return o;
}

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Bounded spin on a chain flag: a protocol bug traps instead of hanging the GPU box.
__device__ __forceinline__ void spin_until_ge(const int* flag, int want) {
  if (ld_acquire(flag) >= want) return;
  const long long t0 = clock64();
  while (ld_acquire(flag) < want) {
    __nanosleep(64);
    if (clock64() - t0 > 20000000000LL) {
      printf("lrag: bm25 chain flag wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
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
  const int* acc; float inv_scale; int64_t slab0; int64_t range_end; unsigned long long thr_key; int slab;
  template <class F> __device__ void operator()(F&& f) const {
    for (int i = threadIdx.x; i < ncand; i += BM25_CONSUMERS) f(cand[i]);
    for (int i = threadIdx.x; i < slab; i += BM25_CONSUMERS) {
      const int64_t doc = slab0 + i;
      if (doc >= range_end) break;
      const uint64_t key = make_key(float(acc[i]) * inv_scale, uint32_t(doc));
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

// ------------------------------------------------------------------------------------------------
// Per launch, before the scan: drops OOV terms, merges repeated query terms into a multiplicity
// (first-occurrence order) and resets the item counter and the chain flags.  One warp per query.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bm25_prepare_kernel(const Bm25Params p) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int gsz = gridDim.x * blockDim.x;
  for (int i = gtid; i < p.nc; i += gsz) { p.ws.cur_flag[i] = 0; p.ws.cand_flag[i] = 0; }
  if (gtid == 0) *p.ws.counter = 0ull;
  const int lane = threadIdx.x & 31;
  for (int q = gtid >> 5; q < p.nq; q += gsz >> 5) {
    const int64_t qs = p.q_indptr[q];
    const int64_t qlen = p.q_indptr[q + 1] - qs;
    const int nraw = int(qlen < p.TS ? qlen : p.TS);
    const int32_t* raw = p.q_term + qs;
    int base = 0, scored = 0;
    for (int r0 = 0; r0 < nraw; r0 += 32) {
      const int i = r0 + lane;
      int t = -1;
      if (i < nraw) { t = raw[i]; if (t < 0 || int64_t(t) >= p.V) t = -1; }
      bool own = t >= 0;
      for (int j = 0; j < i && own; ++j) own = (raw[j] != t);
      int64_t s = 0, e = 0;
      if (own) { s = p.indptr[t]; e = p.indptr[t + 1]; own = e > s; }      // a term without postings scores nothing
      int mult = 0;
      if (own) for (int j = i; j < nraw; ++j) mult += (raw[j] == t);
      const uint32_t m = __ballot_sync(0xffffffffu, own);
      if (own) {
        const size_t slot = size_t(q) * p.TS + base + __popc(m & ((1u << lane) - 1));
        // a term with a dense row is walked by doc range instead of by postings
        int lo_r = 0, hi_r = p.n_dense;
        while (lo_r < hi_r) { const int mid = (lo_r + hi_r) >> 1; if (p.dense_term[mid] < t) lo_r = mid + 1; else hi_r = mid; }
        const bool dense = lo_r < p.n_dense && p.dense_term[lo_r] == t;
        p.ws.tq_start[slot] = dense ? -int64_t(lo_r + 1) : s;
        p.ws.tq_len[slot] = dense ? int32_t(p.N) : int32_t(e - s);
        p.ws.tq_mult[slot] = float(mult);
      }
      base += __popc(m);
      scored += __reduce_add_sync(0xffffffffu, mult);
    }
    // Fixed-point scale of this query's accumulators: a power of two such that no score can reach
    // 2^30 in magnitude (|score| <= impact_bound x scored tokens).
    const float bound = fmaxf(p.impact_bound * float(scored > 0 ? scored : 1), 1e-30f);
    int ex = 29 - ilogbf(bound);                    // bound * 2^ex < 2^30
    ex = ex > 60 ? 60 : (ex < -60 ? -60 : ex);
    const float scale = ldexpf(1.0f, ex);
    __syncwarp();
    if (p.n_dense > 0 && base > 1) {
      // terms with a dense row go first (stable partition): every slab's first table then holds all of them
      constexpr int R = BM25_MAXT / 32;
      int64_t ss[R]; int32_t ll[R]; float mm[R]; int pos_d[R]; bool dd[R], vv[R];
      int nd = 0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int i = r * 32 + lane;
        vv[r] = i < base; ss[r] = 0; ll[r] = 0; mm[r] = 0.f;
        if (vv[r]) { const size_t slot = size_t(q) * p.TS + i; ss[r] = p.ws.tq_start[slot]; ll[r] = p.ws.tq_len[slot]; mm[r] = p.ws.tq_mult[slot]; }
        dd[r] = vv[r] && ss[r] < 0;
        const uint32_t m = __ballot_sync(0xffffffffu, dd[r]);
        pos_d[r] = nd + __popc(m & ((1u << lane) - 1));
        nd += __popc(m);
      }
      __syncwarp();
      int ns = 0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const bool sp = vv[r] && !dd[r];
        const uint32_t m = __ballot_sync(0xffffffffu, sp);
        if (vv[r]) {
          const size_t slot = size_t(q) * p.TS + (dd[r] ? pos_d[r] : nd + ns + __popc(m & ((1u << lane) - 1)));
          p.ws.tq_start[slot] = ss[r]; p.ws.tq_len[slot] = ll[r]; p.ws.tq_mult[slot] = mm[r];
        }
        ns += __popc(m);
      }
      __syncwarp();
    }
    for (int t = lane; t < base; t += 32) p.ws.tq_mult[size_t(q) * p.TS + t] *= scale;
    if (lane == 0) { p.ws.q_nt[q] = base; p.ws.q_inv_scale[q] = ldexpf(1.0f, -ex); }
  }
}

// ------------------------------------------------------------------------------------------------
// What the consumer warps do at a marker.  Kept out of line: the chunk loop holds sixteen loaded postings in
// registers, and this code (barriers, candidate selection, chain state) must not take part in that loop's register
// allocation.  Each returns the fixed-point threshold at or above which an add notes its doc (INT_MAX: never).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int bm25_note_threshold(const Bm25Params& p, const Bm25Shared& sh) {
  if (!p.nonneg) return INT_MAX;             // negative impacts: every slab is scanned in full anyway
  const unsigned long long thr_key = sh.thr_key;
  // "strictly positive" while fewer than k docs have been seen (zero scores are filled in by the merge)
  const float thr_s = (uint32_t(thr_key) == 0xffffffffu && key_score(thr_key) == 0.f) ? 1.4e-45f : key_score(thr_key);
  const float t = thr_s / sh.inv_scale;      // a lower bound of every fixed-point value whose fp32 score reaches thr_s
  const int thr_i = t >= 2147483520.f ? INT_MAX : int(floorf(t - fabsf(t) * 2.4e-7f)) - 1;
  return thr_i < 1 ? 1 : thr_i;
}

// a doc whose running score reached the threshold (rare); false = the list is full, the slab will be scanned in full
__device__ __forceinline__ bool bm25_note_hot(Bm25Shared& sh, int hb, int doc) {
  const uint32_t at = uint32_t(atomicAdd(&sh.hot_cnt[hb], 1));
  if (at < uint32_t(BM25_HOT)) { sh.hot[hb][at] = doc; return true; }
  return false;
}

__device__ __forceinline__ float4 ldg_stream_f32x4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// One pass of a slab's initialisation over one dense row: acc[d] (+)= rn(row[d] * ms).  A thread owns the int4 groups
// tid, tid + 512, ...: plain 16-byte stores, no atomics; BM25_INIT_BATCH row loads are in flight per thread.  `lim` =
// floats of the row from the slab's first doc on (multiple of 4).  The last pass checks the sums against the note threshold.
#ifndef LRAG_BM25_INIT_BATCH
#define LRAG_BM25_INIT_BATCH 4
#endif
constexpr int BM25_INIT_BATCH = LRAG_BM25_INIT_BATCH;
template <bool FIRST>
__device__ __forceinline__ bool bm25_init_pass(Bm25Shared& sh, int4* a4, int iters, const float* row, float ms, int lim, int slab0,
                                               int thr_i, bool last, int hb) {
  const int tid = threadIdx.x;
  bool noted = false;
  for (int i0 = 0; i0 < iters; i0 += BM25_INIT_BATCH) {
    float4 v[BM25_INIT_BATCH];
#pragma unroll
    for (int u = 0; u < BM25_INIT_BATCH; ++u) {
      const int off = 4 * (tid + (i0 + u) * BM25_CONSUMERS);
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i0 + u < iters && off < lim) v[u] = ldg_stream_f32x4(row + off);
    }
#pragma unroll
    for (int u = 0; u < BM25_INIT_BATCH; ++u) {
      if (i0 + u < iters) {
        const int g = tid + (i0 + u) * BM25_CONSUMERS;
        int4 s = make_int4(0, 0, 0, 0);
        if (!FIRST) s = a4[g];
        s.x += __float2int_rn(v[u].x * ms); s.y += __float2int_rn(v[u].y * ms);
        s.z += __float2int_rn(v[u].z * ms); s.w += __float2int_rn(v[u].w * ms);
        a4[g] = s;
        if (last && max(max(s.x, s.y), max(s.z, s.w)) >= thr_i) {
          noted = true;
          const int doc = slab0 + 4 * g;
          if (s.x >= thr_i) bm25_note_hot(sh, hb, doc);
          if (s.y >= thr_i) bm25_note_hot(sh, hb, doc + 1);
          if (s.z >= thr_i) bm25_note_hot(sh, hb, doc + 2);
          if (s.w >= thr_i) bm25_note_hot(sh, hb, doc + 3);
        }
      }
    }
  }
  return noted;
}

// Begin of a slab: the accumulators are set to the summed contributions of the query's dense-row terms (zero without any)
// -- this replaces both the re-zeroing of the slab and every shared-memory atomic of those terms.  `dm` = lanes of the slab's
// first table that hold a dense run; r_lo / r_hi = the lane's run start (element of dense_rows at the slab's first doc),
// r_mult = multiplicity x scale.  Returns whether this thread noted a hot doc.  Ends behind a barrier.
__device__ __noinline__ bool bm25_slab_init(const Bm25Params& p, Bm25Shared& sh, int slab0, uint32_t dm, uint32_t r_lo, int r_hi,
                                            float r_mult, int thr_i, int hb) {
  const int tid = threadIdx.x;
  const NamedBarrier cbar{BM25_BAR_CONSUMERS, BM25_CONSUMERS};
  int4* a4 = reinterpret_cast<int4*>(reinterpret_cast<uint8_t*>(&sh) + BM25_SH_BYTES);
  const int iters = p.slab / BM25_SLAB_STEP;
  bool noted = false;
  if (dm == 0) {
#pragma unroll 4
    for (int i = 0; i < iters; ++i) a4[tid + i * BM25_CONSUMERS] = make_int4(0, 0, 0, 0);
  } else {
    const int64_t left = p.dense_stride - int64_t(slab0);
    const int lim = left < int64_t(p.slab) ? int(left) : p.slab;
    bool first = true;
    while (dm) {
      const int t = __ffs(dm) - 1; dm &= dm - 1;
      const float* row = p.dense_rows + (int64_t(__shfl_sync(0xffffffffu, r_lo, t)) | (int64_t(__shfl_sync(0xffffffffu, r_hi, t)) << 32));
      const float ms = __shfl_sync(0xffffffffu, r_mult, t);
      const bool last = dm == 0;
      noted |= first ? bm25_init_pass<true>(sh, a4, iters, row, ms, lim, slab0, thr_i, last, hb)
                     : bm25_init_pass<false>(sh, a4, iters, row, ms, lim, slab0, thr_i, last, hb);
      first = false;
    }
  }
  cbar();                                                 // the slab is initialised before anybody adds into it
  return noted;
}

// End of a slab: ranks the slab's hot docs (or the whole slab).  The accumulators are left as they are: the next slab's
// initialisation overwrites them.  `noted` = this thread noted a hot doc in the slab; the verdict "somebody did" travels with
// the barrier itself.  Usual hot case (a few docs reached the threshold): warp 0 alone reads their final scores and appends
// them to the candidate buffer; the other warps only wait for the read-out (a flag), not for the appends -- the threshold
// does not move until the buffer overflows.  The hot lists are double-buffered by slab parity, and the candidate count is
// only changed by warp 0 before it arrives at the next barrier, so every thread sees the same counts behind the barrier.
// `seq` = slabs this CTA has ended so far.
#ifdef LRAG_BM25_TIMING
#define BM25_PATH(x) (*path_out = (x))
__device__ __noinline__ int bm25_slab_end(const Bm25Params& p, Bm25Shared& sh, int slab0_, int thr_note, bool noted, int seq, int* path_out) {
#else
#define BM25_PATH(x) ((void)0)
__device__ __noinline__ int bm25_slab_end(const Bm25Params& p, Bm25Shared& sh, int slab0_, int thr_note, bool noted, int seq) {
#endif
  const int tid = threadIdx.x;
  const NamedBarrier cbar{BM25_BAR_CONSUMERS, BM25_CONSUMERS};
  const int SLAB = p.slab, cap = p.cap;
  int* acci = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(&sh) + BM25_SH_BYTES);
  uint64_t* cand = reinterpret_cast<uint64_t*>(acci + SLAB);
  const int64_t slab0 = slab0_;
  const bool any = consumers_bar_or(noted);               // every add of the slab is done
  // nothing reached the threshold (the usual case once a few slabs have been seen): keep the threshold
  BM25_PATH(0);
  if (p.nonneg && !any) return thr_note;
  const int hb = seq & 1;
  const int nh = sh.hot_cnt[hb];
  const int cnt_before = sh.cand_cnt;
  const bool listed = p.nonneg && nh <= BM25_HOT;         // the hot list holds every doc that may have reached the threshold
  const bool fast = listed && cnt_before + nh <= cap;     // ... and the candidate buffer has room for all of them
  BM25_PATH(fast ? 1 : 2);
  if (fast && tid >= 32) {
    // ---- the other warps: the hot docs' accumulators must have been read before the next slab overwrites them ----
    if (lds_volatile(&sh.drained) != seq) {
      const long long t0 = clock64();
      while (lds_volatile(&sh.drained) != seq) {
#ifdef LRAG_BM25_DRAIN_SLEEP_NS
        __nanosleep(LRAG_BM25_DRAIN_SLEEP_NS);
#endif
        if (clock64() - t0 > 20000000000LL) { printf("lrag: bm25 hot-list wait timed out (block %d thread %d)\n", blockIdx.x, tid); __trap(); }
      }
    }
    return thr_note;
  }
  const float inv_scale = sh.inv_scale;
  const int64_t range_end = sh.range_end;
  // warp 0: final score of every hot doc -> candidate (a doc noted twice is taken once: the read clears it).  With
  // `publish`, the flag the other warps wait for is raised as soon as the last read-out has returned, before the appends.
  auto drain = [&](unsigned long long thr_key, float thr_s, int thr_i, bool publish) {
    for (int i0 = 0; i0 < nh; i0 += 32) {
      const int i = i0 + tid;
      int doc = 0, val = 0;
      if (i < nh) { doc = sh.hot[hb][i]; val = atomicExch(acci + (doc - int(slab0)), 0); }
      if (publish && i0 + 32 >= nh) {
        const int dep = __reduce_or_sync(0xffffffffu, val);       // every read-out has returned
        if (tid == 0) {
          sh.hot_cnt[hb] = 0;
          asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(&sh.drained)), "r"(seq), "r"(dep) : "memory");
        }
      }
      const float sc = float(val) * inv_scale;
      bool want = i < nh && val >= thr_i && sc >= thr_s && int64_t(doc) < range_end;
      uint64_t key = 0;
      if (want) { key = make_key(sc, uint32_t(doc)); want = key > thr_key; }
      cand_append(want, key, cand, cap, &sh.cand_cnt);
    }
  };
  if (fast) {
    // nonneg: thr_key is never 0 and the note threshold is the fixed-point bound of the current threshold score
    const unsigned long long thr_key = sh.thr_key;
    const float thr_s = (uint32_t(thr_key) == 0xffffffffu && key_score(thr_key) == 0.f) ? 1.4e-45f : key_score(thr_key);
    drain(thr_key, thr_s, thr_note, true);
    return thr_note;
  }
  if (listed) {
    // ---- the candidate buffer is full: its exact k-th best becomes the threshold and only the k best stay (the slab is
    //      not scanned: every doc of it that can matter is on the hot list) ----
    BM25_PATH(3);
    Bm25Cands cands{cand, cnt_before};
    const unsigned long long pivot = block_select_pivot(cands, p.k, sh.sel, cbar);
    uint64_t keep[BM25_KEEP];       // cap <= 2 * LRAG_MAX_K: at most BM25_KEEP old keys per thread
#pragma unroll
    for (int i = 0; i < BM25_KEEP; ++i) {
      const int idx = tid + i * BM25_CONSUMERS;
      keep[i] = idx < cnt_before ? cand[idx] : 0ull;
    }
    cbar();
    if (tid == 0) { sh.cand_cnt = 0; if (pivot > sh.thr_key) sh.thr_key = pivot; }
    cbar();
#pragma unroll
    for (int i = 0; i < BM25_KEEP; ++i) {
      const bool want = keep[i] != 0ull && keep[i] >= pivot;
      cand_append(want, keep[i], cand, cap, &sh.cand_cnt);
    }
    cbar();
    if (tid < 32) {
      const unsigned long long thr_key = sh.thr_key;
      const float thr_s = (uint32_t(thr_key) == 0xffffffffu && key_score(thr_key) == 0.f) ? 1.4e-45f : key_score(thr_key);
      drain(thr_key, thr_s, bm25_note_threshold(p, sh), false);
    }
  } else {
    const unsigned long long thr_key = sh.thr_key;
    // initial thresholds: nothing yet (-inf), or "strictly positive" when zero scores are filled in later
    const float thr_s = !thr_key ? -INFINITY
                        : (uint32_t(thr_key) == 0xffffffffu && key_score(thr_key) == 0.f) ? 1.4e-45f : key_score(thr_key);
    // thr_i is a lower bound of every fixed-point value whose fp32 score reaches thr_s
    int thr_i = INT_MIN;
    if (thr_s > -INFINITY) {
      const float t = thr_s / inv_scale;                     // inv_scale is a power of two: exact
      thr_i = t >= 2147483520.f ? INT_MAX : (t <= -2147483520.f ? INT_MIN : int(floorf(t - fabsf(t) * 2.4e-7f)) - 1);
      if (p.nonneg && thr_i < 1) thr_i = 1;                  // zero scores are never candidates then
    }
    // ---- scan the slab for candidates ----
    const int4* a4 = reinterpret_cast<const int4*>(acci);
#pragma unroll 2
    for (int i = 0; i < SLAB / BM25_SLAB_STEP; ++i) {
      const int idx = (tid + i * BM25_CONSUMERS) * 4;
      const int4 s4 = a4[tid + i * BM25_CONSUMERS];
      const bool any4 = max(max(s4.x, s4.y), max(s4.z, s4.w)) >= thr_i;
      if (!__any_sync(0xffffffffu, any4)) continue;
      const float sv[4] = {float(s4.x) * inv_scale, float(s4.y) * inv_scale, float(s4.z) * inv_scale, float(s4.w) * inv_scale};
#pragma unroll
      for (int el = 0; el < 4; ++el) {
        const int64_t doc = slab0 + idx + el;
        bool want = (sv[el] >= thr_s) && (doc < range_end);
        uint64_t key = 0;
        if (want) { key = make_key(sv[el], uint32_t(doc)); want = key > thr_key; }
        cand_append(want, key, cand, cap, &sh.cand_cnt);
      }
    }
    cbar();
    if (sh.cand_cnt > cap) {
      BM25_PATH(3);
      // ---- overflow: exact k-th best of (buffer U slab) becomes the new threshold ----
      Bm25Union uni{cand, cnt_before, acci, inv_scale, slab0, range_end, thr_key, SLAB};
      const unsigned long long pivot = block_select_pivot(uni, p.k, sh.sel, cbar);
      uint64_t keep[BM25_KEEP];     // cap <= 2 * LRAG_MAX_K: at most BM25_KEEP old keys per thread
#pragma unroll
      for (int i = 0; i < BM25_KEEP; ++i) {
        const int idx = tid + i * BM25_CONSUMERS;
        keep[i] = idx < cnt_before ? cand[idx] : 0ull;
      }
      cbar();
      if (tid == 0) sh.cand_cnt = 0;
      cbar();
#pragma unroll
      for (int i = 0; i < BM25_KEEP; ++i) {
        const bool want = keep[i] != 0ull && keep[i] >= pivot;
        cand_append(want, keep[i], cand, cap, &sh.cand_cnt);
      }
      for (int i = tid; i < SLAB; i += BM25_CONSUMERS) {
        const int64_t doc = slab0 + i;
        uint64_t key = 0;
        bool want = doc < range_end;
        if (want) { key = make_key(float(acci[i]) * inv_scale, uint32_t(doc)); want = (key > thr_key) && (key >= pivot); }
        cand_append(want, key, cand, cap, &sh.cand_cnt);
      }
      cbar();
      if (tid == 0 && pivot > sh.thr_key) sh.thr_key = pivot;
    }
  }
  if (tid == 0) sh.hot_cnt[hb] = 0;
  cbar();                                                 // the slab is read and the threshold set before the next slab begins
  return bm25_note_threshold(p, sh);
}

// Begin of an item: the chain's candidate state (from the workspace after the first step).
__device__ __noinline__ int bm25_item_begin(const Bm25Params& p, Bm25Shared& sh, int chain, int step, float inv_scale) {
  const int tid = threadIdx.x;
  const NamedBarrier cbar{BM25_BAR_CONSUMERS, BM25_CONSUMERS};
  const int cap = p.cap;
  int* acci = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(&sh) + BM25_SH_BYTES);
  uint64_t* cand = reinterpret_cast<uint64_t*>(acci + p.slab);
  // the previous item ended behind a barrier: the candidate buffer is free
  if (tid == 0) {
    sh.inv_scale = inv_scale;
    sh.range_end = int(min(p.N, int64_t(chain % p.S + 1) * p.dps));
  }
  if (step == 0) {
    if (tid == 0) { sh.cand_cnt = 0; sh.thr_key = p.nonneg ? ((uint64_t(ord32(0.0f)) << 32) | 0xffffffffull) : 0ull; }
  } else {
    spin_until_ge(p.ws.cand_flag + chain, step);
    const int cnt = __ldcg(p.ws.ch_cnt + chain);
    const uint64_t* src = p.ws.ch_cand + size_t(chain) * cap;
    for (int i = tid; i < cnt; i += BM25_CONSUMERS) cand[i] = __ldcg(src + i);
    if (tid == 0) { sh.cand_cnt = cnt; sh.thr_key = __ldcg(p.ws.ch_thr + chain); }
  }
  cbar();
  return bm25_note_threshold(p, sh);
}

// End of an item: the chain state goes to the workspace, or (last item of the chain) its sorted top-k to the key list.
__device__ __noinline__ void bm25_item_end(const Bm25Params& p, Bm25Shared& sh, int chain, int step, int final_step) {
  const int tid = threadIdx.x;
  const NamedBarrier cbar{BM25_BAR_CONSUMERS, BM25_CONSUMERS};
  const int cap = p.cap;
  int* acci = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(&sh) + BM25_SH_BYTES);
  uint64_t* cand = reinterpret_cast<uint64_t*>(acci + p.slab);
  cbar();
  if (final_step) {
    // ---- sorted top-k of the surviving candidates -> this chain's key list ----
    const int ncand = min(sh.cand_cnt, cap);
    Bm25Cands cands{cand, ncand};
    const unsigned long long pivot = block_select_pivot(cands, p.k, sh.sel, cbar);
    uint64_t* sortbuf = reinterpret_cast<uint64_t*>(acci);     // the slab (dead between items) doubles as the sort buffer
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
    uint64_t* out = p.ws.out_keys + size_t(chain) * p.k;
    for (int i = tid; i < p.k; i += BM25_CONSUMERS) out[i] = sortbuf[i];
    cbar();
  } else {
    const int cnt = min(sh.cand_cnt, cap);
    uint64_t* dst = p.ws.ch_cand + size_t(chain) * cap;
    for (int i = tid; i < cnt; i += BM25_CONSUMERS) dst[i] = cand[i];
    if (tid == 0) { p.ws.ch_cnt[chain] = cnt; p.ws.ch_thr[chain] = sh.thr_key; }
    __threadfence();
    cbar();
    if (tid == 0) st_release(p.ws.cand_flag + chain, step + 1);
  }
}

__global__ void __launch_bounds__(BM25_THREADS, 2)
bm25_scan_kernel(const __grid_constant__ Bm25Params p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int SLAB = p.slab;
  Bm25Shared& sh = *reinterpret_cast<Bm25Shared*>(smem_raw);                                // control block first: constant addresses
  float* acc = reinterpret_cast<float*>(smem_raw + BM25_SH_BYTES);                          // [SLAB], then the candidate keys [cap]

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    sh.cand_cnt = 0;
    sh.thr_key = 0ull;
    sh.hot_cnt[0] = sh.hot_cnt[1] = 0;
    sh.drained = -1;
    for (int s = 0; s < 2; ++s) { mbar_init(&sh.bfull_bar[s], 1); mbar_init(&sh.bempty_bar[s], 1); }
    fence_barrier_init();
  }
  if (tid < BM25_TABS) { sh.tab[tid].tag = 0; sh.tab[tid].arrived = BM25_WARPS; }
  if (tid == 0 && p.start_counter) atomicAdd(p.start_counter, 1ull);      // this CTA is resident
  __syncthreads();                 // (the accumulators are initialised at the beginning of every slab)
  if (warp == BM25_CONSUMERS / 32 + 1) {
    // ===================== bounds warp =====================
    const int64_t item_docs = int64_t(p.item_slabs) * SLAB;
    uint32_t gcount = 0;
    for (;;) {
      unsigned long long item = 0;
      if (lane == 0) item = atomicAdd(p.ws.counter, 1ull);
      item = __shfl_sync(0xffffffffu, item, 0);
      if (item >= p.total_items) break;
      const int step = int(item / (unsigned long long)p.nc);
      const int chain = int(item - (unsigned long long)step * p.nc);
      const int q = chain / p.S, split = chain - q * p.S;
      const int64_t split_begin = int64_t(split) * p.dps;
      const int64_t split_end = min(p.N, split_begin + p.dps);
      const int64_t rb = split_begin + int64_t(step) * item_docs;
      const int64_t re = min(split_end, rb + item_docs);
      const int nt = p.ws.q_nt[q];
      for (int t = lane; t < nt; t += 32) {
        sh.t_start[t] = p.ws.tq_start[size_t(q) * p.TS + t];
        sh.t_len[t] = p.ws.tq_len[size_t(q) * p.TS + t];
      }
      __syncwarp();
      if (step > 0) {
        spin_until_ge(p.ws.cur_flag + chain, step);
        for (int t = lane; t < nt; t += 32) sh.t_cur[t] = __ldcg(p.ws.ch_cur + size_t(chain) * p.TS + t);
      } else {
        for (int t = lane; t < nt; t += 32)
          sh.t_cur[t] = split_begin == 0 ? 0
                        : (sh.t_start[t] < 0 ? int(split_begin) : lower_bound_doc(p.doc_id + sh.t_start[t], 0, sh.t_len[t], split_begin));
      }
      __syncwarp();
      int gs = nt > 0 ? BM25_BOUND_CAP / nt - 1 : BM25_MAX_GROUP;
      gs = gs < 1 ? 1 : (gs > BM25_MAX_GROUP ? BM25_MAX_GROUP : gs);
      const int nslab = rb < re ? int((re - rb + SLAB - 1) / SLAB) : 0;
      const int ngroups = nslab > 0 ? (nslab + gs - 1) / gs : 1;
      for (int gi = 0; gi < ngroups; ++gi, ++gcount) {
        const uint32_t bb = gcount & 1;
        mbar_wait_lazy(&sh.bempty_bar[bb], ((gcount >> 1) & 1) ^ 1);
        Bm25Group& G = sh.grp[bb];
        const int ns = max(0, min(gs, nslab - gi * gs));
        const int64_t b0 = rb + int64_t(gi) * gs * SLAB;
        if (lane == 0) {
          G.kind = 0; G.chain = chain; G.step = step; G.first = (gi == 0); G.last = (gi == ngroups - 1);
          G.final_step = (step == p.steps - 1); G.nt = nt; G.ns = ns; G.b0 = int(b0); G.range_end = int(re);
          G.inv_scale = p.ws.q_inv_scale[q];
        }
        for (int t = lane; t < nt; t += 32) {
          G.t_start[t] = sh.t_start[t];
          G.t_mult[t] = p.ws.tq_mult[size_t(q) * p.TS + t];
        }
        if (ns > 0) {
          for (int w = lane; w < nt * (ns + 1); w += 32) {
            const int t = w / (ns + 1), j = w % (ns + 1);
            const int cur = sh.t_cur[t], len = sh.t_len[t];
            int64_t target = b0 + int64_t(j) * SLAB;
            if (target > re) target = re;
            // ids are strictly increasing inside a term: the answer is at most (target - b0) past cur
            const int64_t reach = int64_t(cur) + (target - b0);
            const int hi = int(reach < len ? reach : int64_t(len));
            G.bound[w] = (j == 0) ? cur
                         : (sh.t_start[t] < 0 ? int(target) : lower_bound_doc(p.doc_id + sh.t_start[t], cur, hi, target));   // dense row: doc offset
          }
          __syncwarp();
          if (lane < ns) {
            int last = -1;
            for (int t = 0; t < nt; ++t) if (G.bound[t * (ns + 1) + lane + 1] > G.bound[t * (ns + 1) + lane]) last = t;
            G.last_t[lane] = last;
          }
          for (int t = lane; t < nt; t += 32) sh.t_cur[t] = G.bound[t * (ns + 1) + ns];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.bfull_bar[bb]);
      }
      // publish the cursors for the chain's next item
      if (step + 1 < p.steps) {
        for (int t = lane; t < nt; t += 32) p.ws.ch_cur[size_t(chain) * p.TS + t] = sh.t_cur[t];
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release(p.ws.cur_flag + chain, step + 1);
      }
      __syncwarp();
    }
    const uint32_t bb = gcount & 1;
    mbar_wait_lazy(&sh.bempty_bar[bb], ((gcount >> 1) & 1) ^ 1);
    if (lane == 0) { sh.grp[bb].kind = 1; mbar_arrive(&sh.bfull_bar[bb]); }
  } else if (warp == BM25_CONSUMERS / 32) {
    // ===================== emit warp: (slab, term) runs -> slab tables =====================
    int wr = 0;             // tables written
    // next table slot, free once every consumer warp has taken its previous content
    auto acquire = [&]() -> Bm25Tab* {
      Bm25Tab* tb = &sh.tab[wr & (BM25_TABS - 1)];
      if (lane == 0) { spin_shared(&tb->arrived, BM25_WARPS, false); sts_volatile(&tb->arrived, 0); }
      __syncwarp();
      return tb;
    };
    auto publish = [&](Bm25Tab* tb, int flags, int slab0, int nruns, int nchunks, int chain, int step, float inv_scale) {
      __syncwarp();
      if (lane == 0) {
        tb->flags = flags; tb->slab0 = slab0; tb->nruns = nruns; tb->nchunks = nchunks;
        tb->chain = chain; tb->step = step; tb->inv_scale = inv_scale;
        asm volatile("" ::: "memory");      // shared-memory stores of one thread are performed in order
        sts_volatile(&tb->tag, wr / BM25_TABS + 1);
      }
      ++wr;
      __syncwarp();
    };
    for (uint32_t gc = 0;; ++gc) {
      const uint32_t bb = gc & 1;
      mbar_wait_lazy(&sh.bfull_bar[bb], (gc >> 1) & 1);
      const Bm25Group& G = sh.grp[bb];
      if (G.kind != 0) { publish(acquire(), BM25_F_END, 0, 0, 0, 0, 0, 0.f); break; }
      const int nt = G.nt, ns = G.ns, chain = G.chain, step = G.step;
      if (G.first) publish(acquire(), BM25_F_ITEM_BEGIN, 0, 0, 0, chain, step, G.inv_scale);
      for (int j = 0; j < ns; ++j) {
        const int sl0 = G.b0 + j * SLAB;
        // no postings here: nothing to do, unless impacts can be negative (a slab of zero scores is ranked then)
        if (G.last_t[j] < 0 && p.nonneg) continue;
        for (int t0 = 0; t0 < nt || t0 == 0; t0 += 32) {
          Bm25Tab* tb = acquire();
          const int t = t0 + lane;
          int64_t first = 0; int cnt = 0; float mult = 0.f;
          bool dense = false;
          if (t < nt) {
            const int lo = G.bound[t * (ns + 1) + j];
            cnt = G.bound[t * (ns + 1) + j + 1] - lo;
            const int64_t ts = G.t_start[t];
            dense = ts < 0;
            // dense row: no postings to walk; `first` = element of dense_rows at the slab's first doc (the slab initialisation reads it)
            if (dense) cnt = 0;
            first = dense ? (-ts - 1) * p.dense_stride + sl0 : ts + lo;
            mult = G.t_mult[t];
          }
          // chunks of a run: BM25_SEG postings each, counted from the 32-posting (128-byte) boundary at or below its first
          const int nch = cnt > 0 ? (int(first & 31) + cnt + BM25_SEG - 1) / BM25_SEG : 0;
          int incl = nch;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
          tb->first_lo[lane] = int(uint32_t(first)); tb->first_hi[lane] = int(uint32_t(first >> 32));
          tb->cnt[lane] = cnt | (dense ? BM25_DENSE_FLAG : 0); tb->cpre[lane] = incl - nch; tb->mult[lane] = mult;
          const int total = __shfl_sync(0xffffffffu, incl, 31);
          const bool last_tab = t0 + 32 >= nt;
          publish(tb, (last_tab ? BM25_F_SLAB_END : 0) | (t0 == 0 ? BM25_F_SLAB_BEGIN : 0), sl0, min(32, max(nt - t0, 0)), total, chain,
                  step, 0.f);
        }
      }
      if (G.last) publish(acquire(), BM25_F_ITEM_END | (G.final_step ? BM25_F_FINAL : 0), 0, 0, 0, chain, step, 0.f);
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.bempty_bar[bb]);
    }
  } else {
    // ===================== consumers =====================
    // Scores are accumulated in fixed point (int32, per-query power-of-two scale) with shared-memory
    // integer atomics: adds commute exactly, so the pieces of a slab need no ordering between them,
    // results do not depend on which warp ran first, and one ATOMS replaces a load / add / store.
    const uint32_t hot_u32 = smem_u32(sh.hot), hotc_u32 = smem_u32(sh.hot_cnt);
    int seq = 0;              // slabs ended so far; its parity selects the hot list
    const uint32_t acc_u32 = smem_u32(acc);
    int thr_i = INT_MAX;      // an add whose doc reaches this fixed-point score notes the doc
    int thr_slab = INT_MAX;   // thr_i at the start of the slab (a thread that saw the hot list overflow stops noting)

    // one posting: add; returns the doc's running score
    auto add_posting = [&](uint32_t accb, int doc, float imp, float ms) -> int {
      const int vi = __float2int_rn(imp * ms);
      return atoms_add_s32(accb + 4u * uint32_t(doc), vi) + vi;
    };
    bool noted = false;       // this thread noted a hot doc in the current slab
    // note a doc whose running score reached the threshold (rare)
    auto note_hot = [&](int doc) {
      uint32_t at;
      const uint32_t hb = uint32_t(seq & 1);
      asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(at) : "r"(hotc_u32 + 4u * hb) : "memory");
      noted = true;
      if (at < uint32_t(BM25_HOT)) asm volatile("st.shared.s32 [%0], %1;" ::"r"(hot_u32 + 4u * (hb * BM25_HOT + at)), "r"(doc) : "memory");
      else thr_i = INT_MAX;                                    // overflow: the slab will be scanned in full
    };

#ifdef LRAG_BM25_TIMING
    long long tacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif
    for (int rd = 0;; ++rd) {
      // ---- next slab table ----
      Bm25Tab* tb = &sh.tab[rd & (BM25_TABS - 1)];
#ifdef LRAG_BM25_TIMING
      const long long tw0 = clock64();
#endif
      {
        const int want = rd / BM25_TABS + 1;
        long long t0 = 0;
        while (lds_volatile(&tb->tag) != want) {
          __nanosleep(20);
          if (t0 == 0) t0 = clock64();
          else if (clock64() - t0 > 20000000000LL) {
            printf("lrag: bm25 table wait timed out (block %d warp %d table %d)\n", blockIdx.x, warp, rd);
            __trap();
          }
        }
      }
#ifdef LRAG_BM25_TIMING
      tacc[0] += clock64() - tw0; tacc[5] += 1;
#endif
      asm volatile("" ::: "memory");            // shared-memory loads of one thread are performed in order
      int4 h0, h1;
      asm volatile("ld.volatile.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(h0.x), "=r"(h0.y), "=r"(h0.z), "=r"(h0.w)
                   : "r"(smem_u32(&tb->flags)) : "memory");
      asm volatile("ld.volatile.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(h1.x), "=r"(h1.y), "=r"(h1.z), "=r"(h1.w)
                   : "r"(smem_u32(&tb->chain)) : "memory");
      const int flags = h0.x, slab0 = h0.y, nruns = h0.z, nchunks = h0.w;
      // this lane's run
      uint32_t r_lo = 0; int r_hi = 0, r_cnt = 0, r_cpre = INT_MAX; float r_mult = 0.f;
      if (lane < nruns) {
        r_lo = uint32_t(lds_volatile(&tb->first_lo[lane])); r_hi = lds_volatile(&tb->first_hi[lane]);
        r_cnt = lds_volatile(&tb->cnt[lane]); r_cpre = lds_volatile(&tb->cpre[lane]);
        r_mult = __int_as_float(lds_volatile(reinterpret_cast<const int*>(&tb->mult[lane])));
      }
      __syncwarp();
      if (lane == 0) atomicAdd(&tb->arrived, 1);           // the table is in registers: its slot may be rewritten
      if (flags & BM25_F_ITEM_BEGIN) BM25_T(4, thr_slab = thr_i = bm25_item_begin(p, sh, h1.x, h1.y, __int_as_float(h1.z)));
      // ---- first table of a slab: initialise the accumulators from the dense-row terms (or zero them) ----
      if (flags & BM25_F_SLAB_BEGIN) {
        const uint32_t dm = __ballot_sync(0xffffffffu, lane < nruns && (r_cnt & BM25_DENSE_FLAG) != 0);
        BM25_T(1, noted = bm25_slab_init(p, sh, slab0, dm, r_lo, r_hi, r_mult, thr_i, seq & 1));
      }
      // ---- this warp's chunks of the slab: c = warp, warp + 16, ...; 8 coalesced (doc, impact) loads per lane, then the adds ----
      const uint32_t accb = acc_u32 - 4u * uint32_t(slab0);        // accb + 4 * doc == &acc[doc - slab0]
#ifdef LRAG_BM25_TIMING
      const long long tc0 = clock64();
#endif
      for (int c = warp; c < nchunks; c += BM25_WARPS) {
#ifdef LRAG_BM25_TIMING
        tacc[6] += 1;
#endif
        // the run that holds chunk c: the last one whose chunk prefix is <= c
        const int t = __popc(__ballot_sync(0xffffffffu, r_cpre <= c)) - 1;
        const uint32_t f_lo = __shfl_sync(0xffffffffu, r_lo, t);
        const int f_hi = __shfl_sync(0xffffffffu, r_hi, t);
        const int cnt = __shfl_sync(0xffffffffu, r_cnt, t) & (BM25_DENSE_FLAG - 1);
        const int j = c - __shfl_sync(0xffffffffu, r_cpre, t);
        const float ms = __shfl_sync(0xffffffffu, r_mult, t);      // term multiplicity x scale
        const int64_t first = int64_t(f_lo) | (int64_t(f_hi) << 32);
        const int64_t base = (first & ~int64_t(31)) + int64_t(j) * BM25_SEG;
        const int lo = j == 0 ? int(f_lo & 31) : 0;                // slots [lo, hi) of the chunk belong to the run
        const int64_t left = first + cnt - base;
        const int hi = left < BM25_SEG ? int(left) : BM25_SEG;
        const int32_t* pid = p.doc_id + base + lane;
        const float* pim = p.impact + base + lane;
        int d[BM25_SEG_PER_LANE]; float v[BM25_SEG_PER_LANE];
        if (lo == 0 && hi == BM25_SEG) {
#pragma unroll
          for (int u = 0; u < BM25_SEG_PER_LANE; ++u) { d[u] = ldg_stream_s32(pid + 32 * u); v[u] = ldg_stream_f32(pim + 32 * u); }
          int tot[BM25_SEG_PER_LANE];
          int top = INT_MIN;
#pragma unroll
          for (int u = 0; u < BM25_SEG_PER_LANE; ++u) { tot[u] = add_posting(accb, d[u], v[u], ms); top = max(top, tot[u]); }
          if (top >= thr_i) {                                      // one branch per eight postings; taken rarely
#pragma unroll
            for (int u = 0; u < BM25_SEG_PER_LANE; ++u)
              if (tot[u] >= thr_i) note_hot(d[u]);
          }
        } else {
#pragma unroll
          for (int u = 0; u < BM25_SEG_PER_LANE; ++u) {
            d[u] = -1; v[u] = 0.f;
            const int sl = lane + 32 * u;
            if (sl >= lo && sl < hi) { d[u] = ldg_stream_s32(pid + 32 * u); v[u] = ldg_stream_f32(pim + 32 * u); }
          }
#pragma unroll
          for (int u = 0; u < BM25_SEG_PER_LANE; ++u)
            if (d[u] >= 0 && add_posting(accb, d[u], v[u], ms) >= thr_i) note_hot(d[u]);
        }
      }
#ifdef LRAG_BM25_TIMING
      const long long tc1 = clock64();
      tacc[2] += tc1 - tc0;
      if (flags & BM25_F_SLAB_END) tacc[7] += 1;
#endif
      // ---- markers: every consumer warp sees the same ones ----
      if (flags & BM25_F_SLAB_END) {
#ifdef LRAG_BM25_TIMING
        int path = 0;
        const long long te0 = clock64();
        thr_slab = bm25_slab_end(p, sh, slab0, thr_slab, noted, seq, &path); noted = false;
        const long long te1 = clock64();
        tacc[3] += te1 - te0; tacc[8 + path] += te1 - te0; tacc[12 + path] += 1;
#else
        thr_slab = bm25_slab_end(p, sh, slab0, thr_slab, noted, seq); noted = false;
#endif
#ifdef LRAG_BM25_TIMING
        if (blockIdx.x == 0 && seq < 1024 && lane == 0) {
          long long* tr = g_bm25_trace + (size_t(seq) * 16 + warp) * 4;
          tr[0] = tc0; tr[1] = tc1; tr[2] = (nchunks > warp) ? (nchunks - warp + BM25_WARPS - 1) / BM25_WARPS : 0; tr[3] = clock64();
        }
#endif
        ++seq;
      }
      if (flags & BM25_F_ITEM_END) BM25_T(4, bm25_item_end(p, sh, h1.x, h1.y, flags & BM25_F_FINAL));
      if (flags & BM25_F_END) break;
      thr_i = thr_slab;
    }
#ifdef LRAG_BM25_TIMING
    if (lane == 0) for (int i = 0; i < 16; ++i) atomicAdd(&g_bm25_t[i], (unsigned long long)tacc[i]);
#endif
  }
}

// Merge the per-split key lists of one query; when impacts are non-negative, documents that matched
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

struct Bm25Plan {
  int S, cap, P, TS, slab, item_slabs, steps, nc;
  int64_t dps;
  unsigned long long total_items;
  size_t smem, ws;
  size_t off[13];
};

// Tuning knob (process-wide, like the launch profiler: set it before serving, not concurrently with launches):
// slabs per work item; 0 = derive from BM25_DEFAULT_ITEM_DOCS.
static int g_item_slabs = -1;    // -1 = environment not read yet
static int bm25_item_slabs(int slab) {
  if (g_item_slabs < 0) {
    const char* e = getenv("LRAG_BM25_ITEM_SLABS");
    const int v = e ? atoi(e) : 0;
    g_item_slabs = (v < 0 || v > 4096) ? 0 : v;
  }
  if (g_item_slabs > 0) return g_item_slabs;
  const int d = BM25_DEFAULT_ITEM_DOCS / slab;
  return d < 1 ? 1 : d;
}

// Shared memory one of two resident CTAs of an SM can have: 228 KB per SM, 1 KB reserved per CTA.
constexpr size_t BM25_SMEM_PER_CTA = (228 * 1024 - 2 * 1024) / 2;

static Bm25Plan bm25_plan(int64_t N, int nq, int k, int64_t max_query_terms, int sms) {
  Bm25Plan pl;
  pl.P = next_pow2(k);
  pl.cap = 2 * pl.P < 512 ? 512 : 2 * pl.P;         // <= 2048 = 4 * BM25_CONSUMERS
  pl.TS = int(max_query_terms < 1 ? 1 : (max_query_terms > BM25_MAXT ? BM25_MAXT : max_query_terms));
  // the slab: as many docs as leave room for a second CTA on the SM, and no more than the shard has
  const size_t fixed = size_t(pl.cap) * 8 + BM25_SH_BYTES;
  int64_t slab = int64_t((BM25_SMEM_PER_CTA - fixed) / 4 / BM25_SLAB_STEP) * BM25_SLAB_STEP;
  slab = std::min<int64_t>(slab, BM25_SLAB_MAX);
  slab = std::min<int64_t>(slab, (std::max<int64_t>(N, 1) + BM25_SLAB_STEP - 1) / BM25_SLAB_STEP * BM25_SLAB_STEP);
  slab = std::max<int64_t>(slab, BM25_SLAB_STEP);
  pl.slab = int(slab);
  pl.item_slabs = bm25_item_slabs(pl.slab);
  // few queries: split every query's doc range into independent chains so that the machine is full, with items
  // small enough that there is about one chain per resident CTA (equal doc ranges of one query are equal work)
  const int64_t resident = 2 * int64_t(sms);
  if (nq < 2 * resident) {
    const int64_t nslab = (N + slab - 1) / slab;
    int64_t want = (nslab * nq + resident - 1) / resident;
    if (want < 1) want = 1;
    if (want < pl.item_slabs) pl.item_slabs = int(want);
  }
  const int64_t item_docs = int64_t(pl.item_slabs) * slab;
  int64_t n_items = (N + item_docs - 1) / item_docs;
  if (n_items < 1) n_items = 1;
  int64_t S = 1;
  if (nq < 2 * resident) { S = (2 * resident + nq - 1) / nq; if (S > n_items) S = n_items; }
  const int64_t items_per_split = (n_items + S - 1) / S;
  pl.dps = items_per_split * item_docs;
  S = (N + pl.dps - 1) / pl.dps;
  if (S < 1) S = 1;
  pl.S = int(S);
  pl.steps = int(items_per_split);
  pl.nc = int(int64_t(nq) * S);
  pl.total_items = (unsigned long long)pl.steps * (unsigned long long)pl.nc;
  pl.smem = size_t(pl.slab) * 4 + fixed;
  size_t o = 0;
  auto take = [&](int i, size_t bytes) { pl.off[i] = o; o += align_up(bytes, 256); };
  take(0, 8);                                         // counter
  take(1, size_t(pl.nc) * 4);                         // cur_flag
  take(2, size_t(pl.nc) * 4);                         // cand_flag
  take(3, size_t(nq) * 4);                            // q_nt
  take(4, size_t(nq) * pl.TS * 8);                    // tq_start
  take(5, size_t(nq) * pl.TS * 4);                    // tq_len
  take(6, size_t(nq) * pl.TS * 4);                    // tq_mult
  take(7, size_t(pl.nc) * pl.TS * 4);                 // ch_cur
  take(8, size_t(pl.nc) * 8);                         // ch_thr
  take(9, size_t(pl.nc) * 4);                         // ch_cnt
  take(10, pl.steps > 1 ? size_t(pl.nc) * pl.cap * 8 : 8);   // ch_cand
  take(11, size_t(pl.nc) * k * 8);                    // out_keys
  take(12, size_t(nq) * 4);                           // q_inv_scale
  pl.ws = o;
  return pl;
}

}  // namespace lrag

using namespace lrag;

#ifdef LRAG_BM25_TIMING
extern "C" int lrag_bm25_debug_timing(unsigned long long* out, int reset) {
  unsigned long long z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (reset) return cudaMemcpyToSymbol(g_bm25_t, z, sizeof(z)) == cudaSuccess ? 0 : -1;
  return cudaMemcpyFromSymbol(out, g_bm25_t, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif

#ifdef LRAG_BM25_TIMING
extern "C" int lrag_bm25_debug_trace(long long* out) {
  return cudaMemcpyFromSymbol(out, g_bm25_trace, sizeof(long long) * 1024 * 16 * 4) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int lrag_bm25_set_item_slabs(int slabs) {
  LRAG_REQUIRE(slabs >= 0 && slabs <= 4096, "bm25_set_item_slabs: %d out of range (0 = default, max 4096)", slabs);
  g_item_slabs = slabs;
  return LRAG_OK;
}

extern "C" size_t lrag_bm25_topk_workspace_bytes(int64_t N, int nq, int k, int64_t max_query_terms) {
  if (N < 0 || nq <= 0 || k <= 0) return 0;
  return bm25_plan(N, nq, k, max_query_terms, sm_count()).ws;
}

extern "C" int lrag_bm25_topk(const int64_t* indptr, const int32_t* doc_id, const float* impact, int64_t V, int64_t nnz,
                              const int64_t* q_indptr, const int32_t* q_term, int nq, int64_t max_query_terms,
                              int64_t N, int k, int64_t id_base, int nonneg, float impact_bound, float* out_score,
                              int64_t* out_id, void* ws, size_t ws_bytes, lrag_stream_t stream_) {
  return lrag_bm25_topk_dense(indptr, doc_id, impact, V, nnz, nullptr, nullptr, 0, 0, q_indptr, q_term, nq, max_query_terms, N, k,
                              id_base, nonneg, impact_bound, out_score, out_id, ws, ws_bytes, stream_);
}

extern "C" int lrag_bm25_topk_dense(const int64_t* indptr, const int32_t* doc_id, const float* impact, int64_t V, int64_t nnz,
                                    const int32_t* dense_term, const float* dense_rows, int n_dense, int64_t dense_stride,
                                    const int64_t* q_indptr, const int32_t* q_term, int nq, int64_t max_query_terms,
                                    int64_t N, int k, int64_t id_base, int nonneg, float impact_bound, float* out_score,
                                    int64_t* out_id, void* ws, size_t ws_bytes, lrag_stream_t stream_) {
  return lrag_bm25_topk_part(indptr, doc_id, impact, V, nnz, dense_term, dense_rows, n_dense, dense_stride, q_indptr, q_term, nq,
                             max_query_terms, N, k, id_base, nonneg, impact_bound, 0, nullptr, out_score, out_id, ws, ws_bytes, stream_);
}

// CTAs the scan launches when it may use at most `max_sms` SMs (0 = the whole machine): `occ` resident CTAs on each
static unsigned long long bm25_grid(const Bm25Plan& pl, int max_sms, int occ) {
  const int sms = max_sms > 0 && max_sms < sm_count() ? max_sms : sm_count();
  unsigned long long grid = (unsigned long long)occ * sms;
  if (grid > pl.total_items) grid = pl.total_items;
  return grid;
}

extern "C" int lrag_bm25_grid(int64_t N, int nq, int k, int64_t max_query_terms, int max_sms) {
  if (N < 0 || nq <= 0 || k <= 0 || k > LRAG_MAX_K || max_sms < 0 || !initialised()) return 0;
  const Bm25Plan pl = bm25_plan(N, nq, k, max_query_terms, sm_count());
  int occ = 0;
  cudaFuncSetAttribute(bm25_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl.smem));
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bm25_scan_kernel, BM25_THREADS, pl.smem) != cudaSuccess || occ < 1) return 0;
  return int(bm25_grid(pl, max_sms, occ));
}

extern "C" int lrag_bm25_topk_part(const int64_t* indptr, const int32_t* doc_id, const float* impact, int64_t V, int64_t nnz,
                                   const int32_t* dense_term, const float* dense_rows, int n_dense, int64_t dense_stride,
                                   const int64_t* q_indptr, const int32_t* q_term, int nq, int64_t max_query_terms,
                                   int64_t N, int k, int64_t id_base, int nonneg, float impact_bound, int max_sms,
                                   unsigned long long* start_counter, float* out_score,
                                   int64_t* out_id, void* ws, size_t ws_bytes, lrag_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LRAG_REQUIRE(n_dense >= 0 && n_dense <= 32 && (n_dense == 0 || (dense_term && dense_rows && dense_stride >= N && dense_stride % 32 == 0 &&
                                                (reinterpret_cast<uintptr_t>(dense_rows) & 127) == 0)),
               "bm25_topk: dense rows (at most 32) need sorted term ids, 128-byte aligned rows and a row stride >= N that is a multiple of 32 "
               "(n_dense=%d stride=%lld)", n_dense, (long long)dense_stride);
  LRAG_REQUIRE(initialised(), "lrag_init has not been called");
  LRAG_REQUIRE(nq > 0 && k > 0 && k <= LRAG_MAX_K, "bm25_topk: need nq > 0 and 1 <= k <= %d (nq=%d k=%d)", LRAG_MAX_K, nq, k);
  LRAG_REQUIRE(N >= 0 && N < (int64_t(1) << 31) - BM25_SLAB_MAX, "bm25_topk: N=%lld out of range for one shard", (long long)N);
  LRAG_REQUIRE(V >= 0 && V < (int64_t(1) << 31) && nnz >= 0, "bm25_topk: V=%lld nnz=%lld out of range", (long long)V, (long long)nnz);
  LRAG_REQUIRE(max_query_terms >= 0 && max_query_terms <= LRAG_BM25_MAX_QUERY_TERMS,
               "bm25_topk: a query has %lld terms; at most %d are supported", (long long)max_query_terms,
               LRAG_BM25_MAX_QUERY_TERMS);
  LRAG_REQUIRE(indptr && q_indptr && out_score && out_id, "bm25_topk: null pointer");
  LRAG_REQUIRE(impact_bound >= 0.f && impact_bound < 3.0e38f, "bm25_topk: impact_bound must be a finite value >= max |impact|");
  LRAG_REQUIRE((reinterpret_cast<uintptr_t>(doc_id) & 3) == 0 && (reinterpret_cast<uintptr_t>(impact) & 3) == 0,
               "bm25_topk: doc_id and impact must be 4-byte aligned");
  const int sms = sm_count();
  const Bm25Plan pl = bm25_plan(N, nq, k, max_query_terms, sms);
  if (ws_bytes < pl.ws || !ws) { set_error("bm25_topk: workspace %zu < required %zu", ws_bytes, pl.ws); return LRAG_ENOSPC; }
  LRAG_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "bm25_topk: workspace must be 16-byte aligned");
  uint8_t* w = static_cast<uint8_t*>(ws);
  Bm25Params p;
  p.indptr = indptr; p.doc_id = doc_id; p.impact = impact; p.V = V; p.nnz = nnz; p.q_indptr = q_indptr; p.q_term = q_term;
  p.dense_term = dense_term; p.dense_rows = dense_rows; p.n_dense = n_dense; p.dense_stride = dense_stride;
  p.N = N; p.dps = pl.dps; p.total_items = pl.total_items; p.nq = nq; p.k = k; p.nonneg = nonneg ? 1 : 0;
  p.S = pl.S; p.cap = pl.cap; p.P = pl.P; p.TS = pl.TS; p.item_slabs = pl.item_slabs; p.steps = pl.steps; p.nc = pl.nc; p.slab = pl.slab;
  p.ws.counter = reinterpret_cast<unsigned long long*>(w + pl.off[0]);
  p.ws.cur_flag = reinterpret_cast<int*>(w + pl.off[1]);
  p.ws.cand_flag = reinterpret_cast<int*>(w + pl.off[2]);
  p.ws.q_nt = reinterpret_cast<int*>(w + pl.off[3]);
  p.ws.tq_start = reinterpret_cast<int64_t*>(w + pl.off[4]);
  p.ws.tq_len = reinterpret_cast<int32_t*>(w + pl.off[5]);
  p.ws.tq_mult = reinterpret_cast<float*>(w + pl.off[6]);
  p.ws.ch_cur = reinterpret_cast<int32_t*>(w + pl.off[7]);
  p.ws.ch_thr = reinterpret_cast<unsigned long long*>(w + pl.off[8]);
  p.ws.ch_cnt = reinterpret_cast<int*>(w + pl.off[9]);
  p.ws.ch_cand = reinterpret_cast<uint64_t*>(w + pl.off[10]);
  p.ws.out_keys = reinterpret_cast<uint64_t*>(w + pl.off[11]);
  p.ws.q_inv_scale = reinterpret_cast<float*>(w + pl.off[12]);
  p.impact_bound = impact_bound;
  p.start_counter = start_counter;
  LRAG_REQUIRE(max_sms >= 0, "bm25_topk: max_sms=%d must be >= 0 (0 = the whole machine)", max_sms);

  static size_t smem_set[LRAG_MAX_DEVICES] = {};      // function attributes are per device
  const int dev = device_slot();
  if (pl.smem > smem_set[dev]) {
    LRAG_CHECK_CUDA(cudaFuncSetAttribute(bm25_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl.smem)));
    smem_set[dev] = pl.smem;
  }
  int occ = 0;
  LRAG_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bm25_scan_kernel, BM25_THREADS, pl.smem));
  LRAG_REQUIRE(occ >= 1, "bm25_topk: the scan kernel does not fit on an SM (smem %zu)", pl.smem);
  const unsigned long long grid = bm25_grid(pl, max_sms, occ);
  const int prep_blocks = int(std::min<int64_t>((int64_t(nq) + 7) / 8, 4 * int64_t(sms)));
  bm25_prepare_kernel<<<prep_blocks, 256, 0, stream>>>(p);
  LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
  prof_begin(stream, PROF_BM25_SCAN);
  bm25_scan_kernel<<<unsigned(grid), BM25_THREADS, pl.smem, stream>>>(p);
  prof_end(stream);
  LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
  bm25_merge_kernel<<<nq, SELECT_THREADS, select_smem_bytes(k), stream>>>(p.ws.out_keys, pl.S, k, pl.P, N, id_base,
                                                                          p.nonneg, out_score, out_id);
  LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
  return LRAG_OK;
}
