// BM25 channel: Okapi scoring over term-major CSR postings + exact top-k (replaces
// `bm25.get_scores(tokens)` and the full Python sort at legalrag/retrieval/bm25_retriever.py:74-75
// of the reference; rank_bm25.BM25Okapi semantics, see oracle/bm25.py).
//
// Integer/float streaming work over posting lists, no tensor cores; HBM sees every posting range once,
// the rest is served from L2, so the binding resource is the SM side (issue slots, shared memory).
//
// Work decomposition.  A *chain* is one (query, doc split) pair; it walks its split in *items* of
// `item_slabs` slabs (BM25_SLAB docs = 48 KB of int32 fixed-point accumulators in shared memory).  Items are
// numbered step-major -- item = step * n_chains + chain -- and handed out by an atomic counter to
// persistent CTAs, so at any moment the whole machine works on the same narrow doc range of every
// query: a posting list is fetched from HBM once per range and then served from L2 to all the other
// queries that share the term.  A chain's state between two of its items (term cursors, running
// threshold, surviving candidates) lives in the workspace; item (step, chain) waits on a flag that
// item (step - 1, chain) -- always a lower item number, so already claimed by a running CTA -- sets.
//
// The CTA is warp-specialised so that posting traffic never waits for the arithmetic and the
// arithmetic never does bookkeeping:
//   * bounds warp   -- claims the next item, loads the chain's term cursors and finds the posting
//                      boundaries of every query term at every slab edge with one parallel round of
//                      windowed binary searches (postings are doc-id sorted); double-buffered, one
//                      group of slabs ahead of the copy warps;
//   * 3 copy warps  -- walk the (slab, term) runs in order (the lanes look up one term each) and stream
//                      them, <= BM25_CHUNK postings at a time, into a shared-memory ring with bulk async
//                      copies (cp.async.bulk, 16-byte aligned source windows) that complete on mbarriers;
//                      chunk c is issued by warp c mod 3.  The short runs of a slab (rare terms) are
//                      packed into ONE stage with 4-byte cp.async copies.  Each ring stage carries a
//                      descriptor {count, skip, term multiplicity x scale, slab base, flags}: the
//                      consumers are a plain interpreter of that stream;
//   * 16 consumer warps -- per stage add `mult * impact` into acc[doc - slab0] with one shared-memory
//                      integer atomic per posting: scores are kept in fixed point (int32, a per-query
//                      power-of-two scale sized from `impact_bound`), so adds commute exactly, the
//                      terms of a slab need no barrier between them and the result is independent
//                      of scheduling.  Each thread keeps a running max of what it wrote.  At
//                      SLAB_END one `bar.red.or` tells whether any thread wrote a score that reaches
//                      the query's running k-th best: only then is the slab scanned for candidates
//                      (appended to a shared buffer; an overflow triggers an exact radix select that
//                      raises the threshold).  The slab is re-zeroed.
//                      ITEM_BEGIN / ITEM_END stages load / store the chain's candidate state; the
//                      chain's last item sorts its top-k into the per-chain key list.
// Slabs in which no query term has a posting are skipped when impacts are known non-negative;
// documents that match nothing (score 0) are then added by the merge step, lowest id first, exactly
// as the reference's stable sort does.  bm25_merge_kernel merges the per-split lists of a query.
#include <algorithm>
#include <climits>
#include <cstdlib>

#include "common.cuh"
#include "select.cuh"

namespace lrag {

constexpr int BM25_CONSUMERS = 512;                  // 16 warps
constexpr int BM25_COPY_WARPS = 3;                   // chunk c is issued by copy warp c % BM25_COPY_WARPS; must not exceed BM25_STAGES (phase parity)
constexpr int BM25_THREADS = BM25_CONSUMERS + 32 * BM25_COPY_WARPS + 32;   // consumers, copy warps, bounds warp
constexpr int BM25_SLAB = 12288;
constexpr int BM25_MAX_GROUP = 16;                   // slabs per bounds group (fewer when a query has many terms)
constexpr int BM25_BOUND_CAP = 17 * 32;              // ints per bounds buffer: (group + 1) * nt must fit
constexpr int BM25_MAXT = LRAG_BM25_MAX_QUERY_TERMS;
constexpr int BM25_CHUNK = 2048;                     // postings per ring stage
constexpr int BM25_STAGES = 3;                       // 48 KB slab + 48 KB ring + 4 KB candidates + 11 KB state: two CTAs per SM
constexpr int BM25_RING_BYTES = BM25_STAGES * BM25_CHUNK * 8;
constexpr int BM25_BAR_CONSUMERS = 1;                // named barrier id of the consumer warps
constexpr int BM25_PER_THREAD = BM25_CHUNK / BM25_CONSUMERS;
constexpr int BM25_KEEP = 2 * LRAG_MAX_K / BM25_CONSUMERS;   // candidate keys a thread may hold across a compaction
static_assert(BM25_COPY_WARPS <= BM25_STAGES, "a copy warp may run at most one ring phase ahead of the consumers");
constexpr int BM25_DEFAULT_ITEM_SLABS = 32;
constexpr int BM25_GATHER_MAX = 64;                  // runs this short share one ring stage (32 runs x 64 postings fill it at most)
static_assert(32 * BM25_GATHER_MAX <= BM25_CHUNK, "a gather stage must hold one lane batch of short runs");
enum : int { BM25_F_SLAB_END = 1, BM25_F_ITEM_BEGIN = 2, BM25_F_ITEM_END = 4, BM25_F_FINAL = 8, BM25_F_END = 16 };

struct Bm25Ws {
  unsigned long long* counter;     // next item
  int* cur_flag;                   // [nc] steps whose final term cursors are published
  int* cand_flag;                  // [nc] steps whose candidate state is published
  int* q_nt;                       // [nq] distinct in-vocabulary terms with postings
  int64_t* tq_start;               // [nq, TS] first posting
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
  int64_t N; int64_t dps;          // docs per split (multiple of the item size)
  unsigned long long total_items;
  int nq, k, nonneg, S, cap, P, TS, item_slabs, steps, nc;
  float impact_bound;              // >= max |impact| over the index
  Bm25Ws ws;
};

struct Bm25Group {                 // bounds warp -> copy warp
  int kind;                        // 0 = slabs of an item, 1 = no more work
  int chain, step, first, last, final_step, nt, ns;
  int b0, range_end;
  float inv_scale;
  int64_t t_start[BM25_MAXT];
  float t_mult[BM25_MAXT];
  int32_t bound[BM25_BOUND_CAP];   // [t * (ns + 1) + j]
  int32_t last_t[BM25_MAX_GROUP];  // last term with postings in the slab, -1 = none
};

struct Bm25Shared {
  SelectShared sel;
  int4 sdesc[BM25_STAGES];         // {n | skip << 12 | flags << 16, mult x scale (float bits) or step, slab0 or chain, 1/scale bits}
  uint64_t full_bar[BM25_STAGES], empty_bar[BM25_STAGES];   // posting ring
  uint64_t bfull_bar[2], bempty_bar[2];                      // group buffers
  Bm25Group grp[2];
  int64_t t_start[BM25_MAXT];      // bounds warp's view of the current item's query
  int32_t t_len[BM25_MAXT];
  int32_t t_cur[BM25_MAXT];
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

// 4-byte asynchronous global -> shared copy of the executing thread, and the arrive that fires on an
// mbarrier once all of this thread's earlier copies have landed (pending count +1 now, -1 then)
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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
  const int* acc; float inv_scale; int64_t slab0; int64_t range_end; unsigned long long thr_key;
  template <class F> __device__ void operator()(F&& f) const {
    for (int i = threadIdx.x; i < ncand; i += BM25_CONSUMERS) f(cand[i]);
    for (int i = threadIdx.x; i < BM25_SLAB; i += BM25_CONSUMERS) {
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
        p.ws.tq_start[slot] = s;
        p.ws.tq_len[slot] = int32_t(e - s);
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
    for (int t = lane; t < base; t += 32) p.ws.tq_mult[size_t(q) * p.TS + t] *= scale;
    if (lane == 0) { p.ws.q_nt[q] = base; p.ws.q_inv_scale[q] = ldexpf(1.0f, -ex); }
  }
}

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
  const int cap = p.cap;
  const unsigned long long thr_init = p.nonneg ? ((uint64_t(ord32(0.0f)) << 32) | 0xffffffffull) : 0ull;

  if (tid == 0) {
    sh.cand_cnt = 0;
    sh.thr_key = thr_init;
    for (int s = 0; s < BM25_STAGES; ++s) { mbar_init(&sh.full_bar[s], 1); mbar_init(&sh.empty_bar[s], BM25_CONSUMERS / 32); }
    for (int s = 0; s < 2; ++s) { mbar_init(&sh.bfull_bar[s], 1); mbar_init(&sh.bempty_bar[s], BM25_COPY_WARPS); }
    fence_barrier_init();
  }
  // zero the slab once; every slab end leaves it zeroed again
  for (int i = tid; i < BM25_SLAB / 4; i += BM25_THREADS) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const NamedBarrier cbar{BM25_BAR_CONSUMERS, BM25_CONSUMERS};

  if (warp == BM25_CONSUMERS / 32 + BM25_COPY_WARPS) {
    // ===================== bounds warp =====================
    const int64_t item_docs = int64_t(p.item_slabs) * BM25_SLAB;
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
          sh.t_cur[t] = split_begin == 0 ? 0 : lower_bound_doc(p.doc_id + sh.t_start[t], 0, sh.t_len[t], split_begin);
      }
      __syncwarp();
      int gs = nt > 0 ? BM25_BOUND_CAP / nt - 1 : BM25_MAX_GROUP;
      gs = gs < 1 ? 1 : (gs > BM25_MAX_GROUP ? BM25_MAX_GROUP : gs);
      const int nslab = rb < re ? int((re - rb + BM25_SLAB - 1) / BM25_SLAB) : 0;
      const int ngroups = nslab > 0 ? (nslab + gs - 1) / gs : 1;
      for (int gi = 0; gi < ngroups; ++gi, ++gcount) {
        const uint32_t bb = gcount & 1;
        mbar_wait(&sh.bempty_bar[bb], ((gcount >> 1) & 1) ^ 1);
        Bm25Group& G = sh.grp[bb];
        const int ns = max(0, min(gs, nslab - gi * gs));
        const int64_t b0 = rb + int64_t(gi) * gs * BM25_SLAB;
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
            int64_t target = b0 + int64_t(j) * BM25_SLAB;
            if (target > re) target = re;
            // ids are strictly increasing inside a term: the answer is at most (target - b0) past cur
            const int64_t reach = int64_t(cur) + (target - b0);
            const int hi = int(reach < len ? reach : int64_t(len));
            G.bound[w] = (j == 0) ? cur : lower_bound_doc(p.doc_id + sh.t_start[t], cur, hi, target);
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
    mbar_wait(&sh.bempty_bar[bb], ((gcount >> 1) & 1) ^ 1);
    if (lane == 0) { sh.grp[bb].kind = 1; mbar_arrive(&sh.bfull_bar[bb]); }
  } else if (warp >= BM25_CONSUMERS / 32) {
    // ===================== copy warps: postings -> shared-memory ring =====================
    // Every copy warp walks the same chunk sequence and issues the chunks of its own residue class.
    // Per slab the lanes look up one term's run each; the runs are then issued in term order, one
    // short serial sequence per chunk (wait for a free stage, descriptor, expect_tx, two bulk copies).
    uint32_t ps = 0, pph = 1;
    int turn = warp - BM25_CONSUMERS / 32;      // chunks until this warp's next one
    // returns the stage of the next chunk, or -1 when the chunk belongs to another copy warp
    auto stage_acquire = [&]() {
      const uint32_t s = ps;
      const uint32_t ph = pph;
      if (++ps == BM25_STAGES) { ps = 0; pph ^= 1; }
      if (turn != 0) { --turn; return -1; }
      turn = BM25_COPY_WARPS - 1;
      mbar_wait(&sh.empty_bar[s], ph);
      return int(s);
    };
    // control stage: no postings, just flags and two words for the consumers
    auto emit_ctrl = [&](int flags, int y, int z, float w) {
      const int s = stage_acquire();
      if (s < 0) return;
      if (lane == 0) {
        sh.sdesc[s] = make_int4(flags << 16, y, z, __float_as_int(w));
        mbar_arrive(&sh.full_bar[s]);
      }
      __syncwarp();
    };
    // postings [first, first + n) -> one stage.  The copy window is widened to 16-byte units; the
    // consumers ignore what lies outside [skip, skip + n).
    auto emit = [&](int64_t first, int n, float mult, int sl0, int flags) {
      const int s = stage_acquire();
      if (s < 0) return;
      const int skip = int(first & 3);
      const int64_t a0 = first - skip;
      int cnt4 = (n + skip + 3) & ~3;
      int32_t* dst_id = ring_id + s * BM25_CHUNK;
      float* dst_imp = ring_imp + s * BM25_CHUNK;
      if (a0 + cnt4 > p.nnz) {
        // the arrays end inside the last unit: those (at most 3) postings are moved by hand
        cnt4 -= 4;
        const int tail = int(p.nnz - (a0 + cnt4));
        if (lane < tail) {
          dst_id[cnt4 + lane] = __ldg(p.doc_id + a0 + cnt4 + lane);
          dst_imp[cnt4 + lane] = __ldg(p.impact + a0 + cnt4 + lane);
        }
        __syncwarp();
      }
      if (lane == 0) {
        sh.sdesc[s] = make_int4(n | (skip << 12) | (flags << 16), __float_as_int(mult), sl0, 0);
        mbar_arrive_expect_tx(&sh.full_bar[s], uint32_t(cnt4) * 8u);
        if (cnt4 > 0) {
          bulk_copy_g2s(dst_id, p.doc_id + a0, uint32_t(cnt4) * 4u, &sh.full_bar[s]);
          bulk_copy_g2s(dst_imp, p.impact + a0, uint32_t(cnt4) * 4u, &sh.full_bar[s]);
        }
      }
      __syncwarp();
    };
    // Short runs (rare terms: a handful of postings per slab) with the same multiplier share ONE
    // stage: the lanes copy each run with 4-byte asynchronous copies packed back to back,
    // and the stage's barrier completes when every lane's copies have landed.
    auto emit_gather = [&](uint32_t runs, int64_t first, int cnt, float mult, int sl0, int flags) {
      const int s = stage_acquire();
      if (s < 0) return;
      int32_t* dst_id = ring_id + s * BM25_CHUNK;
      float* dst_imp = ring_imp + s * BM25_CHUNK;
      int total = 0;
      while (runs) {
        const int src = __ffs(runs) - 1;
        runs &= runs - 1;
        const int64_t f = __shfl_sync(0xffffffffu, first, src);
        const int c = __shfl_sync(0xffffffffu, cnt, src);
        for (int l = lane; l < c; l += 32) {
          cp_async_4(dst_id + total + l, p.doc_id + f + l);
          cp_async_4(dst_imp + total + l, p.impact + f + l);
        }
        total += c;
      }
      cp_async_mbar_arrive(&sh.full_bar[s]);
      __syncwarp();
      if (lane == 0) {
        sh.sdesc[s] = make_int4(total | (flags << 16), __float_as_int(mult), sl0, 0);
        mbar_arrive(&sh.full_bar[s]);
      }
      __syncwarp();
    };
    for (uint32_t gc = 0;; ++gc) {
      const uint32_t bb = gc & 1;
      mbar_wait(&sh.bfull_bar[bb], (gc >> 1) & 1);
      const Bm25Group& G = sh.grp[bb];
      if (G.kind != 0) { emit_ctrl(BM25_F_END, 0, 0, 0.f); break; }
      const int nt = G.nt, ns = G.ns, chain = G.chain, step = G.step;
      if (G.first) emit_ctrl(BM25_F_ITEM_BEGIN, step, chain, G.inv_scale);
      for (int j = 0; j < ns; ++j) {
        const int sl0 = G.b0 + j * BM25_SLAB;
        const int last_t = G.last_t[j];
        if (last_t < 0) {
          // no postings here; with negative impacts a slab of zero scores still has to be ranked
          if (!p.nonneg) emit_ctrl(BM25_F_SLAB_END, 0, sl0, 0.f);
          continue;
        }
        for (int t0 = 0; t0 <= last_t; t0 += 32) {
          const int t = t0 + lane;
          int64_t first = 0; int cnt = 0; float mult = 0.f;
          if (t <= last_t) {
            const int lo = G.bound[t * (ns + 1) + j];
            cnt = G.bound[t * (ns + 1) + j + 1] - lo;
            first = G.t_start[t] + lo;
            mult = G.t_mult[t];
          }
          uint32_t live = __ballot_sync(0xffffffffu, cnt > 0);
          // short runs of a (single-batch) slab that share a multiplier travel together
          uint32_t shorts = 0;
          float short_mult = 0.f;
          if (last_t < 32) {
            const uint32_t cand_short = __ballot_sync(0xffffffffu, cnt > 0 && cnt <= BM25_GATHER_MAX);
            if (__popc(cand_short) >= 2) {
              short_mult = __shfl_sync(0xffffffffu, mult, __ffs(cand_short) - 1);
              shorts = __ballot_sync(0xffffffffu, cnt > 0 && cnt <= BM25_GATHER_MAX && mult == short_mult);
              if (__popc(shorts) < 2) shorts = 0;
            }
          }
          live &= ~shorts;
          const int last_run = shorts ? -1 : (last_t < 32 ? 31 - __clz(int(live)) : last_t);   // run that ends the slab
          while (live) {
            const int src = __ffs(live) - 1;
            live &= live - 1;
            int64_t pos = __shfl_sync(0xffffffffu, first, src);
            int rem = __shfl_sync(0xffffffffu, cnt, src);
            const float m = __shfl_sync(0xffffffffu, mult, src);
            const int endflags = (t0 + src == last_run) ? BM25_F_SLAB_END : 0;
            while (rem > 0) {
              const int n = min(rem, BM25_CHUNK - int(pos & 3));
              emit(pos, n, m, sl0, n == rem ? endflags : 0);
              pos += n;
              rem -= n;
            }
          }
          if (shorts) emit_gather(shorts, first, cnt, short_mult, sl0, BM25_F_SLAB_END);
        }
      }
      if (G.last) emit_ctrl(BM25_F_ITEM_END | (G.final_step ? BM25_F_FINAL : 0), step, chain, 0.f);
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.bempty_bar[bb]);
    }
  } else {
    // ===================== consumers =====================
    // Scores are accumulated in fixed point (int32, per-query power-of-two scale) with shared-memory
    // integer atomics: adds commute exactly, so the terms of a slab need no ordering between them,
    // results do not depend on which warp ran first, and one ATOMS replaces a load / add / store.
    int mx = INT_MIN;         // largest score this thread wrote into the current slab
    float inv_scale = 1.f;    // fixed point -> fp32 score of the current item's query
    int64_t range_end = 0;    // docs at or past the end of the chain's split are not ranked
    int* acci = reinterpret_cast<int*>(acc);
    const uint32_t acc_u32 = smem_u32(acc);
    const uint32_t rid0 = smem_u32(ring_id) + tid * 4, rim0 = smem_u32(ring_imp) + tid * 4;
    uint32_t s = 0, ph = 0;
    for (;;) {
      mbar_wait(&sh.full_bar[s], ph);
      const int4 de = sh.sdesc[s];
      const int n = de.x & 0xfff, flags = de.x >> 16, sl0 = de.z;
      const float ms = __int_as_float(de.y);                      // term multiplicity x scale
      const uint32_t accb = acc_u32 - 4u * uint32_t(sl0);        // accb + 4 * doc == &acc[doc - slab0]
      const uint32_t rid = rid0 + s * (BM25_CHUNK * 4), rim = rim0 + s * (BM25_CHUNK * 4);
      int d[BM25_PER_THREAD]; float v[BM25_PER_THREAD];
      if (n == BM25_CHUNK) {
        // whole stage: no predicates.  The stage is handed back as soon as it is in registers.
#pragma unroll
        for (int u = 0; u < BM25_PER_THREAD; ++u) { d[u] = lds_s32(rid + u * (BM25_CONSUMERS * 4)); v[u] = lds_f32(rim + u * (BM25_CONSUMERS * 4)); }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.empty_bar[s]);
#pragma unroll
        for (int u = 0; u < BM25_PER_THREAD; ++u) {
          const int vi = __float2int_rn(v[u] * ms);
          mx = max(mx, atoms_add_s32(accb + 4u * uint32_t(d[u]), vi) + vi);
        }
      } else if (n <= BM25_CONSUMERS) {
        // short run (the usual partial stage): at most one posting per thread, and warps past the
        // end of the run only hand the stage back
        const int skip = (de.x >> 12) & 3;
        const bool ok = tid < n;
        if (ok) { d[0] = lds_s32(rid + skip * 4); v[0] = lds_f32(rim + skip * 4); }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.empty_bar[s]);
        if (ok) {
          const int vi = __float2int_rn(v[0] * ms);
          mx = max(mx, atoms_add_s32(accb + 4u * uint32_t(d[0]), vi) + vi);
        }
      } else {
        const int skip = (de.x >> 12) & 3;
#pragma unroll
        for (int u = 0; u < BM25_PER_THREAD; ++u) {
          const bool ok = tid + u * BM25_CONSUMERS < n;
          d[u] = ok ? lds_s32(rid + (skip + u * BM25_CONSUMERS) * 4) : -1;
          v[u] = ok ? lds_f32(rim + (skip + u * BM25_CONSUMERS) * 4) : 0.f;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.empty_bar[s]);
#pragma unroll
        for (int u = 0; u < BM25_PER_THREAD; ++u)
          if (d[u] >= 0) {
            const int vi = __float2int_rn(v[u] * ms);
            mx = max(mx, atoms_add_s32(accb + 4u * uint32_t(d[u]), vi) + vi);
          }
      }
      if (++s == BM25_STAGES) { s = 0; ph ^= 1; }
      if (flags == 0) continue;

      if (flags & BM25_F_SLAB_END) {
        const int64_t slab0 = sl0;
        const int cnt_before = sh.cand_cnt;
        const unsigned long long thr_key = sh.thr_key;
        // initial thresholds: nothing yet (-inf), or "strictly positive" when zero scores are filled in later
        const float thr_s = !thr_key ? -INFINITY
                            : (uint32_t(thr_key) == 0xffffffffu && key_score(thr_key) == 0.f) ? 1.4e-45f : key_score(thr_key);
        // every add of the slab is done behind this barrier; with negative impacts an untouched doc
        // (score 0) can be a hit too
        const bool hit = consumers_bar_or(!p.nonneg || float(mx) * inv_scale >= thr_s);
        if (hit) {
          // ---- scan the slab for candidates ----
          // Coarse filter in the integer domain (no conversion for the ~all accumulators that cannot matter):
          // thr_i is a lower bound of every fixed-point value whose fp32 score reaches thr_s.
          int thr_i = INT_MIN;
          if (thr_s > -INFINITY) {
            const float t = thr_s / inv_scale;                     // inv_scale is a power of two: exact
            thr_i = t >= 2147483520.f ? INT_MAX : (t <= -2147483520.f ? INT_MIN : int(floorf(t - fabsf(t) * 2.4e-7f)) - 1);
            if (p.nonneg && thr_i < 1) thr_i = 1;                  // zero scores are never candidates then
          }
          const int4* a4 = reinterpret_cast<const int4*>(acc);
#pragma unroll 2
          for (int i = 0; i < BM25_SLAB / 4 / BM25_CONSUMERS; ++i) {
            const int idx = (tid + i * BM25_CONSUMERS) * 4;
            const int4 s4 = a4[tid + i * BM25_CONSUMERS];
            const bool any4 = max(max(s4.x, s4.y), max(s4.z, s4.w)) >= thr_i;
            if (!__any_sync(0xffffffffu, any4)) continue;
            const float sv[4] = {float(s4.x) * inv_scale, float(s4.y) * inv_scale, float(s4.z) * inv_scale, float(s4.w) * inv_scale};
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
            Bm25Union uni{cand, cnt_before, acci, inv_scale, slab0, range_end, thr_key};
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
            for (int i = tid; i < BM25_SLAB; i += BM25_CONSUMERS) {
              const int64_t doc = slab0 + i;
              uint64_t key = 0;
              bool want = doc < range_end;
              if (want) { key = make_key(float(acci[i]) * inv_scale, uint32_t(doc)); want = (key > thr_key) && (key >= pivot); }
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
        mx = INT_MIN;
        cbar();
      }

      if (flags & BM25_F_ITEM_BEGIN) {
        // the previous item ended behind a barrier: the candidate buffer is free
        const int chain = de.z, step = de.y;
        inv_scale = __int_as_float(de.w);
        const int split = chain % p.S;
        range_end = min(p.N, int64_t(split + 1) * p.dps);
        if (step == 0) {
          if (tid == 0) { sh.cand_cnt = 0; sh.thr_key = thr_init; }
        } else {
          spin_until_ge(p.ws.cand_flag + chain, step);
          const int cnt = __ldcg(p.ws.ch_cnt + chain);
          const uint64_t* src = p.ws.ch_cand + size_t(chain) * cap;
          for (int i = tid; i < cnt; i += BM25_CONSUMERS) cand[i] = __ldcg(src + i);
          if (tid == 0) { sh.cand_cnt = cnt; sh.thr_key = __ldcg(p.ws.ch_thr + chain); }
        }
        cbar();
      }
      if (flags & BM25_F_ITEM_END) {
        const int chain = de.z, step = de.y;
        if (flags & BM25_F_FINAL) {
          // ---- sorted top-k of the surviving candidates -> this chain's key list ----
          const int ncand = min(sh.cand_cnt, cap);
          Bm25Cands cands{cand, ncand};
          const unsigned long long pivot = block_select_pivot(cands, p.k, sh.sel, cbar);
          uint64_t* sortbuf = reinterpret_cast<uint64_t*>(acc);      // the (zeroed) slab doubles as the sort buffer
          const int P = p.P;
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
          for (int i = tid; i < P; i += BM25_CONSUMERS) sortbuf[i] = 0;   // leave the slab zeroed
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
      if (flags & BM25_F_END) break;
    }
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
  int S, cap, P, TS, item_slabs, steps, nc;
  int64_t dps;
  unsigned long long total_items;
  size_t smem, ws;
  size_t off[13];
};

static int g_item_slabs = 0;     // 0 = not yet initialised
static int bm25_item_slabs() {
  if (!g_item_slabs) {
    const char* e = getenv("LRAG_BM25_ITEM_SLABS");   // tuning knob: docs per work item = value * 16384
    const int v = e ? atoi(e) : BM25_DEFAULT_ITEM_SLABS;
    g_item_slabs = (v < 1 || v > 4096) ? BM25_DEFAULT_ITEM_SLABS : v;
  }
  return g_item_slabs;
}

static Bm25Plan bm25_plan(int64_t N, int nq, int k, int64_t max_query_terms, int sms) {
  Bm25Plan pl;
  pl.P = next_pow2(k);
  pl.cap = 2 * pl.P < 512 ? 512 : 2 * pl.P;         // <= 2048 = 4 * BM25_CONSUMERS
  pl.TS = int(max_query_terms < 1 ? 1 : (max_query_terms > BM25_MAXT ? BM25_MAXT : max_query_terms));
  pl.item_slabs = bm25_item_slabs();
  // few queries: split every query's doc range into independent chains so that the machine is full, with items
  // small enough that there is about one chain per resident CTA (equal doc ranges of one query are equal work)
  const int64_t resident = 2 * int64_t(sms);
  if (nq < 2 * resident) {
    const int64_t nslab = (N + BM25_SLAB - 1) / BM25_SLAB;
    int64_t want = (nslab * nq + resident - 1) / resident;
    if (want < 1) want = 1;
    if (want < pl.item_slabs) pl.item_slabs = int(want);
  }
  const int64_t item_docs = int64_t(pl.item_slabs) * BM25_SLAB;
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
  pl.smem = size_t(BM25_SLAB) * 4 + BM25_RING_BYTES + size_t(pl.cap) * 8 + sizeof(Bm25Shared);
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

extern "C" int lrag_bm25_set_item_slabs(int slabs) {
  LRAG_REQUIRE(slabs >= 0 && slabs <= 4096, "bm25_set_item_slabs: %d out of range (0 = default, max 4096)", slabs);
  g_item_slabs = slabs ? slabs : BM25_DEFAULT_ITEM_SLABS;
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
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LRAG_REQUIRE(initialised(), "lrag_init has not been called");
  LRAG_REQUIRE(nq > 0 && k > 0 && k <= LRAG_MAX_K, "bm25_topk: need nq > 0 and 1 <= k <= %d (nq=%d k=%d)", LRAG_MAX_K, nq, k);
  LRAG_REQUIRE(N >= 0 && N < (int64_t(1) << 31) - BM25_SLAB, "bm25_topk: N=%lld out of range for one shard", (long long)N);
  LRAG_REQUIRE(V >= 0 && V < (int64_t(1) << 31) && nnz >= 0, "bm25_topk: V=%lld nnz=%lld out of range", (long long)V, (long long)nnz);
  LRAG_REQUIRE(max_query_terms >= 0 && max_query_terms <= LRAG_BM25_MAX_QUERY_TERMS,
               "bm25_topk: a query has %lld terms; at most %d are supported", (long long)max_query_terms,
               LRAG_BM25_MAX_QUERY_TERMS);
  LRAG_REQUIRE(indptr && q_indptr && out_score && out_id, "bm25_topk: null pointer");
  LRAG_REQUIRE(impact_bound >= 0.f && impact_bound < 3.0e38f, "bm25_topk: impact_bound must be a finite value >= max |impact|");
  LRAG_REQUIRE((reinterpret_cast<uintptr_t>(doc_id) & 15) == 0 && (reinterpret_cast<uintptr_t>(impact) & 15) == 0,
               "bm25_topk: doc_id and impact must be 16-byte aligned (bulk async copies)");
  const int sms = sm_count();
  const Bm25Plan pl = bm25_plan(N, nq, k, max_query_terms, sms);
  if (ws_bytes < pl.ws || !ws) { set_error("bm25_topk: workspace %zu < required %zu", ws_bytes, pl.ws); return LRAG_ENOSPC; }
  LRAG_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "bm25_topk: workspace must be 16-byte aligned");
  uint8_t* w = static_cast<uint8_t*>(ws);
  Bm25Params p;
  p.indptr = indptr; p.doc_id = doc_id; p.impact = impact; p.V = V; p.nnz = nnz; p.q_indptr = q_indptr; p.q_term = q_term;
  p.N = N; p.dps = pl.dps; p.total_items = pl.total_items; p.nq = nq; p.k = k; p.nonneg = nonneg ? 1 : 0;
  p.S = pl.S; p.cap = pl.cap; p.P = pl.P; p.TS = pl.TS; p.item_slabs = pl.item_slabs; p.steps = pl.steps; p.nc = pl.nc;
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

  static size_t smem_set = 0;
  if (pl.smem > smem_set) {
    LRAG_CHECK_CUDA(cudaFuncSetAttribute(bm25_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl.smem)));
    smem_set = pl.smem;
  }
  int occ = 0;
  LRAG_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bm25_scan_kernel, BM25_THREADS, pl.smem));
  LRAG_REQUIRE(occ >= 1, "bm25_topk: the scan kernel does not fit on an SM (smem %zu)", pl.smem);
  unsigned long long grid = (unsigned long long)occ * sms;
  if (grid > pl.total_items) grid = pl.total_items;
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
