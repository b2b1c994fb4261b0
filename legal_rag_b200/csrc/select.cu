// Row-wise exact top-k over materialised scores, and the k-way merge that follows the NCCL
// all-gather of per-shard candidate lists.
//   lrag_topk_select_f32  <- `sorted(range(N), key=scores[i], reverse=True)[:k]`
//                            (legalrag/retrieval/bm25_retriever.py:75 of the reference) and the ranking
//                            of MaxSim candidate scores (colbert_retriever.py:152, Searcher.search)
//   lrag_topk_merge       <- no reference counterpart (the reference is single-process)
#include "select.cuh"
#include <algorithm>
#include <climits>

namespace lrag {

struct RowScores {
  const float* s; const int64_t* col_id; int64_t n;
  template <class F> __device__ void operator()(F&& f) const {
    // the tie field is the column: with column ids the id (any width) is looked up when the hit is written
    for (int64_t c = threadIdx.x; c < n; c += SELECT_THREADS) {
      if (col_id && col_id[c] < 0) continue;
      f(make_key(s[c], uint32_t(c)));
    }
  }
};

__global__ void __launch_bounds__(SELECT_THREADS)
topk_select_kernel(const float* S, int64_t ld, int64_t N, int k, int64_t id_base, const int64_t* col_id, int P,
                   float* out_score, int64_t* out_id) {
  extern __shared__ uint8_t sm_raw[];
  __shared__ SelectShared ss;
  const int q = blockIdx.x;
  RowScores rows{S + size_t(q) * ld, col_id ? col_id + size_t(q) * N : nullptr, N};
  block_topk_sorted(rows, k, P, ss, reinterpret_cast<uint64_t*>(sm_raw), id_base,
                    out_score + size_t(q) * k, out_id + size_t(q) * k, static_cast<uint64_t*>(nullptr), rows.col_id);
}

// (score, id) pairs of one row, spread over `shards` lists of `kin` entries: entry c = (shard c / kin, rank c % kin) lives at
// score[shard * shard_stride + c % kin] (one list: shards = 1).  Narrow rows (max id - min id < 2^32 - 1: every sharded corpus
// below 4 G docs) keep the exact (score desc, id asc) order through the key's tie field = id - min id; wider rows key on the
// entry index and look the 64-bit id up when the hit is written.
struct RowPairs {
  const float* s; const int64_t* id; int kin, n; int64_t s_stride, i_stride; int64_t base; bool wide;
  __device__ __forceinline__ float score_at(int c) const { return s[int64_t(c / kin) * s_stride + c % kin]; }
  __device__ __forceinline__ int64_t id_at(int c) const { return id[int64_t(c / kin) * i_stride + c % kin]; }
  template <class F> __device__ void operator()(F&& f) const {
    for (int c = threadIdx.x; c < n; c += SELECT_THREADS) {
      const int64_t i = id_at(c);
      if (i >= 0) f(make_key(score_at(c), wide ? uint32_t(c) : uint32_t(i - base)));
    }
  }
};

// score / id point at shard 0's list of row 0; a row's lists are `kin` apart inside a shard, shards are *_stride apart
__global__ void __launch_bounds__(SELECT_THREADS)
topk_merge_kernel(const float* score, const int64_t* id, int64_t s_stride, int64_t i_stride, int shards, int kin, int k, int P,
                  int ids_fit, float* out_score, int64_t* out_id) {
  extern __shared__ uint8_t sm_raw[];
  __shared__ SelectShared ss;
  __shared__ long long id_min, id_max;
  const int q = blockIdx.x;
  const int L = shards * kin;
  RowPairs rows{score + size_t(q) * kin, id + size_t(q) * kin, kin, L, s_stride, i_stride, 0, false};
  if (threadIdx.x == 0) { id_min = LLONG_MAX; id_max = -1; }
  __syncthreads();
  long long lo = LLONG_MAX, hi = -1;
  for (int c = threadIdx.x; c < L; c += SELECT_THREADS) {
    const long long i = rows.id_at(c);
    if (i >= 0) { lo = i < lo ? i : lo; hi = i > hi ? i : hi; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = l2 < lo ? l2 : lo; hi = h2 > hi ? h2 : hi;
  }
  if ((threadIdx.x & 31) == 0 && hi >= 0) { atomicMin(&id_min, lo); atomicMax(&id_max, hi); }
  __syncthreads();
  rows.wide = id_max >= 0 && (unsigned long long)(id_max - id_min) >= 0xffffffffull;
  rows.base = id_max >= 0 ? id_min : 0;
  // wide rows: the tie field is the entry index; the lookup table of the output stage is the id list itself, which is
  // contiguous only for a single list -- several lists go through a gather of the ids into shared memory first
  const int64_t* lookup = nullptr;
  int64_t* ids_sm = reinterpret_cast<int64_t*>(sm_raw + size_t(P) * 8);
  if (rows.wide && ids_fit) {
    for (int c = threadIdx.x; c < L; c += SELECT_THREADS) ids_sm[c] = rows.id_at(c);
    __syncthreads();
    lookup = ids_sm;
  } else if (rows.wide) {
    lookup = rows.id;          // a single list too long for shared memory is contiguous in global memory
  }
  block_topk_sorted(rows, k, P, ss, reinterpret_cast<uint64_t*>(sm_raw), rows.wide ? 0 : rows.base, out_score + size_t(q) * k,
                    out_id + size_t(q) * k, static_cast<uint64_t*>(nullptr), lookup);
}

// ------------------------------------------------------------------------------------------------
// Long rows (or many medium rows): one CTA per row would read the row once per radix pass from a single SM.  Instead a
// CTA streams a part of a row ONCE, keeps only the scores that beat its running threshold in a shared-memory candidate
// buffer (an exact radix select cuts the buffer back to k and raises the threshold whenever it fills), and writes its k
// best keys; a second kernel merges the lists of a row.  HBM traffic = the matrix, once.
//
// topk_stream_kernel (16-byte aligned rows).  The matrix is cut into trips of 4096 scores, numbered row-major, and every
// CTA of a one-wave grid (2 per SM) takes an equal, contiguous run of trips: no wave tail, and a run that covers
// hundreds of thousands of scores of a row has a threshold that only ~k ln(n) scores ever beat.  A run may cross row ends; the
// CTA then emits its list for the finished row and starts over.  Trips arrive through a shared-memory ring of 16 KB
// stages filled by bulk asynchronous copies (cp.async.bulk, completion on mbarriers) that thread 0 keeps three trips
// ahead of the filter, so the bytes in flight per SM do not depend on how the filter's barriers fall.  The filter is per
// thread: a thread whose 8 scores stay below the threshold does nothing.  The first trip of a row sets a threshold
// without streaming anything twice: the k-th largest of the 512 per-thread maxima of the trip has k scores at or above it.
// topk_slice_kernel: the same filter over plain loads and fixed slices, for rows that are not 16-byte aligned.
// ------------------------------------------------------------------------------------------------
constexpr int64_t SELECT_SLICE = int64_t(1) << 18;      // scores per slice CTA (1 MB)
constexpr int STREAM_THREADS = 512;                      // topk_stream_kernel: 16 warps, two CTAs per SM
constexpr int STREAM_PER_THREAD = 8;                     // scores per thread per trip (two 16-byte shared-memory loads)
constexpr int STREAM_TRIP = STREAM_THREADS * STREAM_PER_THREAD;      // 4096 scores = one 16 KB ring stage
constexpr int STREAM_STAGES = 4;
constexpr int SELECT_UNROLL = 16;                        // topk_slice_kernel: scores per thread per trip
constexpr int SELECT_TRIP = SELECT_THREADS * SELECT_UNROLL;
static_assert(SELECT_TRIP == STREAM_TRIP, "both kernels size the candidate buffer for the same trip");
constexpr int SELECT_CAP = LRAG_MAX_K + SELECT_TRIP;     // topk_slice_kernel's candidate keys: at most LRAG_MAX_K kept + one trip

struct SliceCands {
  const uint64_t* c; int n;
  template <class F> __device__ void operator()(F&& f) const {
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const uint64_t key = c[i]; if (key) f(key); }
  }
};

// State of a streaming CTA's filter: candidate buffer [trig + one trip], survivor buffer [P], running threshold (as a key
// and as a score).  `trig` = candidate count above which the buffer is cut back to its k best after a trip: a few times k,
// so that the threshold follows the scores seen so far closely (a compaction costs a few hundred instructions per warp,
// an insert ~15, and inserts fall as k / scores seen only while the threshold keeps up).
struct SliceFilter {
  uint64_t* cand; uint64_t* keep; SelectShared* ss; int* cnt; int* cnt2; int trig;
  unsigned long long thr_key; float thr_s;
  bool over = false;      // one of this thread's inserts of the current trip landed at or past `trig`

  // one score that reached the threshold: make its key, append it.  `col` = its column in the row, `cid` = the row's
  // column ids (or null; a negative id marks a column that does not take part)
  __device__ __forceinline__ void offer(float s, uint32_t col, const int64_t* cid) {
    const bool ok = !cid || cid[col] >= 0;
    const uint64_t key = make_key(s, col);                   // the tie field is the column; ids are looked up at the end
    if (ok && key > thr_key) {
      // one plain shared-memory atomic per hit: hits are rare and scattered over the warp, so the vote / popc / shuffle
      // sequence of a warp-aggregated increment (what atomicAdd compiles to) costs more than it saves
      uint32_t at;
      asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(at) : "r"(smem_u32(cnt)) : "memory");
      cand[at] = key;                                        // never past trig + one trip: see trip_barrier()
      over |= int(at) >= trig;
    }
  }
  // Score at or below every key >= pivot.  A pivot that the select left as a bin prefix (low bits zero) can decode to a
  // NaN when the k-th score is -inf; nothing compares >= NaN, so that case falls back to "no threshold".
  static __device__ __forceinline__ float floor_score(unsigned long long pivot) {
    const float s = key_score(pivot);
    return s == s ? s : -INFINITY;
  }
  // start of a new row (every thread, behind a __syncthreads)
  __device__ __forceinline__ void reset() {
    if (threadIdx.x == 0) *cnt = 0;
    thr_key = 0ull; thr_s = -INFINITY;
    __syncthreads();
  }
  // End of a trip, every thread: the CTA barrier, which also tells whether more than `trig` candidates are held now.  The
  // verdict travels with the barrier (some thread's insert index reached trig) instead of being read from `cnt`
  // afterwards: a warp that runs ahead into the next trip may already be inserting while a slower warp still decides.
  __device__ __forceinline__ bool trip_barrier() {
    const bool need = __syncthreads_or(over);
    over = false;
    return need;
  }
  // keep the k best (every thread, when trip_barrier() said so)
  __device__ __forceinline__ void compact(int k) {
    const int tid = threadIdx.x, n = *cnt;
    SliceCands cs{cand, n};
    const unsigned long long pivot = block_select_pivot(cs, k, *ss);
    if (tid == 0) *cnt2 = 0;
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
      const uint64_t key = cand[i];
      if (key && key >= pivot) keep[atomicAdd(cnt2, 1)] = key;
    }
    __syncthreads();
    const int m = *cnt2;
    for (int i = tid; i < m; i += blockDim.x) cand[i] = keep[i];
    if (tid == 0) *cnt = m;
    if (pivot > thr_key) { thr_key = pivot; thr_s = floor_score(pivot); }
    __syncthreads();
  }
  // the k best held (unsorted; 0 = empty)
  __device__ __forceinline__ void emit(int k, uint64_t* out) {
    const int tid = threadIdx.x, n = *cnt;
    SliceCands cs{cand, n};
    const unsigned long long pivot = block_select_pivot(cs, k, *ss);
    if (tid == 0) *cnt2 = 0;
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
      const uint64_t key = cand[i];
      if (key && key >= pivot) { const int at = atomicAdd(cnt2, 1); if (at < k) out[at] = key; }
    }
    __syncthreads();
    for (int i = *cnt2 + tid; i < k; i += blockDim.x) out[i] = 0;
    __syncthreads();
  }
};

// Balanced cut of an [nq, N] matrix into runs of trips, one run per CTA (see topk_stream_kernel).
struct StreamPlan {
  int tpr;         // trips per row
  int units;       // nq * tpr
  int grid;        // CTAs
  int seg_trips;   // trips per CTA (the last CTA may have fewer)
  int spr;         // list slots per row: a row is covered by at most this many CTAs
  int trig;        // candidate count that triggers a compaction (the candidate buffer holds trig + one trip)
};

struct OneKey {
  uint64_t key;
  template <class F> __device__ void operator()(F&& f) const { f(key); }
};

// 32-bit shared-window addressing for the per-trip instructions (no generic -> shared conversions in the loop)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 x;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(addr));
  return x;
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  long long t0 = 0;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    if (ok) return;
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > 20000000000LL) { printf("lrag: select ring wait timed out (block %d)\n", blockIdx.x); __trap(); }
  }
}

__global__ void __launch_bounds__(STREAM_THREADS)
topk_stream_kernel(const float* __restrict__ S, int64_t ld, int64_t N, int k, const int64_t* __restrict__ col_id, StreamPlan plan,
                   uint64_t* __restrict__ out_keys) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  float* ring = reinterpret_cast<float*>(sm_raw);                                     // [STREAM_STAGES][STREAM_TRIP]
  uint64_t* cand = reinterpret_cast<uint64_t*>(ring + STREAM_STAGES * STREAM_TRIP);   // [trig + STREAM_TRIP], then [P] survivors
  __shared__ SelectShared ss;
  __shared__ int cnt, cnt2;
  __shared__ __align__(8) uint64_t full[STREAM_STAGES];
  const int tid = threadIdx.x;
  const int u0 = blockIdx.x * plan.seg_trips;                                         // this CTA's run of trips [u0, u1)
  const int u1 = min(plan.units, u0 + plan.seg_trips);
  const int tpr = plan.tpr;
  const int n_last = int(N - int64_t(tpr - 1) * STREAM_TRIP);                         // scores in a row's last trip
  const bool boot_ok = k <= STREAM_THREADS && !col_id;
  SliceFilter f{cand, cand + plan.trig + STREAM_TRIP, &ss, &cnt, &cnt2, plan.trig, 0ull, -INFINITY};

  // Producer side (thread 0): trip (iq, itr) -> the next ring stage.  Whole 16-byte groups come by bulk copy; the <= 3
  // scores behind the last whole group of a row are read directly by the filter.
  int iq = u0 / tpr, itr = u0 - iq * tpr, istage = 0, iu = u0;
  auto issue = [&]() {
    const int n = itr == tpr - 1 ? n_last : STREAM_TRIP;
    const uint32_t bytes = uint32_t(n & ~3) * 4u;
    if (bytes) {
      mbar_arrive_expect_tx(&full[istage], bytes);
      bulk_copy_g2s(ring + istage * STREAM_TRIP, S + size_t(iq) * ld + size_t(itr) * STREAM_TRIP, bytes, &full[istage]);
    } else {
      mbar_arrive(&full[istage]);
    }
    if (++itr == tpr) { itr = 0; ++iq; }
    if (++istage == STREAM_STAGES) istage = 0;
    ++iu;
  };
  if (tid == 0) {
    for (int s = 0; s < STREAM_STAGES; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
    cnt = 0;
  }
  __syncthreads();
  if (tid == 0)
    while (iu < u1 && iu < u0 + STREAM_STAGES) issue();

  // Consumer side.  A thread's 8 scores of a trip: groups g = 0, 1 of four consecutive scores at trip offset (g * 512 + tid) * 4.
  const uint32_t ring_s = smem_u32(ring) + uint32_t(tid) * 16u, full_s = smem_u32(full);
  uint32_t stage = 0, parity = 0;
  int q = u0 / tpr, tr = u0 - q * tpr;
  for (int u = u0; u < u1;) {
    // the part of row q inside this run: trips [tr, tr_end)
    const int tr_end = min(tpr, tr + (u1 - u));
    const float* row = S + size_t(q) * ld;
    const int64_t* cid = col_id ? col_id + size_t(q) * N : nullptr;
    bool row_start = true;
    for (; tr < tr_end; ++tr, ++u) {
      float v[STREAM_PER_THREAD];
      mbar_wait_s(full_s + stage * 8u, parity);
      const uint32_t st = ring_s + stage * uint32_t(STREAM_TRIP * 4);
      if (tr != tpr - 1 || n_last == STREAM_TRIP) {
#pragma unroll
        for (int g = 0; g < STREAM_PER_THREAD / 4; ++g) {
          const float4 x = lds128(st + g * STREAM_THREADS * 16);
          v[4 * g] = x.x; v[4 * g + 1] = x.y; v[4 * g + 2] = x.z; v[4 * g + 3] = x.w;
        }
      } else {
        const int whole = n_last & ~3;
        const float qnan = __int_as_float(0x7fc00000);      // past the row's end: NaN never reaches a threshold
#pragma unroll
        for (int g = 0; g < STREAM_PER_THREAD / 4; ++g) {
          const int e = (g * STREAM_THREADS + tid) * 4;
          float4 x = make_float4(qnan, qnan, qnan, qnan);
          if (e + 3 < whole) x = lds128(st + g * STREAM_THREADS * 16);
          else if (e < n_last) {     // ragged end of the row: its last (partial) group comes straight from global memory
            const float* tail = row + size_t(tr) * STREAM_TRIP + e;
            x.x = __ldg(tail);
            if (e + 1 < n_last) x.y = __ldg(tail + 1);
            if (e + 2 < n_last) x.z = __ldg(tail + 2);
          }
          v[4 * g] = x.x; v[4 * g + 1] = x.y; v[4 * g + 2] = x.z; v[4 * g + 3] = x.w;
        }
      }
      float vmax = v[0];
#pragma unroll
      for (int i = 1; i < STREAM_PER_THREAD; ++i) vmax = fmaxf(vmax, v[i]);
      if (row_start && boot_ok) {
        // a threshold without streaming anything twice: the k-th largest of the 512 per-thread maxima of the row's first
        // trip has k scores of the row at or above it (a thread with no score inside the row counts as -inf)
        const OneKey mine{make_key(vmax == vmax ? vmax : -INFINITY, uint32_t(tid))};
        f.thr_s = SliceFilter::floor_score(block_select_pivot(mine, k, ss));
      }
      row_start = false;
      if (vmax >= f.thr_s) {
        // Which of the thread's eight scores reach the threshold, then one pass of the append code per hit of the warp's
        // busiest lane (one, as a rule) with the score picked by a select tree -- not one pass per SLOT: a warp runs the
        // body whenever any of its lanes has a hit, which early in a row is every trip.
        const uint32_t col0 = uint32_t(tr) * STREAM_TRIP + uint32_t(tid) * 4u;
        uint32_t hm = 0;
#pragma unroll
        for (int i = 0; i < STREAM_PER_THREAD; ++i) hm |= (v[i] >= f.thr_s ? 1u : 0u) << i;
        while (hm) {
          const int i = __ffs(hm) - 1;
          hm &= hm - 1;
          const float a0 = (i & 1) ? v[1] : v[0], a1 = (i & 1) ? v[3] : v[2], a2 = (i & 1) ? v[5] : v[4], a3 = (i & 1) ? v[7] : v[6];
          const float b0 = (i & 2) ? a1 : a0, b1 = (i & 2) ? a3 : a2;
          f.offer((i & 4) ? b1 : b0, col0 + uint32_t(i >> 2) * (STREAM_THREADS * 4) + uint32_t(i & 3), cid);
        }
      }
      const bool full_buf = f.trip_barrier();    // every thread holds its scores in registers: the stage is free
      if (tid == 0 && iu < u1) issue();
      if (full_buf) f.compact(k);
      if (++stage == STREAM_STAGES) { stage = 0; parity ^= 1; }
    }
    // end of the row or of the run: this CTA's list for row q goes to slot (CTA - first CTA that holds trips of q)
    f.emit(k, out_keys + (size_t(q) * plan.spr + (blockIdx.x - (q * tpr) / plan.seg_trips)) * k);
    f.reset();
    tr = 0; ++q;
  }
}

__global__ void __launch_bounds__(SELECT_THREADS)
topk_slice_kernel(const float* __restrict__ S, int64_t ld, int64_t N, int k, int P, const int64_t* __restrict__ col_id, int64_t slice_len,
                  int nslices, uint64_t* __restrict__ out_keys) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  uint64_t* cand = reinterpret_cast<uint64_t*>(sm_raw);           // [SELECT_CAP], then [P] survivors of a compaction
  __shared__ SelectShared ss;
  __shared__ int cnt, cnt2;
  const int q = blockIdx.y, sl = blockIdx.x, tid = threadIdx.x;
  const int64_t lo = int64_t(sl) * slice_len;
  const int n_slice = int((lo + slice_len < N ? lo + slice_len : N) - lo);
  const float* src = S + size_t(q) * ld + lo;
  const int64_t* cid = col_id ? col_id + size_t(q) * N : nullptr;
  const uint32_t tie0 = uint32_t(lo);
  SliceFilter f{cand, cand + SELECT_CAP, &ss, &cnt, &cnt2, SELECT_CAP - SELECT_TRIP, 0ull, -INFINITY};
  if (tid == 0) cnt = 0;
  __syncthreads();
  // one trip of scores per thread; the next trip's loads are issued before this one is filtered
  auto load_trip = [&](int base, float (&v)[SELECT_UNROLL]) {
#pragma unroll
    for (int u = 0; u < SELECT_UNROLL; ++u) {
      const int off = base + tid + u * SELECT_THREADS;
      v[u] = off < n_slice ? __ldg(src + off) : -INFINITY;
    }
  };
  float v[SELECT_UNROLL], vn[SELECT_UNROLL];
  load_trip(0, v);
  for (int base = 0; base < n_slice; base += SELECT_TRIP) {
    if (base + SELECT_TRIP < n_slice) load_trip(base + SELECT_TRIP, vn);
    float vmax = v[0];
#pragma unroll
    for (int u = 1; u < SELECT_UNROLL; ++u) vmax = fmaxf(vmax, v[u]);
    if (vmax >= f.thr_s) {
#pragma unroll
      for (int u = 0; u < SELECT_UNROLL; ++u) {
        const int off = base + tid + u * SELECT_THREADS;
        if (v[u] >= f.thr_s && off < n_slice) f.offer(v[u], tie0 + uint32_t(off), cid);
      }
    }
    if (f.trip_barrier()) f.compact(k);
#pragma unroll
    for (int u = 0; u < SELECT_UNROLL; ++u) v[u] = vn[u];
  }
  f.emit(k, out_keys + (size_t(q) * nslices + sl) * k);
}

struct KeyList {
  const uint64_t* keys; int n;
  template <class F> __device__ void operator()(F&& f) const {
    for (int i = threadIdx.x; i < n; i += SELECT_THREADS) { const uint64_t key = keys[i]; if (key) f(key); }
  }
};

__global__ void __launch_bounds__(SELECT_THREADS)
topk_merge_keys_kernel(const uint64_t* keys, int slots, int k, int P, int64_t id_base, float* out_score, int64_t* out_id, StreamPlan plan,
                       const int64_t* col_id, int64_t N) {
  extern __shared__ uint8_t sm_raw[];
  __shared__ SelectShared ss;
  const int q = blockIdx.x;
  int lists = slots;                           // fixed slices: every slot of the row was written
  if (plan.seg_trips > 0)                      // balanced runs: the CTAs that hold trips of row q wrote the first `lists` slots
    lists = ((q + 1) * plan.tpr - 1) / plan.seg_trips - (q * plan.tpr) / plan.seg_trips + 1;
  KeyList rows{keys + size_t(q) * slots * k, lists * k};
  block_topk_sorted(rows, k, P, ss, reinterpret_cast<uint64_t*>(sm_raw), id_base, out_score + size_t(q) * k,
                    out_id + size_t(q) * k, static_cast<uint64_t*>(nullptr), col_id ? col_id + size_t(q) * N : nullptr);
}

// ------------------------------------------------------------------------------------------------
// Many medium rows (a few thousand to ~10^5 scores each, e.g. 4096 x 20 000): one WARP per row.  A row that is a handful of
// trips long cannot amortise a CTA's per-row costs (threshold bootstrap, final select, merge launch: ~14 barriers each), and
// one CTA per row re-reads the row once per radix pass.  Here a warp streams its row once -- sixteen scores per lane per trip
// in four 16-byte loads -- behind a per-lane threshold filter, keeps candidates in its own slice of shared memory, cuts them
// back to the k best with a warp-level radix select whenever the slice fills (which also raises the threshold), and finally
// sorts its k keys and writes the row's result itself.  No CTA-wide barrier anywhere, no second kernel.
// ------------------------------------------------------------------------------------------------
constexpr int WROW_MAX_WARPS = 16;                       // rows (warps) per CTA: 4 ... 16, chosen per launch so that the grid is whole waves
constexpr int WROW_PER_LANE = 16;
constexpr int WROW_TRIP = 32 * WROW_PER_LANE;            // 512 scores per warp trip
constexpr int WROW_MAX_K = 256;

// pivot such that exactly k of keys[0, n) are >= pivot (n >= k); warp-level counterpart of block_select_pivot
__device__ __forceinline__ unsigned long long warp_select_pivot(const uint64_t* keys, int n, int k, int* hist) {
  const int lane = threadIdx.x & 31;
  unsigned long long prefix = 0, mask = 0;
  int rem = k;
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
#pragma unroll
    for (int j = 0; j < 8; ++j) hist[lane * 8 + j] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      const uint64_t key = keys[i];
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1);
    }
    __syncwarp();
    int c[8], local = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j] = hist[lane * 8 + j]; local += c[j]; }
    int incl = local;                                    // keys in this lane's bins and all higher bins
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < 32) incl += v;
    }
    const int above = incl - local;
    const bool own = above < rem && above + local >= rem;        // exactly one lane: n >= k keys match the prefix
    int bin = 0, nrem = 0;
    if (own) {
      int cum = above;
#pragma unroll
      for (int j = 7; j >= 0; --j) {
        if (cum < rem && cum + c[j] >= rem) { bin = lane * 8 + j; nrem = (cum + c[j] == rem) ? 0 : rem - cum; }
        cum += c[j];
      }
    }
    const int src = __ffs(__ballot_sync(0xffffffffu, own)) - 1;
    bin = __shfl_sync(0xffffffffu, bin, src);
    rem = __shfl_sync(0xffffffffu, nrem, src);
    prefix |= (unsigned long long)bin << shift;
    mask |= 0xffull << shift;
    __syncwarp();
    if (rem == 0) break;                                 // the whole bin is wanted: every key >= prefix is a winner
  }
  return prefix;
}

// keeps the keys >= pivot of keys[0, n) (in place, order preserved); returns how many
__device__ __forceinline__ int warp_keep_ge(uint64_t* keys, int n, unsigned long long pivot) {
  const int lane = threadIdx.x & 31;
  int wr = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    const uint64_t key = i < n ? keys[i] : 0ull;
    const bool keep = i < n && key >= pivot;
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    __syncwarp();                                        // the chunk is in registers before anything lands in it
    if (keep) keys[wr + __popc(m & ((1u << lane) - 1))] = key;
    wr += __popc(m);
  }
  __syncwarp();
  return wr;
}

__global__ void __launch_bounds__(WROW_MAX_WARPS * 32)
topk_warp_rows_kernel(const float* __restrict__ S, int64_t ld, int nq, int N, int k, int P, int cap, int64_t id_base,
                      float* __restrict__ out_score, int64_t* __restrict__ out_id) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + warp;
  if (q >= nq) return;                                   // warps are independent: no CTA barrier below
  const size_t per_warp = size_t(cap + WROW_TRIP) * 8 + 256 * 4 + 16;
  uint64_t* keys = reinterpret_cast<uint64_t*>(sm_raw + warp * per_warp);          // [cap + one trip]
  int* hist = reinterpret_cast<int*>(keys + cap + WROW_TRIP);                      // [256]
  int* cnt = hist + 256;
  const float* row = S + size_t(q) * ld;
  if (lane == 0) *cnt = 0;
  __syncwarp();
  unsigned long long thr_key = 0ull;
  float thr_s = -INFINITY;
  const float qnan = __int_as_float(0x7fc00000);         // past the row's end: NaN never reaches a threshold
  const uint32_t cnt_s = smem_u32(cnt);
  // one trip of the row into registers (scores past the row's end are NaN)
  auto load_trip = [&](int base, float (&v)[WROW_PER_LANE]) {
    if (base + WROW_TRIP <= N) {
#pragma unroll
      for (int g = 0; g < WROW_PER_LANE / 4; ++g) {
        const float4 x = __ldcs(reinterpret_cast<const float4*>(row + base + (g * 32 + lane) * 4));
        v[4 * g] = x.x; v[4 * g + 1] = x.y; v[4 * g + 2] = x.z; v[4 * g + 3] = x.w;
      }
    } else {
#pragma unroll
      for (int g = 0; g < WROW_PER_LANE / 4; ++g) {
        const int e = base + (g * 32 + lane) * 4;
        float4 x = make_float4(qnan, qnan, qnan, qnan);
        if (e + 3 < N) x = __ldcs(reinterpret_cast<const float4*>(row + e));
        else if (e < N) {
          x.x = __ldg(row + e);
          if (e + 1 < N) x.y = __ldg(row + e + 1);
          if (e + 2 < N) x.z = __ldg(row + e + 2);
        }
        v[4 * g] = x.x; v[4 * g + 1] = x.y; v[4 * g + 2] = x.z; v[4 * g + 3] = x.w;
      }
    }
  };
  // one trip: filter the lane's sixteen scores, append the hits, cut the slice back when it is full
  auto step = [&](const float (&v)[WROW_PER_LANE], int base) {
    // which of the lane's sixteen scores reach the threshold (NaN compares false)
    uint32_t hm = 0;
#pragma unroll
    for (int i = 0; i < WROW_PER_LANE; ++i) hm |= (v[i] >= thr_s ? 1u : 0u) << i;
    if (thr_s == -INFINITY && __all_sync(0xffffffffu, hm == 0xffffu)) {
      // no threshold yet and a whole trip inside the row: every score is a candidate, no need to count them one by one
      const int n0 = *reinterpret_cast<volatile int*>(cnt);
#pragma unroll
      for (int i = 0; i < WROW_PER_LANE; ++i)
        keys[n0 + i * 32 + lane] = make_key(v[i], uint32_t(base + ((i >> 2) * 32 + lane) * 4 + (i & 3)));
      __syncwarp();
      if (lane == 0) *cnt = n0 + WROW_TRIP;
      hm = 0;
    }
    // the warp runs this loop as often as its busiest lane has hits -- once, as a rule -- instead of once per slot
    while (hm) {
      const int i = __ffs(hm) - 1;
      hm &= hm - 1;
      // v[i] by a select tree (a register array cannot be indexed dynamically)
      const float a0 = (i & 1) ? v[1] : v[0], a1 = (i & 1) ? v[3] : v[2], a2 = (i & 1) ? v[5] : v[4], a3 = (i & 1) ? v[7] : v[6];
      const float a4 = (i & 1) ? v[9] : v[8], a5 = (i & 1) ? v[11] : v[10], a6 = (i & 1) ? v[13] : v[12], a7 = (i & 1) ? v[15] : v[14];
      const float b0 = (i & 2) ? a1 : a0, b1 = (i & 2) ? a3 : a2, b2 = (i & 2) ? a5 : a4, b3 = (i & 2) ? a7 : a6;
      const float c0 = (i & 4) ? b1 : b0, c1 = (i & 4) ? b3 : b2;
      const float x = (i & 8) ? c1 : c0;
      const uint64_t key = make_key(x, uint32_t(base + ((i >> 2) * 32 + lane) * 4 + (i & 3)));
      if (key > thr_key) {
        uint32_t at;
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(at) : "r"(cnt_s) : "memory");
        keys[at] = key;                                  // never past cap + one trip: the slice is cut back below
      }
    }
    __syncwarp();
    const int n = *reinterpret_cast<volatile int*>(cnt);
    if (n > cap) {
      // ---- the slice is full: keep the k best, their k-th is the new threshold ----
      const unsigned long long pivot = warp_select_pivot(keys, n, k, hist);
      const int m = warp_keep_ge(keys, n, pivot);
      if (lane == 0) *cnt = m;
      if (pivot > thr_key) { thr_key = pivot; const float ps = key_score(pivot); thr_s = ps == ps ? ps : -INFINITY; }
      __syncwarp();
    }
  };
  // the next two trips (4 KB per warp) are in flight while one is filtered; the buffers rotate by register moves (unrolling
  // the rotation away costs 40 more registers and a third of the resident warps)
  float v[WROW_PER_LANE], vn[WROW_PER_LANE], vnn[WROW_PER_LANE];
  load_trip(0, v);
  if (WROW_TRIP < N) load_trip(WROW_TRIP, vn);
  for (int base = 0; base < N; base += WROW_TRIP) {
    if (base + 2 * WROW_TRIP < N) load_trip(base + 2 * WROW_TRIP, vnn);
    step(v, base);
#pragma unroll
    for (int i = 0; i < WROW_PER_LANE; ++i) { v[i] = vn[i]; vn[i] = vnn[i]; }
  }
  // ---- the row's k best, sorted, straight to the output ----
  int n = *reinterpret_cast<volatile int*>(cnt);
  if (n > k) { const unsigned long long pivot = warp_select_pivot(keys, n, k, hist); n = warp_keep_ge(keys, n, pivot); }
  for (int i = n + lane; i < P; i += 32) keys[i] = 0ull;
  __syncwarp();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = lane; i < P / 2; i += 32) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint64_t a = keys[lo], b = keys[hi];
        if (desc ? (a < b) : (a > b)) { keys[lo] = b; keys[hi] = a; }
      }
      __syncwarp();
    }
  }
  float* os = out_score + size_t(q) * k;
  int64_t* oi = out_id + size_t(q) * k;
  for (int r = lane; r < k; r += 32) {
    const uint64_t key = r < n ? keys[r] : 0ull;
    os[r] = key ? key_score(key) : LRAG_PAD_SCORE;
    oi[r] = key ? id_base + int64_t(key_id(key)) : int64_t(-1);
  }
}

static int wrow_cap(int k) { return std::max(256, 2 * next_pow2(k)); }
static size_t wrow_smem_bytes(int k, int warps) { return size_t(warps) * (size_t(wrow_cap(k) + WROW_TRIP) * 8 + 256 * 4 + 16); }
// Rows (warps) per CTA: the choice that wastes the fewest warp slots over the waves the grid needs.  All rows of a call
// cost about the same, so a grid of 1.15 waves takes as long as one of 2; 4096 rows are one wave of 293 CTAs of 14 warps.
static int wrow_pick_warps(int nq, int k) {
  const int sms = sm_count();
  int best = 8; double best_eff = -1.0;
  for (int w = 4; w <= WROW_MAX_WARPS; ++w) {
    const int per_sm = int(std::min<size_t>(std::min<size_t>((227 * 1024) / (wrow_smem_bytes(k, w) + 1024), size_t(64 / w)), 32));
    if (per_sm < 1) continue;
    const int64_t slots = int64_t(sms) * per_sm, grid = (nq + w - 1) / w, waves = (grid + slots - 1) / slots;
    const double eff = double(nq) / double(waves * slots * w) * (per_sm * w >= 24 ? 1.0 : double(per_sm * w) / 24.0);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = w; }
  }
  return best;
}
// one warp per row pays off when there are enough rows to fill the machine with warps and a row is neither tiny nor long
// enough for the balanced runs of the streaming kernel
static bool wrow_applies(int nq, int64_t N, int k, bool aligned, bool has_col_id) {
  return aligned && !has_col_id && k <= WROW_MAX_K && N >= 2048 && N < 32 * STREAM_TRIP && nq >= 64;
}

// Slice length of the fixed-slice path: 2^18 scores when that already gives every SM several CTAs, shorter (down to 2^15, a
// multiple of the trip) when there are few rows.
static int64_t select_slice_len(int nq, int64_t N) {
  const int64_t want_ctas = 8 * int64_t(sm_count());
  int64_t len = SELECT_SLICE;
  while (len > (int64_t(1) << 15) && int64_t(nq) * ((N + len - 1) / len) < want_ctas) len >>= 1;
  return len;
}
static size_t slice_ws_bytes(int nq, int64_t N, int k) {
  if (N <= SELECT_SLICE) return 0;
  const int64_t len = select_slice_len(nq, N);
  return align_up(size_t(nq) * size_t((N + len - 1) / len) * size_t(k) * 8, 256);
}

static size_t stream_smem_bytes(int k, int trig) {
  return size_t(STREAM_STAGES) * STREAM_TRIP * 4 + size_t(trig + STREAM_TRIP + next_pow2(k)) * 8;
}

// Balanced runs pay off once a row is a few dozen trips long (a row costs a threshold bootstrap and a final select) and
// the matrix is a few MB; smaller problems stay with one CTA per row.
static bool stream_plan(int nq, int64_t N, int k, StreamPlan& p) {
  if (N < 32 * STREAM_TRIP || int64_t(nq) * N < (int64_t(1) << 22)) return false;
  const int64_t tpr = (N + STREAM_TRIP - 1) / STREAM_TRIP;
  const int64_t units = int64_t(nq) * tpr;
  if (units > (int64_t(1) << 30)) return false;
  static int trig_mult = 0;                      // candidates held before a compaction, in multiples of k (tuning knob, read once)
  if (!trig_mult) { const char* e = getenv("LRAG_SELECT_TRIG_MULT"); trig_mult = e && atoi(e) > 0 ? atoi(e) : 4; }
  p.trig = std::min(4096, std::max(256, trig_mult * k));
  const int per_sm = stream_smem_bytes(k, p.trig) + 3 * 1024 <= (227 * 1024) / 2 ? 2 : 1;      // CTAs that fit an SM's shared memory
  const int64_t slots = int64_t(per_sm) * sm_count();
  p.tpr = int(tpr);
  p.units = int(units);
  p.seg_trips = int((units + slots - 1) / slots);
  p.grid = int((units + p.seg_trips - 1) / p.seg_trips);
  p.spr = int(std::min<int64_t>(p.grid, (tpr + p.seg_trips - 1) / p.seg_trips + 1));
  return true;
}

size_t topk_select_ws_bytes(int nq, int64_t N, int k) {
  if (nq <= 0 || k <= 0) return 0;
  StreamPlan p;
  const size_t stream = stream_plan(nq, N, k, p) ? align_up(size_t(nq) * size_t(p.spr) * size_t(k) * 8, 256) : 0;
  return std::max(stream, slice_ws_bytes(nq, N, k));        // which path runs depends on the alignment of the rows
}

int launch_topk_select(const float* S, int64_t ld, int nq, int64_t N, int k, int64_t id_base, const int64_t* col_id,
                       float* out_score, int64_t* out_id, cudaStream_t stream, void* ws, size_t ws_bytes) {
  const int P = next_pow2(k);
  const size_t smem_slice = size_t(SELECT_CAP + P) * 8;
  static bool attr_set[LRAG_MAX_DEVICES] = {};      // function attributes are per device
  const int dev = device_slot();
  if (!attr_set[dev]) {
    LRAG_CHECK_CUDA(cudaFuncSetAttribute(topk_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         int(stream_smem_bytes(LRAG_MAX_K, 4096))));
    LRAG_CHECK_CUDA(cudaFuncSetAttribute(topk_slice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (SELECT_CAP + LRAG_MAX_K) * 8));
    attr_set[dev] = true;
  }
  uint64_t* keys = static_cast<uint64_t*>(ws);
  StreamPlan plan{};
  const bool aligned = (reinterpret_cast<uintptr_t>(S) & 15) == 0 && ld % 4 == 0;      // bulk copies need 16-byte aligned rows
  if (wrow_applies(nq, N, k, aligned, col_id != nullptr)) {
    static bool wrow_attr[LRAG_MAX_DEVICES] = {};
    if (!wrow_attr[dev]) {
      LRAG_CHECK_CUDA(cudaFuncSetAttribute(topk_warp_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(wrow_smem_bytes(WROW_MAX_K, WROW_MAX_WARPS))));
      wrow_attr[dev] = true;
    }
    prof_begin(stream, PROF_SELECT);
    const int wr_warps = wrow_pick_warps(nq, k);
    topk_warp_rows_kernel<<<(nq + wr_warps - 1) / wr_warps, wr_warps * 32, wrow_smem_bytes(k, wr_warps), stream>>>(
        S, ld, nq, int(N), k, P, wrow_cap(k), id_base, out_score, out_id);
    prof_end(stream);
    LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
    return LRAG_OK;
  }
  if (aligned && ws && stream_plan(nq, N, k, plan) && ws_bytes >= size_t(nq) * size_t(plan.spr) * size_t(k) * 8) {
    prof_begin(stream, PROF_SELECT);
    topk_stream_kernel<<<plan.grid, STREAM_THREADS, stream_smem_bytes(k, plan.trig), stream>>>(S, ld, N, k, col_id, plan, keys);
    prof_end(stream);
    LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
    topk_merge_keys_kernel<<<nq, SELECT_THREADS, select_smem_bytes(k), stream>>>(keys, plan.spr, k, P, id_base, out_score, out_id, plan, col_id, N);
    LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
    return LRAG_OK;
  }
  const size_t need = slice_ws_bytes(nq, N, k);
  if (need && ws && ws_bytes >= need && nq <= 65535) {
    const int64_t slice_len = select_slice_len(nq, N);
    const int nslices = int((N + slice_len - 1) / slice_len);
    prof_begin(stream, PROF_SELECT);
    topk_slice_kernel<<<dim3(nslices, nq), SELECT_THREADS, smem_slice, stream>>>(S, ld, N, k, P, col_id, slice_len, nslices, keys);
    prof_end(stream);
    LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
    topk_merge_keys_kernel<<<nq, SELECT_THREADS, select_smem_bytes(k), stream>>>(keys, nslices, k, P, id_base, out_score, out_id, StreamPlan{}, col_id, N);
    LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
    return LRAG_OK;
  }
  prof_begin(stream, PROF_SELECT);
  topk_select_kernel<<<nq, SELECT_THREADS, select_smem_bytes(k), stream>>>(S, ld, N, k, id_base, col_id, P, out_score,
                                                                           out_id);
  prof_end(stream);
  LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
  return LRAG_OK;
}

}  // namespace lrag

using namespace lrag;

extern "C" size_t lrag_topk_select_workspace_bytes(int nq, int64_t N, int k) {
  return topk_select_ws_bytes(nq, N, k);      // 0 for rows short enough for one CTA each
}

extern "C" int lrag_topk_select_f32(const float* S, int64_t ld, int nq, int64_t N, int k, int64_t id_base,
                                    const int64_t* col_id, float* out_score, int64_t* out_id, void* ws,
                                    size_t ws_bytes, lrag_stream_t stream) {
  LRAG_REQUIRE(initialised(), "lrag_init has not been called");
  LRAG_REQUIRE(nq > 0 && k > 0 && k <= LRAG_MAX_K, "topk_select: need nq > 0 and 1 <= k <= %d (nq=%d k=%d)", LRAG_MAX_K, nq, k);
  LRAG_REQUIRE(N >= 0 && N < (int64_t(1) << 32) - 1 && ld >= N, "topk_select: bad N=%lld ld=%lld", (long long)N, (long long)ld);
  LRAG_REQUIRE((S || N == 0) && out_score && out_id, "topk_select: null pointer");
  return launch_topk_select(S, ld, nq, N, k, col_id ? 0 : id_base, col_id, out_score, out_id,
                            static_cast<cudaStream_t>(stream), ws, ws_bytes);
}

static int launch_topk_merge(const float* score, const int64_t* id, int64_t s_stride, int64_t i_stride, int shards, int nq, int kin,
                             int k, float* out_score, int64_t* out_id, cudaStream_t stream) {
  const int P = next_pow2(k);
  // sorted keys [P] + (wide ids only) the row's ids [shards * kin]
  size_t smem = select_smem_bytes(k) + size_t(shards) * kin * 8;
  const bool ids_fit = smem <= 200 * 1024;
  LRAG_REQUIRE(ids_fit || shards == 1, "topk_merge: %d lists of %d entries do not fit one CTA's shared memory", shards, kin);
  if (!ids_fit) smem = select_smem_bytes(k);
  static size_t smem_set[LRAG_MAX_DEVICES] = {};      // function attributes are per device
  const int dev = device_slot();
  if (smem > 48 * 1024 && smem > smem_set[dev]) {
    LRAG_CHECK_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    smem_set[dev] = smem;
  }
  topk_merge_kernel<<<nq, SELECT_THREADS, smem, stream>>>(score, id, s_stride, i_stride, shards, kin, k, P, ids_fit ? 1 : 0, out_score, out_id);
  LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
  return LRAG_OK;
}

extern "C" int lrag_topk_merge(const float* score, const int64_t* id, int nq, int L, int k, float* out_score,
                               int64_t* out_id, lrag_stream_t stream) {
  LRAG_REQUIRE(initialised(), "lrag_init has not been called");
  LRAG_REQUIRE(nq > 0 && L >= 0 && k > 0 && k <= LRAG_MAX_K, "topk_merge: bad shape nq=%d L=%d k=%d", nq, L, k);
  LRAG_REQUIRE((score && id) || L == 0, "topk_merge: null pointer");
  LRAG_REQUIRE(out_score && out_id, "topk_merge: null output");
  return launch_topk_merge(score, id, 0, 0, 1, nq, L, k, out_score, out_id, static_cast<cudaStream_t>(stream));
}

extern "C" int lrag_topk_merge_shards(const float* score, const int64_t* id, int64_t score_shard_stride, int64_t id_shard_stride,
                                      int shards, int nq, int kin, int k, float* out_score, int64_t* out_id, lrag_stream_t stream) {
  LRAG_REQUIRE(initialised(), "lrag_init has not been called");
  LRAG_REQUIRE(nq > 0 && shards > 0 && kin > 0 && k > 0 && k <= LRAG_MAX_K, "topk_merge_shards: bad shape shards=%d nq=%d kin=%d k=%d",
               shards, nq, kin, k);
  LRAG_REQUIRE(score && id && out_score && out_id, "topk_merge_shards: null pointer");
  LRAG_REQUIRE(score_shard_stride >= int64_t(nq) * kin && id_shard_stride >= int64_t(nq) * kin,
               "topk_merge_shards: shard strides (%lld, %lld elements) are shorter than one shard's [nq, kin] block",
               (long long)score_shard_stride, (long long)id_shard_stride);
  return launch_topk_merge(score, id, score_shard_stride, id_shard_stride, shards, nq, kin, k, out_score, out_id,
                           static_cast<cudaStream_t>(stream));
}
