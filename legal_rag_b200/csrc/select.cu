// Row-wise exact top-k over materialised scores, and the k-way merge that follows the NCCL
// all-gather of per-shard candidate lists.
//   lrag_topk_select_f32  <- `sorted(range(N), key=scores[i], reverse=True)[:k]`
//                            (legalrag/retrieval/bm25_retriever.py:75 of the reference) and the ranking
//                            of MaxSim candidate scores (colbert_retriever.py:152, Searcher.search)
//   lrag_topk_merge       <- no reference counterpart (the reference is single-process)
#include "select.cuh"

namespace lrag {

struct RowScores {
  const float* s; const int64_t* col_id; int64_t n;
  template <class F> __device__ void operator()(F&& f) const {
    for (int64_t c = threadIdx.x; c < n; c += SELECT_THREADS) {
      uint32_t tie = uint32_t(c);
      if (col_id) {
        const int64_t id = col_id[c];
        if (id < 0) continue;
        tie = uint32_t(id);
      }
      f(make_key(s[c], tie));
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
                    out_score + size_t(q) * k, out_id + size_t(q) * k, static_cast<uint64_t*>(nullptr));
}

struct RowPairs {
  const float* s; const int64_t* id; int n;
  template <class F> __device__ void operator()(F&& f) const {
    for (int c = threadIdx.x; c < n; c += SELECT_THREADS) {
      const int64_t i = id[c];
      if (i >= 0) f(make_key(s[c], uint32_t(i)));
    }
  }
};

__global__ void __launch_bounds__(SELECT_THREADS)
topk_merge_kernel(const float* score, const int64_t* id, int L, int k, int P, float* out_score, int64_t* out_id) {
  extern __shared__ uint8_t sm_raw[];
  __shared__ SelectShared ss;
  const int q = blockIdx.x;
  RowPairs rows{score + size_t(q) * L, id + size_t(q) * L, L};
  block_topk_sorted(rows, k, P, ss, reinterpret_cast<uint64_t*>(sm_raw), 0, out_score + size_t(q) * k,
                    out_id + size_t(q) * k, static_cast<uint64_t*>(nullptr));
}

// ------------------------------------------------------------------------------------------------
// Long rows: one CTA per row would read the row once per radix pass from a single SM.  Instead the row is cut into
// slices; a slice CTA streams its scores ONCE (coalesced), keeps only those that beat its running threshold in a
// shared-memory candidate buffer (an exact radix select cuts the buffer back to k and raises the threshold whenever it
// fills), and writes its k best keys; a second kernel merges the slices' lists.  HBM traffic = the row, once.
// ------------------------------------------------------------------------------------------------
constexpr int64_t SELECT_SLICE = int64_t(1) << 18;      // scores per slice CTA (1 MB)
constexpr int SELECT_UNROLL = 16;                        // scores per thread per trip (four 16-byte loads when rows are aligned)
constexpr int SELECT_TRIP = SELECT_THREADS * SELECT_UNROLL;
constexpr int SELECT_CAP = LRAG_MAX_K + SELECT_TRIP;     // candidate keys: at most k survivors + one trip

struct SliceCands {
  const uint64_t* c; int n;
  template <class F> __device__ void operator()(F&& f) const {
    for (int i = threadIdx.x; i < n; i += SELECT_THREADS) f(c[i]);
  }
};

template <bool VEC>
__global__ void __launch_bounds__(SELECT_THREADS)
topk_slice_kernel(const float* __restrict__ S, int64_t ld, int64_t N, int k, int P, const int64_t* __restrict__ col_id, int64_t slice_len,
                  int nslices, uint64_t* __restrict__ out_keys) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  uint64_t* cand = reinterpret_cast<uint64_t*>(sm_raw);           // [SELECT_CAP]
  uint64_t* keep = cand + SELECT_CAP;                              // [P] survivors of a compaction
  __shared__ SelectShared ss;
  __shared__ int cnt, cnt2;
  const int q = blockIdx.y, sl = blockIdx.x;
  const int64_t lo = int64_t(sl) * slice_len;
  const int64_t hi = lo + slice_len < N ? lo + slice_len : N;
  const float* row = S + size_t(q) * ld;
  const int64_t* cid = col_id ? col_id + size_t(q) * N : nullptr;
  const int tid = threadIdx.x, lane = tid & 31;
  unsigned long long thr_key = 0ull;
  float thr_s = -INFINITY;
  if (tid == 0) cnt = 0;
  __syncthreads();
  // one trip of scores per thread; the next trip's loads are issued before this one is filtered
  auto load_trip = [&](int64_t base, float (&v)[SELECT_UNROLL]) {
    if (VEC) {
#pragma unroll
      for (int g = 0; g < SELECT_UNROLL / 4; ++g) {
        const int64_t c = base + (int64_t(g) * SELECT_THREADS + tid) * 4;
        float4 x = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (c + 3 < hi) x = __ldg(reinterpret_cast<const float4*>(row + c));
        else {
          if (c < hi) x.x = __ldg(row + c);
          if (c + 1 < hi) x.y = __ldg(row + c + 1);
          if (c + 2 < hi) x.z = __ldg(row + c + 2);
        }
        v[4 * g] = x.x; v[4 * g + 1] = x.y; v[4 * g + 2] = x.z; v[4 * g + 3] = x.w;
      }
    } else {
#pragma unroll
      for (int u = 0; u < SELECT_UNROLL; ++u) {
        const int64_t c = base + tid + int64_t(u) * SELECT_THREADS;
        v[u] = c < hi ? __ldg(row + c) : -INFINITY;
      }
    }
  };
  float v[SELECT_UNROLL], vn[SELECT_UNROLL];
  load_trip(lo, v);
  for (int64_t base = lo; base < hi; base += SELECT_TRIP) {
    if (base + SELECT_TRIP < hi) load_trip(base + SELECT_TRIP, vn);
    int64_t c0[SELECT_UNROLL / 4];
#pragma unroll
    for (int g = 0; g < SELECT_UNROLL / 4; ++g) c0[g] = base + (int64_t(g) * SELECT_THREADS + tid) * 4;
    // common case once the threshold has settled: nothing in this warp's 512 scores reaches it (one max tree, one vote)
    float vmax = v[0];
#pragma unroll
    for (int u = 1; u < SELECT_UNROLL; ++u) vmax = fmaxf(vmax, v[u]);
    const bool warp_hit = __any_sync(0xffffffffu, vmax >= thr_s);
#pragma unroll
    for (int u = 0; u < SELECT_UNROLL; ++u) {
      if (!warp_hit) break;
      const int64_t c = VEC ? c0[u / 4] + (u & 3) : base + tid + int64_t(u) * SELECT_THREADS;
      bool want = c < hi && v[u] >= thr_s;
      if (__any_sync(0xffffffffu, want)) {
        uint64_t key = 0;
        if (want) {
          uint32_t tie = uint32_t(c);
          if (cid) { const int64_t id = cid[c]; want = id >= 0; tie = uint32_t(id); }
          key = make_key(v[u], tie);
          want = want && key > thr_key;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, want);
        if (m) {
          int at = 0;
          if (lane == __ffs(m) - 1) at = atomicAdd(&cnt, __popc(m));
          at = __shfl_sync(0xffffffffu, at, __ffs(m) - 1);
          if (want) cand[at + __popc(m & ((1u << lane) - 1))] = key;       // never past SELECT_CAP: see the trigger below
        }
      }
    }
    __syncthreads();
    if (cnt > SELECT_CAP - SELECT_TRIP) {
      // the next trip could overflow: keep the k best, their worst becomes the threshold
      const int n = cnt;
      SliceCands cs{cand, n};
      const unsigned long long pivot = block_select_pivot(cs, k, ss);
      if (tid == 0) cnt2 = 0;
      __syncthreads();
      for (int i = tid; i < n; i += SELECT_THREADS) {
        const uint64_t key = cand[i];
        if (key >= pivot) keep[atomicAdd(&cnt2, 1)] = key;
      }
      __syncthreads();
      const int m = cnt2;
      for (int i = tid; i < m; i += SELECT_THREADS) cand[i] = keep[i];
      if (tid == 0) cnt = m;
      if (pivot > thr_key) { thr_key = pivot; thr_s = key_score(pivot); }
      __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < SELECT_UNROLL; ++u) v[u] = vn[u];
  }
  // this slice's k best (unsorted; 0 = empty)
  const int n = cnt;
  SliceCands cs{cand, n};
  const unsigned long long pivot = block_select_pivot(cs, k, ss);
  uint64_t* out = out_keys + (size_t(q) * nslices + sl) * k;
  if (tid == 0) cnt2 = 0;
  __syncthreads();
  for (int i = tid; i < n; i += SELECT_THREADS) {
    const uint64_t key = cand[i];
    if (key >= pivot) { const int at = atomicAdd(&cnt2, 1); if (at < k) out[at] = key; }
  }
  __syncthreads();
  for (int i = cnt2 + tid; i < k; i += SELECT_THREADS) out[i] = 0;
}

struct KeyList {
  const uint64_t* keys; int n;
  template <class F> __device__ void operator()(F&& f) const {
    for (int i = threadIdx.x; i < n; i += SELECT_THREADS) { const uint64_t key = keys[i]; if (key) f(key); }
  }
};

__global__ void __launch_bounds__(SELECT_THREADS)
topk_merge_keys_kernel(const uint64_t* keys, int L, int k, int P, int64_t id_base, float* out_score, int64_t* out_id) {
  extern __shared__ uint8_t sm_raw[];
  __shared__ SelectShared ss;
  const int q = blockIdx.x;
  KeyList lists{keys + size_t(q) * L, L};
  block_topk_sorted(lists, k, P, ss, reinterpret_cast<uint64_t*>(sm_raw), id_base, out_score + size_t(q) * k,
                    out_id + size_t(q) * k, static_cast<uint64_t*>(nullptr));
}

// Slice length for a problem: 2^18 scores when that already gives every SM several CTAs, shorter (down to 2^15, a multiple
// of the trip) when there are few rows.
static int64_t select_slice_len(int nq, int64_t N) {
  const int64_t want_ctas = 8 * int64_t(sm_count());
  int64_t len = SELECT_SLICE;
  while (len > (int64_t(1) << 15) && int64_t(nq) * ((N + len - 1) / len) < want_ctas) len >>= 1;
  return len;
}

size_t topk_select_ws_bytes(int nq, int64_t N, int k) {
  if (N <= SELECT_SLICE || nq <= 0 || k <= 0) return 0;
  const int64_t len = select_slice_len(nq, N);
  return align_up(size_t(nq) * size_t((N + len - 1) / len) * size_t(k) * 8, 256);
}

int launch_topk_select(const float* S, int64_t ld, int nq, int64_t N, int k, int64_t id_base, const int64_t* col_id,
                       float* out_score, int64_t* out_id, cudaStream_t stream, void* ws, size_t ws_bytes) {
  const int P = next_pow2(k);
  const size_t need = topk_select_ws_bytes(nq, N, k);
  if (need && ws && ws_bytes >= need && nq <= 65535) {
    const int64_t slice_len = select_slice_len(nq, N);
    const int nslices = int((N + slice_len - 1) / slice_len);
    const size_t smem = size_t(SELECT_CAP + P) * 8;
    static bool attr_set = false;
    if (!attr_set) {
      LRAG_CHECK_CUDA(cudaFuncSetAttribute(topk_slice_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (SELECT_CAP + LRAG_MAX_K) * 8));
      LRAG_CHECK_CUDA(cudaFuncSetAttribute(topk_slice_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (SELECT_CAP + LRAG_MAX_K) * 8));
      attr_set = true;
    }
    uint64_t* keys = static_cast<uint64_t*>(ws);
    const bool vec = (reinterpret_cast<uintptr_t>(S) & 15) == 0 && ld % 4 == 0;      // 16-byte loads need aligned rows
    prof_begin(stream, PROF_SELECT);
    if (vec) topk_slice_kernel<true><<<dim3(nslices, nq), SELECT_THREADS, smem, stream>>>(S, ld, N, k, P, col_id, slice_len, nslices, keys);
    else topk_slice_kernel<false><<<dim3(nslices, nq), SELECT_THREADS, smem, stream>>>(S, ld, N, k, P, col_id, slice_len, nslices, keys);
    prof_end(stream);
    LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
    topk_merge_keys_kernel<<<nq, SELECT_THREADS, select_smem_bytes(k), stream>>>(keys, nslices * k, k, P, id_base, out_score, out_id);
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

extern "C" int lrag_topk_merge(const float* score, const int64_t* id, int nq, int L, int k, float* out_score,
                               int64_t* out_id, lrag_stream_t stream) {
  LRAG_REQUIRE(initialised(), "lrag_init has not been called");
  LRAG_REQUIRE(nq > 0 && L >= 0 && k > 0 && k <= LRAG_MAX_K, "topk_merge: bad shape nq=%d L=%d k=%d", nq, L, k);
  LRAG_REQUIRE((score && id) || L == 0, "topk_merge: null pointer");
  LRAG_REQUIRE(out_score && out_id, "topk_merge: null output");
  const int P = next_pow2(k);
  topk_merge_kernel<<<nq, SELECT_THREADS, select_smem_bytes(k), static_cast<cudaStream_t>(stream)>>>(
      score, id, L, k, P, out_score, out_id);
  LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
  return LRAG_OK;
}
