// Block-level exact top-k over 64-bit selection keys: MSD radix select (8 bits per pass, early exit
// as soon as a bin boundary lands exactly on k) + bitonic sort of the winners.
// One CTA of SELECT_THREADS threads per row.  Keys are (ord32(score) << 32) | (~tie id), so a
// larger key is a better hit under (score desc, id asc); key 0 is "empty".
#pragma once
#include "common.cuh"

namespace lrag {

constexpr int SELECT_THREADS = 256;

// Barrier over the threads that take part in a block-level primitive: the whole CTA by default, or a
// named barrier over the first `n` threads when other warps of the CTA are busy elsewhere.
struct CtaBarrier {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
  __device__ __forceinline__ int threads() const { return blockDim.x; }
};
struct NamedBarrier {
  int id, n;
  __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
  __device__ __forceinline__ int threads() const { return n; }
};

struct SelectShared {
  int hist[256];
  unsigned long long prefix;
  unsigned long long mask;
  int remaining;   // > 0: still narrowing; 0: pivot final; -1: fewer than k candidates, keep all
  int nsel;
};

// Finds the pivot such that exactly k candidate keys are >= pivot (all of them when there are fewer
// than k).  ForEach: `template <class F> __device__ void operator()(F&& f) const` calls f(key) for
// every candidate this thread owns.  Every thread of the block (>= 256 threads) must call.
template <class ForEach, class Bar = CtaBarrier>
__device__ unsigned long long block_select_pivot(const ForEach& for_each, int k, SelectShared& sm, Bar bar = Bar()) {
  const int tid = threadIdx.x;
  bar();
  if (tid == 0) { sm.prefix = 0; sm.mask = 0; sm.remaining = k; sm.nsel = 0; }
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    if (tid < 256) sm.hist[tid] = 0;
    bar();
    if (sm.remaining <= 0) break;
    const unsigned long long prefix = sm.prefix, mask = sm.mask;
    for_each([&](uint64_t key) {
      if ((key & mask) == prefix) atomicAdd(&sm.hist[(key >> shift) & 255], 1);
    });
    bar();
    if (tid < 32) {
      // lane l owns bins [8l, 8l+8); find the bin holding the `remaining`-th largest key
      int c[8], local = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = sm.hist[tid * 8 + j]; local += c[j]; }
      int incl = local;  // suffix sum over lanes: keys in this lane's bins and all higher bins
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_down_sync(0xffffffffu, incl, o);
        if (tid + o < 32) incl += v;
      }
      const int above = incl - local;
      const int total = __shfl_sync(0xffffffffu, incl, 0);
      const int rem = sm.remaining;
      __syncwarp();
      if (total < rem) {
        if (tid == 0) { sm.prefix = 0; sm.mask = 0; sm.remaining = -1; }   // keep everything
      } else if (above < rem && above + local >= rem) {
        int cum = above;
#pragma unroll
        for (int j = 7; j >= 0; --j) {
          if (cum < rem && cum + c[j] >= rem) {
            sm.prefix = prefix | ((unsigned long long)(tid * 8 + j) << shift);
            sm.mask = mask | (0xffull << shift);
            // the whole bin is wanted: every key >= (prefix with zero low bits) is a winner
            sm.remaining = (cum + c[j] == rem || pass == 7) ? 0 : rem - cum;
          }
          cum += c[j];
        }
      }
    }
    bar();
  }
  const unsigned long long pivot = (sm.remaining < 0) ? 0ull : sm.prefix;
  bar();
  return pivot;
}

// Bitonic sort of P (power of two) keys in shared memory, largest first.
template <class Bar = CtaBarrier>
__device__ __forceinline__ void block_sort_desc(uint64_t* key, int P, Bar bar = Bar()) {
  const int tid = threadIdx.x;
  const int nthr = bar.threads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < P / 2; i += nthr) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint64_t a = key[lo], b = key[hi];
        if (desc ? (a < b) : (a > b)) { key[lo] = b; key[hi] = a; }
      }
      bar();
    }
  }
}

// Exact sorted top-k of the candidates `for_each` enumerates.  sel_key: shared, P = pow2 >= k.
// Returns (in every thread) the number of real hits written; rows are padded with
// (LRAG_PAD_SCORE, -1).  out_key, if non-null, receives the k winning keys (0 = empty).
// `lookup` == null: a key's tie field is a local id and the hit's id is id_base + that.
// `lookup` != null: the tie field is an index into `lookup` (ids of any width; 64-bit ids never pass through
// the 32-bit field): the hit's id is id_base + lookup[index], and hits of equal score are put in ascending id order
// afterwards (which of several equal scores make the cut at rank k is then decided by index, not by id).
template <class ForEach>
__device__ int block_topk_sorted(const ForEach& for_each, int k, int P, SelectShared& sm, uint64_t* sel_key,
                                 int64_t id_base, float* out_score, int64_t* out_id, uint64_t* out_key,
                                 const int64_t* lookup = nullptr) {
  const int tid = threadIdx.x;
  const unsigned long long pivot = block_select_pivot(for_each, k, sm);
  for (int i = tid; i < P; i += SELECT_THREADS) sel_key[i] = 0;
  __syncthreads();
  for_each([&](uint64_t key) {
    if (key >= pivot) {
      const int pos = atomicAdd(&sm.nsel, 1);
      if (pos < P) sel_key[pos] = key;
    }
  });
  __syncthreads();
  block_sort_desc(sel_key, P);
  const int n = min(min(sm.nsel, k), P);
  for (int r = tid; r < k; r += SELECT_THREADS) {
    const uint64_t key = (r < n) ? sel_key[r] : 0;
    if (out_key) out_key[r] = key;
    if (!out_score) continue;
    if (!key) { out_score[r] = LRAG_PAD_SCORE; out_id[r] = -1; continue; }
    if (!lookup) { out_score[r] = key_score(key); out_id[r] = id_base + int64_t(key_id(key)); continue; }
    // rank inside the run of equal scores by (id, index): runs are one hit long unless scores tie exactly
    const uint32_t so = uint32_t(key >> 32);
    const int64_t id = lookup[key_id(key)];
    int lo = r, hi = r;
    while (lo > 0 && uint32_t(sel_key[lo - 1] >> 32) == so) --lo;
    while (hi + 1 < n && uint32_t(sel_key[hi + 1] >> 32) == so) ++hi;
    int pos = lo;
    for (int j = lo; j <= hi; ++j) {
      if (j == r) continue;
      const int64_t idj = lookup[key_id(sel_key[j])];
      pos += (idj < id || (idj == id && j < r)) ? 1 : 0;
    }
    out_score[pos] = key_score(key);
    out_id[pos] = id_base + id;
  }
  __syncthreads();
  return n;
}

static inline int next_pow2(int x) { int p = 32; while (p < x) p <<= 1; return p; }
static inline size_t select_smem_bytes(int k) { return size_t(next_pow2(k)) * 8; }

}  // namespace lrag
