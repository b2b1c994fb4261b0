// Fusion: HybridRetriever._fuse + the min_final_score filter + the final slice, one CTA per query
// (legalrag/retrieval/hybrid_retriever.py:389-551, _minmax :24-30, _rrf_with_breakdown :33-56,
// filter :309-310, slice :384 of the reference; restated in oracle/fuse.py, which is pinned to
// outputs of the reference's own code in tests/golden/fuse_golden.json).
//
// <= 3 * kc candidates per query live in shared memory for the whole pass:
//   load -> per-channel min/max (one fused reduction) -> id -> {position in each channel} through a
//   shared-memory hash table (the reference's dict lookups; a linear search here made the pass
//   quadratic in kc) -> RRF min/max -> score -> filter -> bitonic sort of packed 64-bit keys
//   (ord32(score) << 32 | ~id when every id fits 32 bits: a plain integer compare; else | ~entry with an int64 id
//   compare on equal scores) -> top-k + breakdown.
// Arithmetic is fp64 like the reference's Python floats; only the stored scores are fp32.
#include "common.cuh"

namespace lrag {

constexpr int FUSE_THREADS = 256;
constexpr int FUSE_WARPS = FUSE_THREADS / 32;
constexpr int FUSE_NONE = 0x7fffffff;

struct FuseParams {
  const float* s[3]; const int64_t* i[3];
  int nq, kc, k, method, rrf_k, nch, P, H;    // nch = channels up to the last one present; P = pow2 >= nch*kc (sort width), H = pow2 hash slots
  double w[3]; double alpha; double min_final;
  float* out_score; int64_t* out_id; float* out_breakdown;
};

// Block-wide min of v[0..NV/2) and max of v[NV/2..NV) in one pass (every thread gets the results back in v).
template <class T, int NV>
__device__ __forceinline__ void block_minmax(T (&v)[NV], T (*scratch)[NV]) {
#pragma unroll
  for (int j = 0; j < NV; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T other = __shfl_xor_sync(0xffffffffu, v[j], o);
      v[j] = j < NV / 2 ? min(v[j], other) : max(v[j], other);
    }
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int j = 0; j < NV; ++j) scratch[threadIdx.x >> 5][j] = v[j];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    T r = scratch[0][j];
#pragma unroll
    for (int w = 1; w < FUSE_WARPS; ++w) r = j < NV / 2 ? min(r, scratch[w][j]) : max(r, scratch[w][j]);
    v[j] = r;
  }
}

__device__ __forceinline__ uint32_t fuse_hash(int64_t id, int H) {
  return uint32_t((uint64_t(id) * 0x9E3779B97F4A7C15ull) >> 40) & uint32_t(H - 1);
}

__global__ void __launch_bounds__(FUSE_THREADS)
fuse_kernel(const FuseParams p) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int kc = p.kc, n3 = p.nch * kc, P = p.P, H = p.H;      // entries of absent trailing channels are not even laid out
  int64_t* ids = reinterpret_cast<int64_t*>(sm_raw);                                   // [nch*kc]
  double* rrf_tot = reinterpret_cast<double*>(ids + n3);                               // [3*kc]
  double* wsum = rrf_tot + n3;                                                         // [3*kc]
  unsigned long long* skey = reinterpret_cast<unsigned long long*>(wsum + n3);         // [P] sort keys (0 = dropped)
  unsigned long long* hkey = skey + P;                                                 // [H] hash slots: id, ~0 = empty
  int* hpos = reinterpret_cast<int*>(hkey + H);                                        // [H][3] first position of the id in each channel
  float* sc = reinterpret_cast<float*>(hpos + 3 * H);                                  // [3*kc]
  __shared__ double scratch[FUSE_WARPS][6];
  __shared__ int nvalid[3];

  const int q = blockIdx.x, tid = threadIdx.x;
  if (tid < 3) nvalid[tid] = 0;
  for (int h = tid; h < H; h += FUSE_THREADS) { hkey[h] = ~0ull; hpos[3 * h] = hpos[3 * h + 1] = hpos[3 * h + 2] = FUSE_NONE; }
  __syncthreads();
  // ---- load; valid entries of a channel form a prefix (id == -1 padding at the tail) ----
  bool wide_id = false;
  for (int e = tid; e < n3; e += FUSE_THREADS) {
    const int ch = e / kc, pos = e - ch * kc;
    int64_t id = -1; float s = 0.f;
    if (p.i[ch]) { id = p.i[ch][size_t(q) * kc + pos]; s = p.s[ch][size_t(q) * kc + pos]; }
    ids[e] = id; sc[e] = s;
    wide_id |= id > int64_t(0xffffffffll);
    // the last valid entry of a channel is followed by padding (or by the end of the list): few threads reach the atomic
    if (id >= 0 && (pos == kc - 1 || p.i[ch][size_t(q) * kc + pos + 1] < 0)) atomicMax(&nvalid[ch], pos + 1);
  }
  // Sort keys carry ~id in their low word when every id of the query fits 32 bits (the order (score desc, id asc) is then
  // a plain integer compare); otherwise ~entry, with an id lookup on equal scores.
  const bool wide = __syncthreads_or(wide_id);
  // ---- per-channel min / max (hybrid_retriever.py:24-30) and the id table, one pass ----
  float mm[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};     // fp32 inputs: their min / max are exact in fp32
  for (int e = tid; e < n3; e += FUSE_THREADS) {
    const int ch = e / kc, pos = e - ch * kc;
    if (pos >= nvalid[ch]) continue;
    const float v = sc[e];
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (c == ch) { mm[c] = fminf(mm[c], v); mm[3 + c] = fmaxf(mm[3 + c], v); }
    const int64_t id = ids[e];
    if (id < 0) continue;
    uint32_t h = fuse_hash(id, H);
    for (;;) {
      const unsigned long long prev = atomicCAS(&hkey[h], ~0ull, (unsigned long long)id);
      if (prev == ~0ull || prev == (unsigned long long)id) break;
      h = (h + 1) & uint32_t(H - 1);
    }
    atomicMin(&hpos[3 * h + ch], pos);
  }
  block_minmax<float, 6>(mm, reinterpret_cast<float (*)[6]>(&scratch[0][0]));
  const double lo[3] = {double(mm[0]), double(mm[1]), double(mm[2])}, hi[3] = {double(mm[3]), double(mm[4]), double(mm[5])};
  __syncthreads();            // table complete
  // ---- union by id: the first channel holding an id owns it (its first occurrence there) ----
  double rr[2] = {INFINITY, -INFINITY};
  for (int e = tid; e < P; e += FUSE_THREADS) {
    skey[e] = 0ull;
    if (e >= n3) continue;
    const int ch = e / kc, pos = e - ch * kc;
    const int64_t id = ids[e];
    rrf_tot[e] = 0.0; wsum[e] = __longlong_as_double(0x7ff8000000000000ll);       // NaN marks "not an owner"
    if (id < 0 || pos >= nvalid[ch]) continue;
    uint32_t h = fuse_hash(id, H);
    while (hkey[h] != (unsigned long long)id) h = (h + 1) & uint32_t(H - 1);
    const int at[3] = {hpos[3 * h], hpos[3 * h + 1], hpos[3 * h + 2]};
    bool own = at[ch] == pos;
#pragma unroll
    for (int c = 0; c < 3; ++c) if (c < ch && at[c] != FUSE_NONE) own = false;
    if (!own) continue;
    double tot = 0.0, ws = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (at[c] == FUSE_NONE) continue;
      const double wr = (p.method == 2) ? p.w[c] : 1.0;
      tot += wr * (1.0 / double(p.rrf_k + at[c] + 1));
      const double range = hi[c] - lo[c];
      const double nv = (range < 1e-12) ? 0.0 : (double(sc[c * kc + at[c]]) - lo[c]) / range;
      ws += p.w[c] * nv;
    }
    rrf_tot[e] = tot; wsum[e] = ws;
    rr[0] = fmin(rr[0], tot); rr[1] = fmax(rr[1], tot);
  }
  block_minmax<double, 2>(rr, reinterpret_cast<double (*)[2]>(&scratch[0][0]));
  const double rmn = rr[0], rrange = rr[1] - rr[0];
  // ---- score per method, min_final_score filter, sort key ----
  for (int e = tid; e < n3; e += FUSE_THREADS) {
    if (!(wsum[e] == wsum[e])) continue;
    const double rn = (rrange < 1e-12) ? 0.0 : (rrf_tot[e] - rmn) / rrange;
    double s;
    if (p.method == 0) s = wsum[e];
    else if (p.method == 3) s = p.alpha * rn + (1.0 - p.alpha) * wsum[e];
    else s = rn;
    // rank on the value that is returned (fp32), so equal outputs are ordered by id
    if (s >= p.min_final)      // + 0.0f: no -0
      skey[e] = ((unsigned long long)ord32(float(s) + 0.0f) << 32) | (unsigned long long)(0xffffffffu - (wide ? uint32_t(e) : uint32_t(ids[e])));
  }
  __syncthreads();
  // ---- bitonic sort by (score desc, id asc); dropped entries (key 0) sink ----
  auto before = [&](unsigned long long a, unsigned long long b) {
    if ((a ^ b) >> 32) return a > b;                                   // different scores (or one of them dropped)
    if (a == 0ull || b == 0ull) return false;
    return ids[0xffffffffu - uint32_t(a)] < ids[0xffffffffu - uint32_t(b)];       // same fp32 score: lower id first
  };
  // Pair i of a step touches elements l = 2i - (i & (stride - 1)) and l + stride: for stride <= 32 the 32 pairs of a
  // warp stay inside one aligned 64-element block, so consecutive steps with stride <= 32 only need a warp barrier.
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < P / 2; i += FUSE_THREADS) {
        const int l = 2 * i - (i & (stride - 1)), h = l + stride;
        const unsigned long long a = skey[l], b = skey[h];
        const bool desc = ((l & size) == 0);
        const bool swap = wide ? (desc ? before(b, a) : before(a, b)) : (desc ? b > a : a > b);
        if (swap) { skey[l] = b; skey[h] = a; }
      }
      const int next = stride > 1 ? stride >> 1 : size;       // stride of the step that follows (size = 2 * size / 2)
      if (stride > 32 || next > 32) __syncthreads(); else __syncwarp();
    }
  }
  __syncthreads();
  // ---- top-k + breakdown ----
  for (int r = tid; r < p.k; r += FUSE_THREADS) {
    const unsigned long long key = (r < P) ? skey[r] : 0ull;
    float* os = p.out_score + size_t(q) * p.k + r;
    int64_t* oi = p.out_id + size_t(q) * p.k + r;
    float* bd = p.out_breakdown ? p.out_breakdown + (size_t(q) * p.k + r) * 8 : nullptr;
    if (key == 0ull) {
      *os = LRAG_PAD_SCORE; *oi = -1;
      if (bd) for (int j = 0; j < 8; ++j) bd[j] = 0.f;
      continue;
    }
    int e = int(0xffffffffu - uint32_t(key));                   // wide ids: the entry; else the id itself
    const int64_t id = wide ? ids[e] : int64_t(0xffffffffu - uint32_t(key));
    *os = unord32(uint32_t(key >> 32)); *oi = id;
    if (!bd) continue;
    uint32_t h = fuse_hash(id, H);
    while (hkey[h] != (unsigned long long)id) h = (h + 1) & uint32_t(H - 1);
    if (!wide)                                                  // the owning entry: first occurrence in the first channel that lists the id
      e = hpos[3 * h] != FUSE_NONE ? hpos[3 * h] : (hpos[3 * h + 1] != FUSE_NONE ? kc + hpos[3 * h + 1] : 2 * kc + hpos[3 * h + 2]);
    double norm[3] = {0, 0, 0}, raw[3] = {0, 0, 0};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int found = hpos[3 * h + c];
      if (found == FUSE_NONE) continue;
      const double range = hi[c] - lo[c];
      norm[c] = (range < 1e-12) ? 0.0 : (double(sc[c * kc + found]) - lo[c]) / range;
      raw[c] = ((p.method == 2) ? p.w[c] : 1.0) * (1.0 / double(p.rrf_k + found + 1));
    }
    const double tot = rrf_tot[e];
    const double rn = (rrange < 1e-12) ? 0.0 : (tot - rmn) / rrange;
    double contrib[3] = {0, 0, 0};
    const double mass = (p.method == 3) ? p.alpha * rn : ((p.method == 0) ? 0.0 : rn);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (p.method == 0) contrib[c] = p.w[c] * norm[c];
      else if (p.method == 3) contrib[c] = (1.0 - p.alpha) * p.w[c] * norm[c];
      if (p.method != 0 && mass > 0.0 && tot > 1e-18) contrib[c] += mass * raw[c] / tot;
    }
    bd[0] = float(rn); bd[1] = float(wsum[e]);
    bd[2] = float(norm[0]); bd[3] = float(norm[1]); bd[4] = float(norm[2]);
    bd[5] = float(contrib[0]); bd[6] = float(contrib[1]); bd[7] = float(contrib[2]);
  }
}

}  // namespace lrag

using namespace lrag;

extern "C" int lrag_fuse_topk(const float* s_dense, const int64_t* i_dense, const float* s_bm25,
                              const int64_t* i_bm25, const float* s_colb, const int64_t* i_colb, int nq,
                              int kc, int k, int method, double w_dense, double w_bm25, double w_colb,
                              int rrf_k, double alpha, double min_final, float* out_score, int64_t* out_id,
                              float* out_breakdown, lrag_stream_t stream) {
  LRAG_REQUIRE(initialised(), "lrag_init has not been called");
  LRAG_REQUIRE(nq > 0 && kc > 0 && kc <= LRAG_MAX_K && k > 0 && k <= 3 * LRAG_MAX_K, "fuse_topk: bad shape nq=%d kc=%d k=%d", nq, kc, k);
  LRAG_REQUIRE(method >= 0 && method <= 3, "fuse_topk: unknown method %d", method);
  LRAG_REQUIRE((!i_dense) == (!s_dense) && (!i_bm25) == (!s_bm25) && (!i_colb) == (!s_colb), "fuse_topk: a channel needs both scores and ids");
  LRAG_REQUIRE(out_score && out_id, "fuse_topk: null output");
  FuseParams p;
  p.s[0] = s_dense; p.s[1] = s_bm25; p.s[2] = s_colb; p.i[0] = i_dense; p.i[1] = i_bm25; p.i[2] = i_colb;
  p.nq = nq; p.kc = kc; p.k = k; p.method = method; p.rrf_k = rrf_k;
  p.nch = i_colb ? 3 : (i_bm25 ? 2 : 1);
  int P = 32; while (P < p.nch * kc) P <<= 1;
  p.P = P;
  // open-addressing table: load factor <= 0.5; only three channels of kc > 682 (P = 4096) take H = P, load <= 0.75,
  // because 2 P slots would not fit in shared memory
  auto smem_for = [&](int H) { return size_t(p.nch * kc) * (8 + 8 + 8 + 4) + size_t(P) * 8 + size_t(H) * (8 + 12) + 16; };
  p.H = smem_for(2 * P) <= 227 * 1024 ? 2 * P : P;
  p.w[0] = w_dense; p.w[1] = w_bm25; p.w[2] = w_colb;
  p.alpha = alpha; p.min_final = min_final;
  p.out_score = out_score; p.out_id = out_id; p.out_breakdown = out_breakdown;
  const size_t smem = smem_for(p.H);
  static size_t smem_set[LRAG_MAX_DEVICES] = {};      // function attributes are per device
  const int dev = device_slot();
  if (smem > 48 * 1024 && smem > smem_set[dev]) {
    LRAG_CHECK_CUDA(cudaFuncSetAttribute(fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    smem_set[dev] = smem;
  }
  prof_begin(static_cast<cudaStream_t>(stream), PROF_FUSE);
  fuse_kernel<<<nq, FUSE_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(p);
  prof_end(static_cast<cudaStream_t>(stream));
  LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
  return LRAG_OK;
}
