// Fusion: HybridRetriever._fuse + the min_final_score filter + the final slice, one CTA per query
// (legalrag/retrieval/hybrid_retriever.py:389-551, _minmax :24-30, _rrf_with_breakdown :33-56,
// filter :309-310, slice :384 of the reference; restated in oracle/fuse.py, which is pinned to
// outputs of the reference's own code in tests/golden/fuse_golden.json).
//
// Tiny and latency-bound: <= 3 * kc candidates per query live in shared memory for the whole pass
// (load -> per-channel min/max -> union by id with per-candidate channel lookups -> RRF min/max ->
// score -> filter -> bitonic sort by (score desc, id asc) -> top-k + breakdown).  Arithmetic is
// fp64 like the reference's Python floats; only the stored scores are fp32.
#include "common.cuh"

namespace lrag {

constexpr int FUSE_THREADS = 256;

struct FuseParams {
  const float* s[3]; const int64_t* i[3];
  int nq, kc, k, method, rrf_k, P;    // P = pow2 >= 3*kc
  double w[3]; double alpha; double min_final;
  float* out_score; int64_t* out_id; float* out_breakdown;
};

__device__ __forceinline__ double block_reduce(double v, bool is_max, double* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmax(v, other) : fmin(v, other);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = scratch[0];
  for (int w = 1; w < FUSE_THREADS / 32; ++w) r = is_max ? fmax(r, scratch[w]) : fmin(r, scratch[w]);
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(FUSE_THREADS)
fuse_kernel(const FuseParams p) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int kc = p.kc, n3 = 3 * kc, P = p.P;
  int64_t* ids = reinterpret_cast<int64_t*>(sm_raw);                   // [3*kc]
  double* rrf_tot = reinterpret_cast<double*>(ids + n3);               // [3*kc]
  double* wsum = rrf_tot + n3;                                         // [3*kc]
  double* score = wsum + n3;                                           // [P]  (sort key, -inf = dropped)
  float* sc = reinterpret_cast<float*>(score + P);                     // [3*kc]
  int* order = reinterpret_cast<int*>(sc + n3);                        // [P]
  uint8_t* owns = reinterpret_cast<uint8_t*>(order + P);               // [3*kc] entry owns its id
  __shared__ double scratch[FUSE_THREADS / 32];
  __shared__ int nvalid[3];
  __shared__ double lo[3], hi[3];

  const int q = blockIdx.x, tid = threadIdx.x;
  if (tid < 3) nvalid[tid] = 0;
  __syncthreads();
  // ---- load; valid entries of a channel form a prefix (id == -1 padding at the tail) ----
  for (int e = tid; e < n3; e += FUSE_THREADS) {
    const int ch = e / kc, pos = e % kc;
    int64_t id = -1; float s = 0.f;
    if (p.i[ch]) { id = p.i[ch][size_t(q) * kc + pos]; s = p.s[ch][size_t(q) * kc + pos]; }
    ids[e] = id; sc[e] = s;
    if (id >= 0) atomicMax(&nvalid[ch], pos + 1);
  }
  __syncthreads();
  // ---- per-channel min / max (hybrid_retriever.py:24-30) ----
  for (int ch = 0; ch < 3; ++ch) {
    double mn = INFINITY, mx = -INFINITY;
    for (int pos = tid; pos < nvalid[ch]; pos += FUSE_THREADS) {
      const double v = double(sc[ch * kc + pos]);
      mn = fmin(mn, v); mx = fmax(mx, v);
    }
    mn = block_reduce(mn, false, scratch);
    mx = block_reduce(mx, true, scratch);
    if (tid == 0) { lo[ch] = mn; hi[ch] = mx; }
  }
  __syncthreads();
  // ---- union by id: the first channel holding an id owns it ----
  double rmn = INFINITY, rmx = -INFINITY;
  for (int e = tid; e < P; e += FUSE_THREADS) {
    if (e < P) { score[e] = -INFINITY; order[e] = e; }
    if (e >= n3) continue;
    const int ch = e / kc, pos = e % kc;
    const int64_t id = ids[e];
    bool own = (id >= 0) && pos < nvalid[ch];
    int at[3] = {-1, -1, -1};
    if (own) {
      at[ch] = pos;
      for (int c = 0; c < 3 && own; ++c) {
        if (c == ch) continue;
        int found = -1;
        for (int j = 0; j < nvalid[c]; ++j) if (ids[c * kc + j] == id) { found = j; break; }
        if (found >= 0 && c < ch) own = false;
        at[c] = found;
      }
    }
    owns[e] = own ? 1 : 0;
    if (!own) { rrf_tot[e] = 0.0; wsum[e] = 0.0; continue; }
    double tot = 0.0, ws = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (at[c] < 0) continue;
      const double wr = (p.method == 2) ? p.w[c] : 1.0;
      tot += wr * (1.0 / double(p.rrf_k + at[c] + 1));
      const double range = hi[c] - lo[c];
      const double nv = (range < 1e-12) ? 0.0 : (double(sc[c * kc + at[c]]) - lo[c]) / range;
      ws += p.w[c] * nv;
    }
    rrf_tot[e] = tot; wsum[e] = ws;
    rmn = fmin(rmn, tot); rmx = fmax(rmx, tot);
  }
  rmn = block_reduce(rmn, false, scratch);
  rmx = block_reduce(rmx, true, scratch);
  const double rrange = rmx - rmn;
  // ---- score per method, min_final_score filter ----
  for (int e = tid; e < n3; e += FUSE_THREADS) {
    if (!owns[e]) continue;
    const double rn = (rrange < 1e-12) ? 0.0 : (rrf_tot[e] - rmn) / rrange;
    double s;
    if (p.method == 0) s = wsum[e];
    else if (p.method == 3) s = p.alpha * rn + (1.0 - p.alpha) * wsum[e];
    else s = rn;
    // rank on the value that is returned (fp32), so equal outputs are ordered by id
    if (s >= p.min_final) score[e] = double(float(s));
  }
  __syncthreads();
  // ---- bitonic sort of entry indices by (score desc, id asc); dropped entries sink ----
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < P / 2; i += FUSE_THREADS) {
        const int l = 2 * i - (i & (stride - 1)), h = l + stride;
        const int a = order[l], b = order[h];
        const double sa = score[a], sb = score[b];
        const int64_t ia = a < n3 ? ids[a] : -1, ib = b < n3 ? ids[b] : -1;
        const bool a_first = sa > sb || (sa == sb && ia >= 0 && (ib < 0 || ia < ib));
        const bool b_first = sb > sa || (sa == sb && ib >= 0 && (ia < 0 || ib < ia));
        const bool desc = ((l & size) == 0);
        if (desc ? b_first : a_first) { order[l] = b; order[h] = a; }
      }
      __syncthreads();
    }
  }
  // ---- top-k + breakdown ----
  for (int r = tid; r < p.k; r += FUSE_THREADS) {
    const int e = (r < P) ? order[r] : -1;
    const bool ok = e >= 0 && e < n3 && owns[e] && score[e] > -INFINITY;
    float* os = p.out_score + size_t(q) * p.k + r;
    int64_t* oi = p.out_id + size_t(q) * p.k + r;
    float* bd = p.out_breakdown ? p.out_breakdown + (size_t(q) * p.k + r) * 8 : nullptr;
    if (!ok) {
      *os = LRAG_PAD_SCORE; *oi = -1;
      if (bd) for (int j = 0; j < 8; ++j) bd[j] = 0.f;
      continue;
    }
    const int64_t id = ids[e];
    *os = float(score[e]); *oi = id;
    if (!bd) continue;
    const int ch = e / kc;
    double norm[3] = {0, 0, 0}, raw[3] = {0, 0, 0};
    for (int c = 0; c < 3; ++c) {
      int found = -1;
      if (c == ch) found = e % kc;
      else for (int j = 0; j < nvalid[c]; ++j) if (ids[c * kc + j] == id) { found = j; break; }
      if (found < 0) continue;
      const double range = hi[c] - lo[c];
      norm[c] = (range < 1e-12) ? 0.0 : (double(sc[c * kc + found]) - lo[c]) / range;
      raw[c] = ((p.method == 2) ? p.w[c] : 1.0) * (1.0 / double(p.rrf_k + found + 1));
    }
    const double tot = rrf_tot[e];
    const double rn = (rrange < 1e-12) ? 0.0 : (tot - rmn) / rrange;
    double contrib[3] = {0, 0, 0};
    const double mass = (p.method == 3) ? p.alpha * rn : ((p.method == 0) ? 0.0 : rn);
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
  int P = 32; while (P < 3 * kc) P <<= 1;
  p.P = P;
  p.w[0] = w_dense; p.w[1] = w_bm25; p.w[2] = w_colb;
  p.alpha = alpha; p.min_final = min_final;
  p.out_score = out_score; p.out_id = out_id; p.out_breakdown = out_breakdown;
  const size_t smem = size_t(3 * kc) * (8 + 8 + 8 + 4) + size_t(P) * (8 + 4) + size_t(3 * kc) + 16;
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    LRAG_CHECK_CUDA(cudaFuncSetAttribute(fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    smem_set = smem;
  }
  prof_begin(static_cast<cudaStream_t>(stream), PROF_FUSE);
  fuse_kernel<<<nq, FUSE_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(p);
  prof_end(static_cast<cudaStream_t>(stream));
  LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
  return LRAG_OK;
}
