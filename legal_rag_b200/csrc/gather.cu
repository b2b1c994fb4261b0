// Gathered inner products: score[q, c] = <Q[q], X[rows[q, c]]> for a short candidate list per query.
// Replaces the re-embedding + `_cosine_sim` loop of the reference's graph-expansion channel
// (legalrag/retrieval/graph_retriever.py:177-186: the <= graph_limit = 800 neighbour texts are embedded
// again and compared with the question one by one); here the neighbours are index rows, so their unit-norm
// embeddings are already resident in the dense corpus and one warp per (query, candidate) reads the row once.
// HBM-bound gather of C rows x d x 2 bytes per query; no tensor cores (one query row against scattered
// corpus rows is a GEMV, not a GEMM).
#include "common.cuh"

namespace lrag {

constexpr int GATHER_WARPS = 8;

__device__ __forceinline__ float dot8_bf16(const uint4& a, const uint4& b) {
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 fa = __bfloat1622float2(pa[j]), fb = __bfloat1622float2(pb[j]);
    s = fmaf(fa.x, fb.x, s);
    s = fmaf(fa.y, fb.y, s);
  }
  return s;
}

__global__ void __launch_bounds__(GATHER_WARPS * 32)
gather_scores_kernel(const __nv_bfloat16* __restrict__ X, int64_t N, int d, const __nv_bfloat16* __restrict__ Q, int nq,
                     const int64_t* __restrict__ rows, int C, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t pair = int64_t(blockIdx.x) * GATHER_WARPS + (threadIdx.x >> 5);
  if (pair >= int64_t(nq) * C) return;
  const int q = int(pair / C);
  const int64_t r = rows[pair];
  if (r < 0 || r >= N) { if (lane == 0) out[pair] = -INFINITY; return; }
  const uint4* x = reinterpret_cast<const uint4*>(X + r * d);
  const uint4* qv = reinterpret_cast<const uint4*>(Q + int64_t(q) * d);
  float s = 0.f;
  for (int i = lane; i < d / 8; i += 32) s += dot8_bf16(__ldg(x + i), __ldg(qv + i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[pair] = s;
}

}  // namespace lrag

using namespace lrag;

extern "C" int lrag_dense_gather_scores_bf16(const void* X, int64_t N, int d, const void* Q, int nq, const int64_t* rows,
                                             int C, float* out_score, lrag_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LRAG_REQUIRE(initialised(), "lrag_init has not been called");
  LRAG_REQUIRE(nq > 0 && C > 0, "dense_gather_scores: need nq > 0 and C > 0 (nq=%d C=%d)", nq, C);
  LRAG_REQUIRE(N >= 0 && d > 0 && d % 8 == 0, "dense_gather_scores: N=%lld, d=%d must be a positive multiple of 8", (long long)N, d);
  LRAG_REQUIRE(Q && rows && out_score && (X || N == 0), "dense_gather_scores: null pointer");
  LRAG_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Q) & 15) == 0,
               "dense_gather_scores: X and Q must be 16-byte aligned");
  const int64_t pairs = int64_t(nq) * C;
  gather_scores_kernel<<<unsigned((pairs + GATHER_WARPS - 1) / GATHER_WARPS), GATHER_WARPS * 32, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(X), N, d, static_cast<const __nv_bfloat16*>(Q), nq, rows, C, out_score);
  LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
  return LRAG_OK;
}
