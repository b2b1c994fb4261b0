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

int launch_topk_select(const float* S, int64_t ld, int nq, int64_t N, int k, int64_t id_base, const int64_t* col_id,
                       float* out_score, int64_t* out_id, cudaStream_t stream) {
  const int P = next_pow2(k);
  topk_select_kernel<<<nq, SELECT_THREADS, select_smem_bytes(k), stream>>>(S, ld, N, k, id_base, col_id, P, out_score,
                                                                           out_id);
  LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
  return LRAG_OK;
}

}  // namespace lrag

using namespace lrag;

extern "C" size_t lrag_topk_select_workspace_bytes(int nq, int64_t N, int k) {
  (void)nq; (void)N; (void)k;
  return 0;
}

extern "C" int lrag_topk_select_f32(const float* S, int64_t ld, int nq, int64_t N, int k, int64_t id_base,
                                    const int64_t* col_id, float* out_score, int64_t* out_id, void* ws,
                                    size_t ws_bytes, lrag_stream_t stream) {
  (void)ws; (void)ws_bytes;
  LRAG_REQUIRE(initialised(), "lrag_init has not been called");
  LRAG_REQUIRE(nq > 0 && k > 0 && k <= LRAG_MAX_K, "topk_select: need nq > 0 and 1 <= k <= %d (nq=%d k=%d)", LRAG_MAX_K, nq, k);
  LRAG_REQUIRE(N >= 0 && N < (int64_t(1) << 32) - 1 && ld >= N, "topk_select: bad N=%lld ld=%lld", (long long)N, (long long)ld);
  LRAG_REQUIRE((S || N == 0) && out_score && out_id, "topk_select: null pointer");
  return launch_topk_select(S, ld, nq, N, k, col_id ? 0 : id_base, col_id, out_score, out_id,
                            static_cast<cudaStream_t>(stream));
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
