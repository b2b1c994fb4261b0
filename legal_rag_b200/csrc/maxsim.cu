// ColBERT channel: MaxSim late interaction  score(q, d) = sum_i max_j <q_i, d_j>  over a candidate
// list per query (replaces the scoring inside `Searcher.search(query, k)` at
// legalrag/retrieval/colbert_retriever.py:152 of the reference; colbert_score semantics, see
// oracle/maxsim.py).  An exact full-corpus scan is the same call with cand = arange(Nd).
//
// HBM-bound gather (Ld*dim*2 bytes per candidate, 32 FLOP/B), so the kernel is organised around
// keeping TMA loads in flight; the contraction itself runs on tcgen05 because that frees the SM's
// issue slots for the epilogue, not because it is the bound.
//   * persistent CTAs, one per SM; work item = (query, chunk of <= 32 candidates);
//   * operand A (UMMA M = 128) is the query's token matrix [Lq <= 32, 128] replicated into the four
//     32-row groups, so that every TMEM lane quadrant holds the full result and any epilogue warp
//     can reduce any candidate; operand B (UMMA N = Ld) is the candidate's token tile [Ld, 128],
//     fetched by two TMA boxes straight into the 128B-swizzled K-major layout;
//   * one candidate per pipeline stage: smem stage s feeds TMEM accumulator s (128 lanes x Ld
//     fp32 columns); epilogue warp s reads lane = query token, columns = doc tokens with
//     tcgen05.ld, masks columns >= doclen, takes the row max in registers and the sum over query
//     tokens with one warp shuffle tree.  The [Lq, Ld] similarity matrix never leaves TMEM.
#include "common.cuh"
#include "select.cuh"

namespace lrag {

constexpr int MS_THREADS = 192;          // warp0 TMA, warp1 MMA, warps 2..5 epilogue
constexpr int MS_DIM = 128;
constexpr int MS_A_BYTES = 128 * MS_DIM * 2;   // 32 KB: two K-halves of [128 rows x 64]
constexpr int MS_A_BUFS = 2;
constexpr int MS_MAX_STAGES = 4;
constexpr int MS_CHUNK = 32;
constexpr float MS_PAD_FILL = -9999.0f;  // what colbert_score puts in padded doc-token slots

struct MaxsimParams {
  const int32_t* doclen; const int64_t* cand; float* out;
  int64_t Nd; int Ld, Lq, nq, C;
  int stages;        // smem stages == TMEM accumulators
  int col_stride;    // TMEM columns per accumulator (128 or 256)
  int chunks;        // work items per query
  int64_t items;
};

__global__ void __launch_bounds__(MS_THREADS, 1)
maxsim_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d,
              const MaxsimParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = p.Ld * MS_DIM * 2;
  uint8_t* a_buf = smem;                                   // [MS_A_BUFS][32 KB]
  uint8_t* b_buf = smem + MS_A_BUFS * MS_A_BYTES;          // [stages][stage_bytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_buf + size_t(p.stages) * stage_bytes);
  uint64_t* full_bar = bars;                               // [4] TMA -> MMA
  uint64_t* empty_bar = bars + 4;                          // [4] MMA -> TMA
  uint64_t* tfull_bar = bars + 8;                          // [4] MMA -> epilogue
  uint64_t* tempty_bar = bars + 12;                        // [4] epilogue -> MMA
  uint64_t* afull_bar = bars + 16;                         // [2]
  uint64_t* aempty_bar = bars + 18;                        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  volatile int* a_nv = reinterpret_cast<volatile int*>(bars + 21);   // [2] valid candidates of the item in A buffer ab; -1 = no more items

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NS = p.stages;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_d);
    for (int s = 0; s < 4; ++s) {
      mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1);
      mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) { mbar_init(&afull_bar[s], 1); mbar_init(&aempty_bar[s], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // this lane's candidate row of item `it` (-1: no such slot).  Every role loads the NEXT item's rows while it works on the
  // current one, so the load's latency never sits in front of an item.
  auto lane_row = [&](int64_t it) -> int64_t {
    if (it >= p.items) return -1;
    const int q = int(it / p.chunks);
    const int c0 = int(it % p.chunks) * MS_CHUNK;
    return (lane < min(MS_CHUNK, p.C - c0)) ? p.cand[size_t(q) * p.C + c0 + lane] : int64_t(-1);
  };

  if (warp == 0) {
    // ===================== TMA producer (whole warp reads candidate ids, lane 0 issues) ============
    // Skipped slots (row < 0 or past the store: padding, or candidates another shard owns) never enter the pipeline: the three
    // roles derive the same mask of valid slots from the candidate ids and count stages over valid slots only; an item
    // without a valid slot costs nothing.  (The first version pushed row 0 through the ring for a skipped slot, so a shard
    // that owns 1/8 of the candidates paid for all of them.)
    uint32_t n = 0, itc = 0;
    int64_t row_next = lane_row(blockIdx.x);
    for (int64_t it = blockIdx.x; it < p.items; it += gridDim.x) {
      const int q = int(it / p.chunks);
      const int64_t row = row_next;
      row_next = lane_row(it + gridDim.x);
      const uint32_t vm = __ballot_sync(0xffffffffu, row >= 0 && row < p.Nd);
      if (vm == 0) continue;
      const uint32_t ab = itc & 1;
      if (lane == 0) {
        mbar_wait_spin(&aempty_bar[ab], ((itc >> 1) & 1) ^ 1);
        a_nv[ab] = __popc(vm);                                   // read by the MMA thread behind the barrier
        mbar_arrive_expect_tx(&afull_bar[ab], MS_A_BYTES);
        uint8_t* a = a_buf + ab * MS_A_BYTES;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int r = 0; r < 4; ++r)
            tma_load_2d(a + h * (MS_A_BYTES / 2) + r * 4096, &tmap_q, &afull_bar[ab], h * 64, q * p.Lq);
      }
      for (uint32_t m = vm; m; m &= m - 1, ++n) {
        const int64_t rw = __shfl_sync(0xffffffffu, row, __ffs(m) - 1);
        if (lane == 0) {
          const uint32_t s = n % NS, use = n / NS;
          mbar_wait_spin(&empty_bar[s], (use & 1) ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
          uint8_t* b = b_buf + size_t(s) * stage_bytes;
          tma_load_2d(b, &tmap_d, &full_bar[s], 0, int32_t(rw * p.Ld));
          tma_load_2d(b + stage_bytes / 2, &tmap_d, &full_bar[s], 64, int32_t(rw * p.Ld));
        }
      }
      ++itc;
    }
    if (lane == 0) {                                             // end marker for the MMA thread
      const uint32_t ab = itc & 1;
      mbar_wait_spin(&aempty_bar[ab], ((itc >> 1) & 1) ^ 1);
      a_nv[ab] = -1;
      mbar_arrive(&afull_bar[ab]);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread; the producer tells it how many candidates an item holds) ==========
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, p.Ld);
      uint32_t n = 0;
      for (uint32_t itc = 0;; ++itc) {
        const uint32_t ab = itc & 1;
        mbar_wait_spin(&afull_bar[ab], (itc >> 1) & 1);
        const int nv = a_nv[ab];
        if (nv < 0) break;
        const uint32_t a_addr = smem_u32(a_buf + ab * MS_A_BYTES);
        for (int c = 0; c < nv; ++c, ++n) {
          const uint32_t s = n % NS, use = n / NS;
          mbar_wait_spin(&tempty_bar[s], (use & 1) ^ 1);
          mbar_wait_spin(&full_bar[s], use & 1);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(b_buf + size_t(s) * stage_bytes);
          const uint32_t tmem_d = tmem_base + s * p.col_stride;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t da = umma_desc_k_sw128(a_addr + h * (MS_A_BYTES / 2));
            const uint64_t db = umma_desc_k_sw128(b_addr + h * (stage_bytes / 2));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16_ss(tmem_d, da + uint64_t(kk * 2), db + uint64_t(kk * 2), idesc, (h | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
          umma_commit(&tfull_bar[s]);
        }
        umma_commit(&aempty_bar[ab]);
      }
    }
  } else {
    // ===================== epilogue: row max over doc tokens, sum over query tokens ================
    const int e = warp - 2;                        // accumulator this warp owns
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may read
    if (e < NS) {
      const int nchunk = (p.Ld + 31) / 32;
      uint32_t n = 0;
      int64_t row_next = lane_row(blockIdx.x);
      for (int64_t it = blockIdx.x; it < p.items; it += gridDim.x) {
        const int q = int(it / p.chunks);
        const int c0 = int(it % p.chunks) * MS_CHUNK;
        const int nc = min(MS_CHUNK, p.C - c0);
        const int64_t rowl = row_next;
        row_next = lane_row(it + gridDim.x);
        const bool ok = rowl >= 0 && rowl < p.Nd;
        const uint32_t vm = __ballot_sync(0xffffffffu, ok);
        if (e == 0 && lane < nc && !ok) p.out[size_t(q) * p.C + c0 + lane] = -INFINITY;       // skipped slots
        for (uint32_t left = vm; left; left &= left - 1, ++n) {
          if (int(n % NS) != e) continue;
          const int c = __ffs(left) - 1;
          const uint32_t use = n / NS;
          const int64_t row = __shfl_sync(0xffffffffu, rowl, c);
          const int dl = p.doclen ? min(p.doclen[row], p.Ld) : p.Ld;
          mbar_wait_spin(&tfull_bar[e], use & 1);
          tc_fence_after();
          const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + e * p.col_stride;
          float m = MS_PAD_FILL;
          for (int ch = 0; ch < nchunk; ++ch) {
            if (ch * 32 >= dl) break;              // warp-uniform: dl is the same in every lane
            uint32_t v[32];
            tmem_ld_32x32(taddr + ch * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (ch * 32 + j < dl) m = fmaxf(m, __uint_as_float(v[j]));
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[e]);
          float sum = (lane < p.Lq) ? m : 0.f;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          if (lane == 0) p.out[size_t(q) * p.C + c0 + c] = sum;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_topk_select(const float* S, int64_t ld, int nq, int64_t N, int k, int64_t id_base, const int64_t* col_id,
                       float* out_score, int64_t* out_id, cudaStream_t stream, void* ws = nullptr, size_t ws_bytes = 0);
size_t topk_select_ws_bytes(int nq, int64_t N, int k);

static int maxsim_launch(const void* D, const int32_t* doclen, int64_t Nd, int Ld, int dim, const void* Q, int nq, int Lq,
                         const int64_t* cand, int C, float* out, cudaStream_t stream) {
  LRAG_REQUIRE(initialised(), "lrag_init has not been called");
  LRAG_REQUIRE(dim == MS_DIM, "maxsim: dim=%d, only 128-d token vectors are supported", dim);
  LRAG_REQUIRE(Ld >= 16 && Ld <= 256 && Ld % 16 == 0, "maxsim: Ld=%d must be a multiple of 16 in [16, 256]", Ld);
  LRAG_REQUIRE(Lq >= 1 && Lq <= 32, "maxsim: Lq=%d must be in [1, 32]", Lq);
  LRAG_REQUIRE(nq > 0 && C > 0 && Nd > 0, "maxsim: empty problem (nq=%d C=%d Nd=%lld)", nq, C, (long long)Nd);
  LRAG_REQUIRE(Nd * int64_t(Ld) < (int64_t(1) << 31), "maxsim: token store of %lld x %d rows exceeds one shard", (long long)Nd, Ld);
  LRAG_REQUIRE(D && Q && cand && out, "maxsim: null pointer");
  LRAG_REQUIRE((reinterpret_cast<uintptr_t>(D) & 15) == 0 && (reinterpret_cast<uintptr_t>(Q) & 15) == 0,
               "maxsim: D and Q must be 16-byte aligned");
  MaxsimParams p;
  p.doclen = doclen; p.cand = cand; p.out = out; p.Nd = Nd; p.Ld = Ld; p.Lq = Lq; p.nq = nq; p.C = C;
  p.col_stride = Ld <= 128 ? 128 : 256;
  const int stage_bytes = Ld * MS_DIM * 2;
  int stages = 512 / p.col_stride;
  const int smem_budget = 227 * 1024 - 1024 - 256 - MS_A_BUFS * MS_A_BYTES;
  if (stages * stage_bytes > smem_budget) stages = smem_budget / stage_bytes;
  if (stages > MS_MAX_STAGES) stages = MS_MAX_STAGES;
  p.stages = stages;
  p.chunks = (C + MS_CHUNK - 1) / MS_CHUNK;
  p.items = int64_t(nq) * p.chunks;
  const size_t smem = size_t(MS_A_BUFS) * MS_A_BYTES + size_t(stages) * stage_bytes + 1024 + 256;

  CUtensorMap tq, td;
  int rc = make_tmap_bf16_2d(&tq, Q, uint64_t(nq) * Lq, MS_DIM, MS_DIM, 32, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&td, D, uint64_t(Nd) * Ld, MS_DIM, MS_DIM, uint32_t(Ld), 64);
  if (rc) return rc;
  LRAG_CHECK_CUDA(cudaFuncSetAttribute(maxsim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int grid = int(p.items < sm_count() ? p.items : sm_count());
  prof_begin(stream, PROF_MAXSIM);
  maxsim_kernel<<<grid, MS_THREADS, smem, stream>>>(tq, td, p);
  prof_end(stream);
  LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
  return LRAG_OK;
}

}  // namespace lrag

using namespace lrag;

extern "C" size_t lrag_maxsim_rerank_workspace_bytes(int nq, int C, int k) {
  (void)k;
  if (nq <= 0 || C <= 0) return 0;
  return align_up(size_t(nq) * size_t(C) * 4, 256);
}

extern "C" int lrag_maxsim_scores_bf16(const void* D, const int32_t* doclen, int64_t Nd, int Ld, int dim,
                                       const void* Q, int nq, int Lq, const int64_t* cand, int C,
                                       float* out_score, lrag_stream_t stream) {
  return maxsim_launch(D, doclen, Nd, Ld, dim, Q, nq, Lq, cand, C, out_score, static_cast<cudaStream_t>(stream));
}

extern "C" int lrag_maxsim_rerank_bf16(const void* D, const int32_t* doclen, int64_t Nd, int Ld, int dim,
                                       const void* Q, int nq, int Lq, const int64_t* cand, int C, int k,
                                       int64_t id_base, float* out_score, int64_t* out_id, void* ws,
                                       size_t ws_bytes, lrag_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LRAG_REQUIRE(k > 0 && k <= LRAG_MAX_K, "maxsim_rerank: need 1 <= k <= %d (k=%d)", LRAG_MAX_K, k);
  LRAG_REQUIRE(out_score && out_id, "maxsim_rerank: null output");
  const size_t need = lrag_maxsim_rerank_workspace_bytes(nq, C, k);
  if (ws_bytes < need || !ws) { set_error("maxsim_rerank: workspace %zu < required %zu", ws_bytes, need); return LRAG_ENOSPC; }
  float* S = static_cast<float*>(ws);
  int rc = maxsim_launch(D, doclen, Nd, Ld, dim, Q, nq, Lq, cand, C, S, stream);
  if (rc) return rc;
  // candidates are LOCAL rows; ties go to the lower row, returned ids are id_base + row
  return launch_topk_select(S, C, nq, C, k, id_base, cand, out_score, out_id, stream);
}
