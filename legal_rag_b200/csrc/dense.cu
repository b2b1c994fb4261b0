// Dense channel: exact flat inner-product top-k (replaces faiss IndexFlatIP.search at
// legalrag/retrieval/dense_retriever.py:42 of the reference).
//
// dense_scan_kernel: persistent, warp-specialised tcgen05 GEMM  S = Q . X^T  whose epilogue never
// writes S.  Tile = 128 queries x 256 docs, K streamed in 64-element (128 B, one swizzle atom)
// blocks by TMA through a 4-stage mbarrier ring; fp32 accumulators double-buffered in TMEM
// (2 x 256 columns).  Epilogue warps read the accumulator with tcgen05.ld, compare each score with a
// per-query running threshold and append survivors as 64-bit keys to a per-(CTA, query) candidate
// buffer in global memory (L2 resident); a full buffer is cut back to its k best by the owning
// warp, which also raises the query's global threshold (atomicMax).  dense_finalize_kernel then
// selects the exact top-k of each query's surviving candidates.
//
// Tile order: t = doc_tile * QB + query_block, CTA c takes t = c, c+G, ... so that all query blocks
// of one doc tile are in flight together and X is fetched from HBM once (re-reads hit L2).
#include "common.cuh"
#include "select.cuh"

#include <cstdlib>

namespace lrag {

constexpr int DENSE_BM = 128;          // queries per tile (UMMA M)
constexpr int DENSE_BN = 256;          // docs per tile    (UMMA N)
constexpr int DENSE_BK = 64;           // bf16 elements per k-block = one 128 B swizzle row
constexpr int DENSE_STAGES = 4;
constexpr int DENSE_A_BYTES = DENSE_BM * DENSE_BK * 2;   // 16 KB
constexpr int DENSE_B_BYTES = DENSE_BN * DENSE_BK * 2;   // 32 KB
constexpr int DENSE_STAGE_BYTES = DENSE_A_BYTES + DENSE_B_BYTES;
constexpr int DENSE_THREADS = 192;     // warp0 TMA, warp1 MMA, warps 2..5 epilogue
constexpr int DENSE_SMEM = DENSE_STAGES * DENSE_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int DENSE_TMEM_COLS = 512;

struct DenseParams {
  int64_t N;
  int nq, d, k;
  int QB;            // query blocks
  int64_t DT;        // doc tiles
  int g;             // gcd(grid, QB): CTA c only ever sees query blocks == c (mod g)
  int slots;         // (QB / g) * 128 candidate buffers per CTA
  int cap;           // entries per candidate buffer
  int debug;         // 0 normal; 1 = TMA only (no MMA, no epilogue); 2 = TMA + MMA, epilogue skipped (timing experiments)
  int x_evict_first; // corpus tiles are loaded with an L2 evict-first hint (the scan shares the machine with the BM25 scan)
  uint32_t* thr;     // [QB*128] orderable k-th best score so far, per query
  int32_t* cnt;      // [grid * slots]
  uint64_t* buf;     // [grid * slots * cap]
  uint64_t* best;    // [nq * k] keys: each query's exact top-k over the epochs already tightened (0 = empty)
  int64_t t_begin, t_end;   // tile range of this launch (one epoch)
};

__global__ void dense_init_kernel(uint32_t* thr, int nthr, int32_t* cnt, int ncnt, uint64_t* best, int nbest) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nthr) thr[i] = LRAG_ORD_NEG_INF;
  if (i < ncnt) cnt[i] = 0;
  if (i < nbest) best[i] = 0;
}

__global__ void __launch_bounds__(DENSE_THREADS, 1)
dense_scan_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                  const DenseParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DENSE_STAGES * DENSE_STAGE_BYTES);
  uint64_t* full_bar = bars;                         // [STAGES]
  uint64_t* empty_bar = bars + DENSE_STAGES;         // [STAGES]
  uint64_t* tfull_bar = bars + 2 * DENSE_STAGES;     // [2]
  uint64_t* tempty_bar = bars + 2 * DENSE_STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * DENSE_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int KB = (p.d + DENSE_BK - 1) / DENSE_BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < DENSE_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, DENSE_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint64_t x_policy = l2_policy_evict_first();
      for (int64_t t = p.t_begin + blockIdx.x; t < p.t_end; t += gridDim.x) {
        const int qb = int(t % p.QB);
        const int64_t dt = t / p.QB;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * DENSE_STAGE_BYTES;
          uint8_t* sb = sa + DENSE_A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], DENSE_STAGE_BYTES);
          tma_load_2d(sa, &tmap_q, &full_bar[stage], kb * DENSE_BK, qb * DENSE_BM);
          if (p.x_evict_first) tma_load_2d_hint(sb, &tmap_x, &full_bar[stage], kb * DENSE_BK, int32_t(dt * DENSE_BN), x_policy);
          else tma_load_2d(sb, &tmap_x, &full_bar[stage], kb * DENSE_BK, int32_t(dt * DENSE_BN));
          if (++stage == DENSE_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(DENSE_BM, DENSE_BN);
      int stage = 0; uint32_t phase = 0;
      uint32_t it = 0;
      for (int64_t t = p.t_begin + blockIdx.x; t < p.t_end; t += gridDim.x, ++it) {
        const uint32_t as = it & 1, aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * DENSE_BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * DENSE_STAGE_BYTES);
          const uint32_t b_addr = a_addr + DENSE_A_BYTES;
          const uint64_t da = umma_desc_k_sw128(a_addr);
          const uint64_t db = umma_desc_k_sw128(b_addr);
          if (p.debug == 1) {
            mbar_arrive(&empty_bar[stage]);
            if (kb == KB - 1) mbar_arrive(&tfull_bar[as]);
          } else {
#pragma unroll
            for (int kk = 0; kk < DENSE_BK / 16; ++kk) {
              // advance 16 elements (32 B) along K inside the swizzle atom: +2 in 16 B units
              umma_bf16_ss(tmem_d, da + uint64_t(kk * 2), db + uint64_t(kk * 2), idesc,
                           (kb | kk) != 0 ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);            // smem slot reusable once these MMAs retire
            if (kb == KB - 1) umma_commit(&tfull_bar[as]);
          }
          if (++stage == DENSE_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue: threshold filter out of TMEM =====================
    // Per 32-column chunk the common case is "nothing beats the threshold": one tcgen05.ld, a max
    // tree and one vote.  Only lanes whose chunk max reaches the threshold walk their columns.
    const int quad = warp & 3;                      // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;
    const int cap = p.cap;
    uint32_t it = 0;
    // this tile's counter / threshold are fetched one tile ahead (their L2 latency would otherwise
    // sit on the critical path of every tile)
    int64_t t = p.t_begin + blockIdx.x;
    size_t sidx_n = 0; int q_n = 0; int cnt_n = 0; uint32_t thr_n = LRAG_ORD_NEG_INF;
    if (t < p.t_end) {
      const int qb = int(t % p.QB);
      q_n = qb * DENSE_BM + row;
      sidx_n = size_t(blockIdx.x) * p.slots + (qb / p.g) * DENSE_BM + row;
      cnt_n = p.cnt[sidx_n];
      thr_n = *reinterpret_cast<volatile uint32_t*>(&p.thr[q_n]);
    }
    for (; t < p.t_end; t += gridDim.x, ++it) {
      const int64_t dt = t / p.QB;
      const int q = q_n;
      const size_t sidx = sidx_n;
      const bool qvalid = q < p.nq;
      uint64_t* mybuf = p.buf + sidx * cap;
      int cnt = cnt_n;
      const int cnt0 = cnt;
      float thr = unord32(thr_n);
      const int64_t n0 = dt * DENSE_BN;
      const bool tail = n0 + DENSE_BN > p.N;        // only the last doc tile holds out-of-range columns
      const int64_t tn = t + gridDim.x;
      if (tn < p.t_end) {
        const int qb = int(tn % p.QB);
        q_n = qb * DENSE_BM + row;
        sidx_n = size_t(blockIdx.x) * p.slots + (qb / p.g) * DENSE_BM + row;
        if (sidx_n != sidx) cnt_n = p.cnt[sidx_n];
        thr_n = *reinterpret_cast<volatile uint32_t*>(&p.thr[q_n]);
      }

      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + as * DENSE_BN;
#pragma unroll 1
      for (int c = 0; c < (p.debug ? 0 : DENSE_BN / 32); ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
        float m8[4];
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          float m = __uint_as_float(v[g8 * 8]);
#pragma unroll
          for (int j = 1; j < 8; ++j) m = fmaxf(m, __uint_as_float(v[g8 * 8 + j]));
          m8[g8] = m;
        }
        const bool hit = qvalid && fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])) >= thr;
        if (__any_sync(0xffffffffu, hit)) {
          if (hit) {
            const int64_t nbase = n0 + c * 32;
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              if (m8[g8] >= thr) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float sc = __uint_as_float(v[g8 * 8 + j]);
                  const int64_t n = nbase + g8 * 8 + j;
                  if (sc >= thr && (!tail || n < p.N)) mybuf[cnt++] = make_key(sc, uint32_t(n));
                }
              }
            }
          }
          // a buffer that could overflow in the next chunk is cut back to its k best, warp-wide
          uint32_t m = __ballot_sync(0xffffffffu, cnt > cap - 32);
          while (m) {
            const int L = __ffs(m) - 1;
            m &= m - 1;
            uint64_t* bufL = reinterpret_cast<uint64_t*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(mybuf), L));
            const int nL = __shfl_sync(0xffffffffu, cnt, L);
            const uint64_t pivot = warp_keep_topk(bufL, nL, p.k, lane);
            if (lane == L) {
              cnt = p.k;
              const uint32_t o = uint32_t(pivot >> 32);
              atomicMax(&p.thr[q], o);
              thr = fmaxf(thr, unord32(o));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (cnt != cnt0) p.cnt[sidx] = cnt;
      if (tn < p.t_end) {
        if (sidx_n == sidx) cnt_n = cnt;                                   // same buffer again next tile
        if (q_n == q) thr_n = max(thr_n, ord32(thr));                      // keep a locally raised threshold
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, DENSE_TMEM_COLS);
  }
}

// ---- exact top-k over each query's surviving candidates ----
struct DenseSegments {
  const uint64_t* buf; const int32_t* cnt; const uint64_t* best;
  int grid, g, slots, cap, first_cta, slot, k;
  template <class F> __device__ void operator()(F&& f) const {
    for (int i = threadIdx.x; i < k; i += SELECT_THREADS) { const uint64_t key = best[i]; if (key) f(key); }
    // one warp per CTA segment, several segments in flight: the loop is a chain of dependent global
    // loads (count, then keys), so its latency is what a finalize costs at small query batches
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = first_cta + g * warp; c < grid; c += g * (SELECT_THREADS / 32)) {
      const size_t sidx = size_t(c) * slots + slot;
      const int n = cnt[sidx];
      const uint64_t* b = buf + sidx * cap;
      for (int i = lane; i < n; i += 32) f(b[i]);
    }
  }
};

// FINAL: write (score, id) rows.  Otherwise ("tighten", between epochs): fold every CTA's candidates
// into best[q], raise thr[q] to the exact k-th best score seen so far and empty the CTA buffers, so
// that the next epoch filters against the global threshold instead of each CTA's local one.
template <bool FINAL>
__global__ void __launch_bounds__(SELECT_THREADS)
dense_finalize_kernel(DenseParams p, int grid, int64_t id_base, int P, float* out_score, int64_t* out_id) {
  extern __shared__ uint8_t sm_raw[];
  __shared__ SelectShared ss;
  uint64_t* sel_key = reinterpret_cast<uint64_t*>(sm_raw);
  const int q = blockIdx.x;
  const int qb = q / DENSE_BM, row = q % DENSE_BM;
  uint64_t* best = p.best + size_t(q) * p.k;
  DenseSegments seg{p.buf, p.cnt, best, grid, p.g, p.slots, p.cap, qb % p.g, (qb / p.g) * DENSE_BM + row, p.k};
  if (FINAL) {
    block_topk_sorted(seg, p.k, P, ss, sel_key, id_base, out_score + size_t(q) * p.k, out_id + size_t(q) * p.k,
                      static_cast<uint64_t*>(nullptr));
  } else {
    const int n = block_topk_sorted(seg, p.k, P, ss, sel_key, 0, static_cast<float*>(nullptr),
                                    static_cast<int64_t*>(nullptr), best);
    if (threadIdx.x == 0 && n >= p.k) atomicMax(&p.thr[q], uint32_t(sel_key[p.k - 1] >> 32));
    for (int c = seg.first_cta + p.g * threadIdx.x; c < grid; c += p.g * SELECT_THREADS)
      p.cnt[size_t(c) * p.slots + seg.slot] = 0;
  }
}

static int gcd_int(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

struct DensePlan {
  int grid, QB, g, slots, cap; int64_t DT;
  size_t off_thr, off_cnt, off_best, off_buf, off_qpad, total;
  int64_t epoch0;    // tiles in the first epoch (a multiple of lcm(grid, QB)); later epochs grow 4x
};

static DensePlan dense_plan(int64_t N, int d, int nq, int k, int sms) {
  DensePlan pl;
  pl.QB = (nq + DENSE_BM - 1) / DENSE_BM;
  pl.DT = (N + DENSE_BN - 1) / DENSE_BN;
  int64_t tiles = pl.DT * pl.QB;
  pl.grid = int(tiles < sms ? (tiles > 0 ? tiles : 1) : sms);
  pl.g = gcd_int(pl.grid, pl.QB);
  pl.slots = (pl.QB / pl.g) * DENSE_BM;
  int cap = 2 * k; if (cap < k + 64) cap = k + 64;
  pl.cap = (cap + 63) / 64 * 64;
  if (pl.cap < 128) pl.cap = 128;
  pl.off_thr = 0;
  pl.off_cnt = align_up(size_t(pl.QB) * DENSE_BM * 4, 256);
  pl.off_best = pl.off_cnt + align_up(size_t(pl.grid) * pl.slots * 4, 256);
  pl.off_buf = pl.off_best + align_up(size_t(pl.QB) * DENSE_BM * k * 8, 256);
  pl.epoch0 = int64_t(pl.grid) / pl.g * pl.QB;
  pl.off_qpad = pl.off_buf + align_up(size_t(pl.grid) * pl.slots * pl.cap * 8, 256);
  // a query batch that does not fill its last 128-row block is copied into a zero-padded block: TMA boxes
  // that hang over the end of the tensor are several times slower than in-bounds ones
  pl.total = pl.off_qpad + (nq % DENSE_BM ? size_t(pl.QB) * DENSE_BM * size_t(d) * 2 : 0);
  return pl;
}

}  // namespace lrag

using namespace lrag;

// SMs the scan's persistent grid may use: all of them, or `max_ctas` of them when the scan shares the machine (partition.cu)
static int dense_sms(int max_ctas) { const int sms = sm_count(); return max_ctas > 0 && max_ctas < sms ? max_ctas : sms; }

extern "C" size_t lrag_dense_topk_workspace_bytes_part(int64_t N, int d, int nq, int k, int max_ctas) {
  if (N < 0 || nq <= 0 || k <= 0 || d <= 0 || max_ctas < 0) return 0;
  return dense_plan(N, d, nq, k, dense_sms(max_ctas)).total;
}

extern "C" size_t lrag_dense_topk_workspace_bytes(int64_t N, int d, int nq, int k) {
  return lrag_dense_topk_workspace_bytes_part(N, d, nq, k, 0);
}

extern "C" int lrag_dense_topk_bf16(const void* X, int64_t N, int d, const void* Q, int nq, int k,
                                    int64_t id_base, float* out_score, int64_t* out_id, void* ws,
                                    size_t ws_bytes, lrag_stream_t stream_) {
  return lrag_dense_topk_bf16_part(X, N, d, Q, nq, k, id_base, 0, out_score, out_id, ws, ws_bytes, stream_);
}

extern "C" int lrag_dense_topk_bf16_part(const void* X, int64_t N, int d, const void* Q, int nq, int k,
                                         int64_t id_base, int max_ctas, float* out_score, int64_t* out_id, void* ws,
                                         size_t ws_bytes, lrag_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LRAG_REQUIRE(max_ctas >= 0, "dense_topk: max_ctas=%d must be >= 0 (0 = every SM)", max_ctas);
  LRAG_REQUIRE(initialised(), "lrag_init has not been called");
  LRAG_REQUIRE(nq > 0 && k > 0 && k <= LRAG_MAX_K, "dense_topk: need nq > 0 and 1 <= k <= %d (nq=%d k=%d)", LRAG_MAX_K, nq, k);
  LRAG_REQUIRE(N >= 0 && N < (int64_t(1) << 31), "dense_topk: N=%lld out of range for one shard", (long long)N);
  LRAG_REQUIRE(d > 0 && d % 8 == 0, "dense_topk: d=%d must be a positive multiple of 8 (16-byte rows)", d);
  LRAG_REQUIRE(Q && out_score && out_id && (X || N == 0), "dense_topk: null pointer");
  LRAG_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Q) & 15) == 0,
               "dense_topk: X and Q must be 16-byte aligned");
  const DensePlan pl = dense_plan(N, d, nq, k, dense_sms(max_ctas));
  if (ws_bytes < pl.total || !ws) { set_error("dense_topk: workspace %zu < required %zu", ws_bytes, pl.total); return LRAG_ENOSPC; }

  DenseParams p;
  { const char* dbg = getenv("LRAG_DENSE_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
  { const char* ev = getenv("LRAG_DENSE_X_EVICT_FIRST"); p.x_evict_first = ev ? atoi(ev) : 0; }
  p.N = N; p.nq = nq; p.d = d; p.k = k; p.QB = pl.QB; p.DT = pl.DT; p.g = pl.g; p.slots = pl.slots; p.cap = pl.cap;
  uint8_t* w = static_cast<uint8_t*>(ws);
  p.thr = reinterpret_cast<uint32_t*>(w + pl.off_thr);
  p.cnt = reinterpret_cast<int32_t*>(w + pl.off_cnt);
  p.buf = reinterpret_cast<uint64_t*>(w + pl.off_buf);
  p.best = reinterpret_cast<uint64_t*>(w + pl.off_best);
  p.t_begin = 0; p.t_end = 0;

  const int nthr = pl.QB * DENSE_BM, ncnt = pl.grid * pl.slots, nbest = nq * k;
  int ninit = nthr > ncnt ? nthr : ncnt;
  if (nbest > ninit) ninit = nbest;
  dense_init_kernel<<<(ninit + 255) / 256, 256, 0, stream>>>(p.thr, nthr, p.cnt, ncnt, p.best, nbest);
  LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();

  const int P = next_pow2(k);
  if (N > 0) {
    CUtensorMap tq, tx;
    const void* Qmap = Q;
    uint64_t q_rows = uint64_t(nq);
    if (nq % DENSE_BM) {
      uint8_t* qpad = w + pl.off_qpad;
      q_rows = uint64_t(pl.QB) * DENSE_BM;
      LRAG_CHECK_CUDA(cudaMemsetAsync(qpad + size_t(nq) * d * 2, 0, (q_rows - nq) * size_t(d) * 2, stream));
      LRAG_CHECK_CUDA(cudaMemcpyAsync(qpad, Q, size_t(nq) * d * 2, cudaMemcpyDeviceToDevice, stream));
      Qmap = qpad;
    }
    int rc = make_tmap_bf16_2d(&tq, Qmap, q_rows, uint64_t(d), uint64_t(d), DENSE_BM, DENSE_BK);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tx, X, uint64_t(N), uint64_t(d), uint64_t(d), DENSE_BN, DENSE_BK);
    if (rc) return rc;
    static bool attr_set[LRAG_MAX_DEVICES] = {};      // function attributes are per device
    const int dev = device_slot();
    if (!attr_set[dev]) {
      LRAG_CHECK_CUDA(cudaFuncSetAttribute(dense_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DENSE_SMEM));
      attr_set[dev] = true;
    }
    // Epochs: [0, e0), [e0, 4 e0), [4 e0, 16 e0), ... ; between two epochs every query's threshold is
    // tightened to its exact k-th best so far.  An epoch boundary is a multiple of lcm(grid, QB) tiles,
    // which keeps each CTA on the same residue class of query blocks (its slot map) in every epoch.
    const int64_t num_tiles = pl.DT * pl.QB;
    int64_t t0 = 0, span = pl.epoch0;
    const bool epochs = !getenv("LRAG_DENSE_ONE_EPOCH");
    while (t0 < num_tiles) {
      int64_t t1 = t0 + span;
      if (!epochs || t1 + span / 2 >= num_tiles) t1 = num_tiles;     // do not leave a sliver for the last launch
      p.t_begin = t0; p.t_end = t1;
      prof_begin(stream, PROF_DENSE_SCAN);
      dense_scan_kernel<<<pl.grid, DENSE_THREADS, DENSE_SMEM, stream>>>(tq, tx, p);
      prof_end(stream);
      LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
      if (t1 < num_tiles) {
        dense_finalize_kernel<false><<<nq, SELECT_THREADS, select_smem_bytes(k), stream>>>(p, pl.grid, id_base, P, nullptr, nullptr);
        LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
      }
      span = (t1 - 0) * 3;      // next epoch is three times everything scanned so far (4x growth)
      t0 = t1;
    }
  }
  dense_finalize_kernel<true><<<nq, SELECT_THREADS, select_smem_bytes(k), stream>>>(p, pl.grid, id_base, P, out_score, out_id);
  LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
  return LRAG_OK;
}

// ---------------------------------------------------------------------------------------------
// CUDA-core cross-check of the same contract: fp32 FMA over the bf16 inputs, scores materialised.
// Test instrument for the tcgen05 path (small shapes only).
// ---------------------------------------------------------------------------------------------
namespace lrag {
int launch_topk_select(const float* S, int64_t ld, int nq, int64_t N, int k, int64_t id_base, const int64_t* col_id,
                       float* out_score, int64_t* out_id, cudaStream_t stream, void* ws = nullptr, size_t ws_bytes = 0);
size_t topk_select_ws_bytes(int nq, int64_t N, int k);

__global__ void __launch_bounds__(256)
dense_scores_ref_kernel(const __nv_bfloat16* __restrict__ X, int64_t N, int d, const __nv_bfloat16* __restrict__ Q,
                        float* __restrict__ S) {
  extern __shared__ float qs[];
  const int q = blockIdx.y;
  for (int i = threadIdx.x; i < d; i += blockDim.x) qs[i] = __bfloat162float(Q[size_t(q) * d + i]);
  __syncthreads();
  const int64_t n = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const __nv_bfloat16* x = X + size_t(n) * d;
  float acc = 0.f;
  for (int i = 0; i < d; ++i) acc = fmaf(qs[i], __bfloat162float(x[i]), acc);
  S[size_t(q) * N + n] = acc;
}
}  // namespace lrag

extern "C" size_t lrag_dense_topk_ref_workspace_bytes(int64_t N, int d, int nq, int k) {
  (void)d; (void)k;
  if (N < 0 || nq <= 0) return 0;
  return align_up(size_t(N) * size_t(nq) * 4 + 4, 256);
}

extern "C" int lrag_dense_topk_bf16_ref(const void* X, int64_t N, int d, const void* Q, int nq, int k,
                                        int64_t id_base, float* out_score, int64_t* out_id, void* ws,
                                        size_t ws_bytes, lrag_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LRAG_REQUIRE(initialised(), "lrag_init has not been called");
  LRAG_REQUIRE(nq > 0 && nq <= 65535 && k > 0 && k <= LRAG_MAX_K && N >= 0 && d > 0, "dense_topk_ref: bad shape");
  LRAG_REQUIRE(Q && out_score && out_id && (X || N == 0), "dense_topk_ref: null pointer");
  const size_t need = lrag_dense_topk_ref_workspace_bytes(N, d, nq, k);
  if (ws_bytes < need || !ws) { set_error("dense_topk_ref: workspace %zu < required %zu", ws_bytes, need); return LRAG_ENOSPC; }
  float* S = static_cast<float*>(ws);
  if (N > 0) {
    dim3 grid(unsigned((N + 255) / 256), unsigned(nq));
    dense_scores_ref_kernel<<<grid, 256, size_t(d) * 4, stream>>>(static_cast<const __nv_bfloat16*>(X), N, d,
                                                                  static_cast<const __nv_bfloat16*>(Q), S);
    LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
  }
  return launch_topk_select(S, N, nq, N, k, id_base, nullptr, out_score, out_id, stream);
}
