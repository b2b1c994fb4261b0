// ColBERT channel, full-corpus mode for query batches: MaxSim of EVERY document against a batch of queries
// (the reference's ColBERT channel is a first-stage retriever over the whole corpus,
// legalrag/retrieval/colbert_retriever.py:152 / hybrid_retriever.py:299; PLAID approximates this scan).
//
// maxsim_kernel (maxsim.cu) gathers candidates per query: every query re-reads the token rows it scores,
// which is the right shape for a rerank (HBM-bound, 32 FLOP/B).  A batch that scans the whole store would
// re-read it once per query; this kernel instead runs ONE tensor-core contraction
//     [nq x 32 query-token rows, 128] . [Nd x Ld doc-token rows, 128]^T
// (persistent CTAs, the CTA's 4-query block resident in shared memory as the A operand, doc tiles streamed
// by TMA through a 3-stage ring, tcgen05 128 x 256 tiles, accumulators double-buffered in TMEM; the CTAs that
// hold different query blocks walk the doc tiles in step, so the store is read from HBM once) and a MaxSim epilogue: a tile row block of 32 TMEM lanes is one query's tokens,
// a run of Ld columns is one document; each lane takes the max over the document's (unmasked) columns,
// one shuffle tree sums over the query's tokens, lane 0 writes score[query, doc].  The [Lq, Ld]
// similarity matrices never leave TMEM; what is written is the [nq, Nd] per-document score matrix
// (1/4096 of the token-level products), ranked by lrag_topk_select_f32.
// Arithmetic intensity = nq * 32 / ... >= 2048 FLOP per store byte at 64 queries: tensor-bound.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "select.cuh"

namespace lrag {

constexpr int SC_BM = 128;            // 4 queries x 32 token rows (UMMA M)
constexpr int SC_BN = 256;            // doc-token rows per tile   (UMMA N)
constexpr int SC_BK = 64;             // one 128 B swizzle row of bf16
constexpr int SC_DIM = 128;
constexpr int SC_KB = SC_DIM / SC_BK;
constexpr int SC_STAGES = 3;          // doc tiles in flight (a stage = one whole 256 x 128 doc tile, 64 KB)
constexpr int SC_A_BYTES = SC_BM * SC_DIM * 2;        // 32 KB: the CTA's query block, resident for a whole pass
constexpr int SC_B_BYTES = SC_BN * SC_DIM * 2;        // 64 KB
#ifndef LRAG_SCAN_EPI_PER_QUAD
#define LRAG_SCAN_EPI_PER_QUAD 2
#endif
constexpr int SC_EPQ = LRAG_SCAN_EPI_PER_QUAD;        // epilogue warps per TMEM lane quadrant; each takes 256 / SC_EPQ columns of the tile
constexpr int SC_EPI_WARPS = 4 * SC_EPQ;
constexpr int SC_THREADS = 64 + 32 * SC_EPI_WARPS;    // warp0 TMA, warp1 MMA, the rest epilogue
constexpr int SC_WCOLS = SC_BN / SC_EPQ;              // columns per epilogue warp
constexpr int SC_WCHUNKS = SC_WCOLS / 32;             // 32-column TMEM loads per epilogue warp and tile
constexpr int SC_PART_BYTES = 4 * (SC_EPQ - 1) * 32 * 4;  // partial maxima of the warps that do not finish their document (the last warp of a quadrant never hands one over)
constexpr int SC_SMEM = SC_A_BYTES + SC_STAGES * SC_B_BYTES + 1024 /*align*/ + 256 /*barriers*/ + SC_PART_BYTES;
static_assert(SC_SMEM <= 227 * 1024, "the tiles, barriers and partial maxima must fit one CTA's shared memory");
constexpr float SC_PAD_FILL = -9999.0f;   // value of a masked (padding) doc token, as in maxsim.cu / the oracle

struct ScanParams {
  const int32_t* doclen;
  float* out;          // [nq, ld_out]
  int64_t Nd, ld_out;
  int Ld, Lq, nq, QB;
  int64_t DT;          // doc tiles
  int64_t units;       // work units = QB * nchunks, unit u = (chunk u / QB, query block u % QB)
  int64_t chunk_tiles; // doc tiles per chunk
  int debug;           // timing experiments: 1 = epilogue reads nothing (accumulators handed straight back)
};

// 32 accumulator words -> their maximum (and m) through three-input maxima (FMNMX3): 16 instructions in 4 dependent levels
__device__ __forceinline__ float scan_max32(const uint32_t (&vv)[32], float m) {
  float l1[11];
#pragma unroll
  for (int g = 0; g < 10; ++g)
    l1[g] = fmaxf(fmaxf(__uint_as_float(vv[3 * g]), __uint_as_float(vv[3 * g + 1])), __uint_as_float(vv[3 * g + 2]));
  l1[10] = fmaxf(__uint_as_float(vv[30]), __uint_as_float(vv[31]));
  const float a0 = fmaxf(fmaxf(l1[0], l1[1]), l1[2]), a1 = fmaxf(fmaxf(l1[3], l1[4]), l1[5]);
  const float a2 = fmaxf(fmaxf(l1[6], l1[7]), l1[8]), a3 = fmaxf(fmaxf(l1[9], l1[10]), m);
  return fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
}

// The same for a chunk whose document ends inside it: only the first `rem` (1 .. 31, warp-uniform) words count.  Whole groups of
// eight go through max trees; the one group that holds the document's end selects its words against the pad value first --
// independent selects and a tree, not a 32-deep chain of predicated maxima through m.
__device__ __forceinline__ float scan_max32_masked(const uint32_t (&vv)[32], float m, int rem) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (rem > 8 * g) {
      float x[8];
      if (rem >= 8 * g + 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(vv[8 * g + j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = (8 * g + j < rem) ? __uint_as_float(vv[8 * g + j]) : SC_PAD_FILL;
      }
      const float a = fmaxf(fmaxf(x[0], x[1]), x[2]), b = fmaxf(fmaxf(x[3], x[4]), x[5]), c = fmaxf(fmaxf(x[6], x[7]), m);
      m = fmaxf(fmaxf(a, b), c);
    }
  }
  return m;
}

// Epilogue of one warp: SC_EPQ warps per TMEM lane quadrant (= query inside the block); warp `slot` of a quadrant takes columns
// [slot * SC_WCOLS, (slot + 1) * SC_WCOLS) of every tile.  A document longer than that (LD = 256) spans two warps: their partial
// maxima meet in shared memory and the warp that holds the document's last columns finishes it.  LD (32, 64, 128 or 256) and
// HAS_DL (per-document lengths given: tokens past a document's length are masked) are compile-time: a warp executes ~110
// instructions per tile in the unmasked LD = 128 instance against ~290 with run-time column maps and mask tests, and the
// epilogue of a tile is one warp's dependent instruction stream (tools/micro/tmem_ld.cu: TMEM read-back itself delivers
// 730-940 B/clk/SM, 6-7x what is needed).
template <int LD, bool HAS_DL>
__device__ __forceinline__ void scan_epilogue(const ScanParams& p, uint32_t tmem_base, uint64_t* tfull_bar, uint64_t* tempty_bar,
                                              float* part, int quad, int slot, int lane) {
  constexpr int DPT = SC_BN / LD;                              // documents per tile
  constexpr int SPAN = LD > SC_WCOLS ? LD / SC_WCOLS : 1;      // warps that share one document
  static_assert(SPAN <= 2 || SC_EPQ > 2, "partial hand-over below assumes the finisher reads its predecessors' slots");
  const int quad_bar = quad + 1;                               // named barrier of the quadrant's warps
  const int c0 = slot * SC_WCHUNKS;
  uint32_t it = 0;
  if (p.debug == 1) {
    // timing experiment (LRAG_SCAN_DEBUG=1): the mainloop alone, accumulators handed straight back
    for (int64_t u = blockIdx.x; u < p.units; u += gridDim.x) {
      const int64_t dt0 = (u / p.QB) * p.chunk_tiles, dt1 = min(p.DT, dt0 + p.chunk_tiles);
      for (int64_t dt = dt0; dt < dt1; ++dt, ++it) {
        mbar_wait(&tfull_bar[it & 1], (it >> 1) & 1);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[it & 1]);
      }
    }
    return;
  }
  for (int64_t u = blockIdx.x; u < p.units; u += gridDim.x) {
    const int qb = int(u % p.QB);
    const int64_t dt0 = (u / p.QB) * p.chunk_tiles;
    const int64_t dt1 = min(p.DT, dt0 + p.chunk_tiles);
    const int q = qb * 4 + quad;
    float* out_row = p.out + size_t(q < p.nq ? q : 0) * p.ld_out;
    // document lengths of a tile (lane d holds document d's), fetched one tile ahead: when the epilogue is what paces the
    // kernel nothing hides a load issued at the top of the tile it is needed in
    auto load_dl = [&](int64_t t) {
      int dl = 0;
      if (HAS_DL && lane < DPT && t < dt1) {
        const int64_t doc = t * DPT + lane;
        dl = doc < p.Nd ? min(__ldg(p.doclen + doc), LD) : 0;
      }
      return dl;
    };
    int dl_next = load_dl(dt0);
    for (int64_t dt = dt0; dt < dt1; ++dt, ++it) {
      const int64_t doc0 = dt * DPT;
      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      const int dl_mine = dl_next;
      dl_next = load_dl(dt + 1);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + as * SC_BN;
      float m = SC_PAD_FILL;
      // the next chunk's TMEM load is in flight while this one is reduced.  (Requesting all of a warp's columns at once and
      // waiting once was measured: 86.5 ms against 67.1 ms for 64 queries x 1M documents.)
      uint32_t v[2][32];
      // chunk i of this warp: columns (c0 + i) * 32 ...: document `din` of the tile, its tokens off .. off + 31
      auto din_of = [&](int i) { return ((c0 + i) * 32) / LD; };
      auto off_of = [&](int i) { return LD > SC_WCOLS ? ((c0 + i) * 32) % LD : (i * 32) % LD; };   // LD <= SC_WCOLS: the same for both slots
      int dl_c[SC_WCHUNKS];
#pragma unroll
      for (int i = 0; i < SC_WCHUNKS; ++i)
        dl_c[i] = HAS_DL ? __shfl_sync(0xffffffffu, dl_mine, din_of(i)) : LD;
      tmem_ld_32x32(taddr + c0 * 32, v[0]);      // every chunk is read, masked or not: TMEM bandwidth is plentiful, branches around an aligned load are not free
#pragma unroll
      for (int i = 0; i < SC_WCHUNKS; ++i) {
        const int c = c0 + i;
        const int off = off_of(i), dl = dl_c[i];
        const bool live = off < dl;                                             // warp-uniform: the chunk holds unmasked tokens
        tmem_ld_wait();
        if (i + 1 < SC_WCHUNKS) tmem_ld_32x32(taddr + (c + 1) * 32, v[(i + 1) & 1]);
        if (live) {
          const uint32_t (&vv)[32] = v[i & 1];
          if (!HAS_DL || off + 32 <= dl) {
            m = scan_max32(vv, m);
          } else {
            m = scan_max32_masked(vv, m, dl - off);
          }
        }
        const bool doc_ends = off + 32 == LD;                                     // the document's last columns are in this chunk
        const bool mine_ends = i == SC_WCHUNKS - 1 && !doc_ends && SPAN > 1;       // my columns end inside a longer document
        if (mine_ends) {
          // One partial slot per warp (shared memory is full: 227 KB of tiles).  The previous tile's partial is read by its
          // finisher before that warp hands the previous accumulator back, so "every epilogue warp has drained the previous
          // tile" (the accumulator's empty barrier) is what makes the slot free.
          if (it > 0) mbar_wait(&tempty_bar[(it - 1) & 1], ((it - 1) >> 1) & 1);
          part[(quad * (SC_EPQ - 1) + slot) * 32 + lane] = m;                           // handed to the warp that finishes the document
        } else if (doc_ends) {
          if (SPAN > 1) {
            asm volatile("bar.sync %0, %1;" ::"r"(quad_bar), "r"(32 * SC_EPQ) : "memory");   // the partners' partial maxima are in smem
            for (int s2 = slot - SPAN + 1; s2 < slot; ++s2) m = fmaxf(m, part[(quad * (SC_EPQ - 1) + s2) * 32 + lane]);
          }
          float sum = (lane < p.Lq) ? m : 0.f;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          const int64_t doc = doc0 + din_of(i);
          if (lane == 0 && q < p.nq && doc < p.Nd) out_row[doc] = sum;
          m = SC_PAD_FILL;
        }
      }
      // a warp that only handed a partial over still meets the quadrant's barrier (once per tile, every warp)
      if (SPAN > 1 && (slot % SPAN) != SPAN - 1) asm volatile("bar.sync %0, %1;" ::"r"(quad_bar), "r"(32 * SC_EPQ) : "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }
}

// Work split: the doc tiles are cut into chunks; unit u = (chunk u / QB, query block u % QB) goes to CTA u mod grid.
// A unit's query block sits in shared memory (the A operand of every tile of the unit), only doc tiles stream.
// The QB units of a chunk are consecutive, so they run on neighbouring CTAs at the same time: a doc tile comes from
// HBM once and from L2 for the other query blocks.  Units are much smaller than a CTA's share (about 64 per CTA), so
// every query-batch size keeps all SMs busy to within a couple of percent.
__global__ void __launch_bounds__(SC_THREADS, 1)
maxsim_scan_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d, const ScanParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + SC_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + SC_STAGES * SC_B_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES] doc tile landed
  uint64_t* empty_bar = bars + SC_STAGES;          // [STAGES] doc tile consumed by the MMAs
  uint64_t* tfull_bar = bars + 2 * SC_STAGES;      // [2] accumulator ready
  uint64_t* tempty_bar = bars + 2 * SC_STAGES + 2; // [2] accumulator drained
  uint64_t* afull_bar = bars + 2 * SC_STAGES + 4;  // query block landed
  uint64_t* aempty_bar = bars + 2 * SC_STAGES + 5; // last MMA of the pass retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SC_STAGES + 6);
  float* part = reinterpret_cast<float*>(smem_b + SC_STAGES * SC_B_BYTES + 256);   // [4 quads][SC_EPQ - 1 warps][32 lanes]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_d);
    for (int s = 0; s < SC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], SC_EPI_WARPS); }
    mbar_init(afull_bar, 1); mbar_init(aempty_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int cur_qb = -1; uint32_t a_uses = 0;
      for (int64_t u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int qb = int(u % p.QB);
        const int64_t dt0 = (u / p.QB) * p.chunk_tiles;
        const int64_t dt1 = min(p.DT, dt0 + p.chunk_tiles);
        if (qb != cur_qb) {
          // the MMA thread commits aempty when it reaches this unit: every MMA that read the old block has retired
          mbar_wait(aempty_bar, a_uses & 1);
          mbar_arrive_expect_tx(afull_bar, SC_A_BYTES);
          for (int kb = 0; kb < SC_KB; ++kb)
            tma_load_2d(smem_a + kb * (SC_BM * SC_BK * 2), &tmap_q, afull_bar, kb * SC_BK, qb * SC_BM);
          cur_qb = qb; ++a_uses;
        }
        for (int64_t dt = dt0; dt < dt1; ++dt) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sb = smem_b + stage * SC_B_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], SC_B_BYTES);
          for (int kb = 0; kb < SC_KB; ++kb)
            tma_load_2d(sb + kb * (SC_BN * SC_BK * 2), &tmap_d, &full_bar[stage], kb * SC_BK, int32_t(dt * SC_BN));
          if (++stage == SC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(SC_BM, SC_BN);
      int stage = 0; uint32_t phase = 0;
      uint32_t it = 0;
      int cur_qb = -1; uint32_t a_uses = 0;
      const uint32_t a_addr = smem_u32(smem_a);
      for (int64_t u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int qb = int(u % p.QB);
        const int64_t dt0 = (u / p.QB) * p.chunk_tiles;
        const int64_t dt1 = min(p.DT, dt0 + p.chunk_tiles);
        if (qb != cur_qb) {
          umma_commit(aempty_bar);                   // fires once every MMA issued so far has retired
          mbar_wait(afull_bar, a_uses & 1);
          tc_fence_after();
          cur_qb = qb; ++a_uses;
        }
        for (int64_t dt = dt0; dt < dt1; ++dt, ++it) {
          const uint32_t as = it & 1, aphase = (it >> 1) & 1;
          mbar_wait(&tempty_bar[as], aphase ^ 1);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + as * SC_BN;
          const uint32_t b_addr = smem_u32(smem_b + stage * SC_B_BYTES);
#pragma unroll
          for (int kb = 0; kb < SC_KB; ++kb) {
            const uint64_t da = umma_desc_k_sw128(a_addr + kb * (SC_BM * SC_BK * 2));
            const uint64_t db = umma_desc_k_sw128(b_addr + kb * (SC_BN * SC_BK * 2));
#pragma unroll
            for (int kk = 0; kk < SC_BK / 16; ++kk)
              umma_bf16_ss(tmem_d, da + uint64_t(kk * 2), db + uint64_t(kk * 2), idesc, (kb | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          umma_commit(&tfull_bar[as]);
          if (++stage == SC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue: max over a document's tokens, sum over the query's tokens =====
    // SC_EPQ warps per TMEM lane quadrant (= query inside the block); warp `slot` of a quadrant takes columns
    // [slot * SC_WCOLS, (slot + 1) * SC_WCOLS) of the tile.  A document longer than that spans several warps: their partial
    // maxima meet in shared memory and the warp that holds the document's last columns finishes it.
    // tools/micro/tmem_ld.cu: tcgen05.ld delivers 730-940 B/clk/SM to 8-16 reader warps, 6-7x what this epilogue needs; its
    // cost is the dependent max chains, which is why it is spread over 16 warps (4 per scheduler).
    const int quad = warp & 3;
    const int slot = (warp - 2) >> 2;
    // one specialised loop per (document length, masked or not): the tile's column -> (document, token) map, the chunks that end
    // a document and the hand-over between warps are compile-time facts of each instance, not per-tile branches
    if (p.doclen) {
      switch (p.Ld) {
        case 32: scan_epilogue<32, true>(p, tmem_base, tfull_bar, tempty_bar, part, quad, slot, lane); break;
        case 64: scan_epilogue<64, true>(p, tmem_base, tfull_bar, tempty_bar, part, quad, slot, lane); break;
        case 128: scan_epilogue<128, true>(p, tmem_base, tfull_bar, tempty_bar, part, quad, slot, lane); break;
        default: scan_epilogue<256, true>(p, tmem_base, tfull_bar, tempty_bar, part, quad, slot, lane); break;
      }
    } else {
      switch (p.Ld) {
        case 32: scan_epilogue<32, false>(p, tmem_base, tfull_bar, tempty_bar, part, quad, slot, lane); break;
        case 64: scan_epilogue<64, false>(p, tmem_base, tfull_bar, tempty_bar, part, quad, slot, lane); break;
        case 128: scan_epilogue<128, false>(p, tmem_base, tfull_bar, tempty_bar, part, quad, slot, lane); break;
        default: scan_epilogue<256, false>(p, tmem_base, tfull_bar, tempty_bar, part, quad, slot, lane); break;
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

static size_t scan_qpad_bytes(int nq) { return align_up(size_t((nq + 3) / 4) * SC_BM * SC_DIM * 2, 256); }

static int scan_launch(const void* D, const int32_t* doclen, int64_t Nd, int Ld, int dim, const void* Q, int nq, int Lq,
                       float* out, int64_t ld_out, void* qpad, cudaStream_t stream) {
  LRAG_REQUIRE(initialised(), "lrag_init has not been called");
  LRAG_REQUIRE(dim == SC_DIM, "maxsim_scan: dim=%d, only 128-d token vectors are supported", dim);
  LRAG_REQUIRE(Ld >= 32 && Ld <= 256 && SC_BN % Ld == 0, "maxsim_scan: Ld=%d must be 32, 64, 128 or 256 (documents tile the 256-row block)", Ld);
  LRAG_REQUIRE(Lq >= 1 && Lq <= 32, "maxsim_scan: Lq=%d must be in [1, 32]", Lq);
  LRAG_REQUIRE(nq > 0 && Nd > 0 && ld_out >= Nd, "maxsim_scan: empty problem or short output rows (nq=%d Nd=%lld ld=%lld)", nq, (long long)Nd, (long long)ld_out);
  LRAG_REQUIRE(Nd * int64_t(Ld) < (int64_t(1) << 31), "maxsim_scan: token store of %lld x %d rows exceeds one shard", (long long)Nd, Ld);
  LRAG_REQUIRE(D && Q && out && qpad, "maxsim_scan: null pointer");
  LRAG_REQUIRE((reinterpret_cast<uintptr_t>(D) & 15) == 0 && (reinterpret_cast<uintptr_t>(Q) & 15) == 0,
               "maxsim_scan: D and Q must be 16-byte aligned");
  ScanParams p;
  p.doclen = doclen; p.out = out; p.Nd = Nd; p.ld_out = ld_out; p.Ld = Ld; p.Lq = Lq; p.nq = nq;
  p.QB = (nq + 3) / 4;
  { const char* dbg = getenv("LRAG_SCAN_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
  p.DT = (Nd * Ld + SC_BN - 1) / SC_BN;
  // queries as 32-row blocks: [nq, Lq, 128] -> zero-padded [QB * 4, 32, 128] (in-bounds TMA boxes, zero rows add 0)
  LRAG_CHECK_CUDA(cudaMemsetAsync(qpad, 0, scan_qpad_bytes(nq), stream));
  LRAG_CHECK_CUDA(cudaMemcpy2DAsync(qpad, 32 * SC_DIM * 2, Q, size_t(Lq) * SC_DIM * 2, size_t(Lq) * SC_DIM * 2, nq,
                                    cudaMemcpyDeviceToDevice, stream));
  CUtensorMap tq, td;
  int rc = make_tmap_bf16_2d(&tq, qpad, uint64_t(p.QB) * SC_BM, SC_DIM, SC_DIM, SC_BM, SC_BK);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&td, D, uint64_t(Nd) * Ld, SC_DIM, SC_DIM, SC_BN, SC_BK);
  if (rc) return rc;
  static bool attr_set[LRAG_MAX_DEVICES] = {};      // function attributes are per device
  const int dev = device_slot();
  if (!attr_set[dev]) {
    LRAG_CHECK_CUDA(cudaFuncSetAttribute(maxsim_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM));
    attr_set[dev] = true;
  }
  const int sms = sm_count();
  int64_t nchunks = (int64_t(sms) * 64 + p.QB - 1) / p.QB;             // about 64 units per CTA ...
  if (nchunks > p.DT) nchunks = p.DT;
  p.chunk_tiles = (p.DT + nchunks - 1) / nchunks;
  // ... but chunks small enough that the ones in flight stay in L2.  The machine works on sms / QB (+ 1) chunks at a time, and
  // when sms is not a multiple of QB a chunk's query blocks straddle two rounds of the grid: the late ones find the chunk's
  // tiles in L2 only if a whole round's footprint fits (55 MB chunks re-read the store 2.2 x from DRAM, ncu round 2).
  {
    const int64_t in_flight = (sms + p.QB - 1) / p.QB + 1;
    const char* e = getenv("LRAG_SCAN_L2_MB");
    const int64_t budget_tiles = (int64_t(e ? atoi(e) : 8) << 20) / SC_B_BYTES;   // 8 MB: measured best of 4 ... 96 MB at 64, 256 and 1024 queries
    const int64_t lim = std::max<int64_t>(8, budget_tiles / in_flight);
    if (p.chunk_tiles > lim) p.chunk_tiles = lim;
  }
  nchunks = (p.DT + p.chunk_tiles - 1) / p.chunk_tiles;
  p.units = nchunks * p.QB;
  const int grid = int(p.units < sms ? p.units : sms);
  prof_begin(stream, PROF_MAXSIM_SCAN);
  maxsim_scan_kernel<<<grid, SC_THREADS, SC_SMEM, stream>>>(tq, td, p);
  prof_end(stream);
  LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
  return LRAG_OK;
}

}  // namespace lrag

using namespace lrag;

extern "C" size_t lrag_maxsim_scan_workspace_bytes(int64_t Nd, int nq, int k) {
  if (Nd <= 0 || nq <= 0) return 0;
  // padded query block (+ the [nq, Nd] score matrix when the top-k entry point is used: k > 0)
  return scan_qpad_bytes(nq) + (k > 0 ? align_up(size_t(nq) * size_t(Nd) * 4, 256) + topk_select_ws_bytes(nq, Nd, k) : 0);
}

extern "C" int lrag_maxsim_scan_scores_bf16(const void* D, const int32_t* doclen, int64_t Nd, int Ld, int dim, const void* Q,
                                            int nq, int Lq, float* out_score, int64_t ld_out, void* ws, size_t ws_bytes,
                                            lrag_stream_t stream) {
  const size_t need = lrag_maxsim_scan_workspace_bytes(Nd, nq, 0);
  if (ws_bytes < need || !ws) { set_error("maxsim_scan_scores: workspace %zu < required %zu", ws_bytes, need); return LRAG_ENOSPC; }
  return scan_launch(D, doclen, Nd, Ld, dim, Q, nq, Lq, out_score, ld_out, ws, static_cast<cudaStream_t>(stream));
}

extern "C" int lrag_maxsim_scan_topk_bf16(const void* D, const int32_t* doclen, int64_t Nd, int Ld, int dim, const void* Q,
                                          int nq, int Lq, int k, int64_t id_base, float* out_score, int64_t* out_id,
                                          void* ws, size_t ws_bytes, lrag_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LRAG_REQUIRE(k > 0 && k <= LRAG_MAX_K, "maxsim_scan_topk: need 1 <= k <= %d (k=%d)", LRAG_MAX_K, k);
  LRAG_REQUIRE(out_score && out_id, "maxsim_scan_topk: null output");
  const size_t need = lrag_maxsim_scan_workspace_bytes(Nd, nq, k);
  if (ws_bytes < need || !ws) { set_error("maxsim_scan_topk: workspace %zu < required %zu", ws_bytes, need); return LRAG_ENOSPC; }
  float* S = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + scan_qpad_bytes(nq));
  int rc = scan_launch(D, doclen, Nd, Ld, dim, Q, nq, Lq, S, Nd, ws, stream);
  if (rc) return rc;
  uint8_t* sel_ws = reinterpret_cast<uint8_t*>(S) + align_up(size_t(nq) * size_t(Nd) * 4, 256);
  return launch_topk_select(S, Nd, nq, Nd, k, id_base, nullptr, out_score, out_id, stream, sel_ws, topk_select_ws_bytes(nq, Nd, k));
}
