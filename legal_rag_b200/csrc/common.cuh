// Shared device/host helpers for liblrag (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>

#include "../../include/lrag.h"

namespace lrag {

// ----------------------------------------------------------------------------------
// host-side error plumbing
// ----------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int sm_count();                 // cached by lrag_init (148 on B200)
bool initialised();
// Index of the calling thread's current device into per-device tables (cudaFuncSetAttribute is per device, so
// "already set" flags must be too).
constexpr int LRAG_MAX_DEVICES = 64;
int device_slot();
// driver entry point for cuTensorMapEncodeTiled, fetched through the runtime
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_fn tensor_map_encoder();
// 2-D bf16 row-major [rows, cols] tensor map, box [box_rows, 64 cols], 128B swizzle
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_cols);

// optional per-launch timing of the dominant kernels (lrag_prof_enable / lrag_prof_collect)
enum ProfTag { PROF_DENSE_SCAN = 0, PROF_BM25_SCAN = 1, PROF_MAXSIM = 2, PROF_FUSE = 3, PROF_SELECT = 4, PROF_MAXSIM_SCAN = 5 };
void prof_begin(cudaStream_t stream, int tag);
void prof_end(cudaStream_t stream);
void note_launch(int n = 1);   // every kernel the library launches is counted (lrag_launch_count)

#define LRAG_CHECK_CUDA(expr)                                                          \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      lrag::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                      __LINE__);                                                       \
      return LRAG_ECUDA;                                                               \
    }                                                                                  \
  } while (0)

#define LRAG_REQUIRE(cond, ...)   \
  do {                            \
    if (!(cond)) {                \
      lrag::set_error(__VA_ARGS__); \
      return LRAG_EINVAL;         \
    }                             \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ----------------------------------------------------------------------------------
// 64-bit selection keys: (orderable score) << 32 | (0xFFFFFFFF - local id)
// larger key == better hit under (score desc, id asc).  Keys are unique per row.
// ----------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t ord32(float f) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t b = c.u;
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float unord32(uint32_t o) {
  uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}
#define LRAG_ORD_NEG_INF 0x007fffffu  // ord32(-inf)

__device__ __forceinline__ uint64_t make_key(float s, uint32_t local_id) {
  return (uint64_t(ord32(s)) << 32) | uint64_t(0xffffffffu - local_id);
}
__device__ __forceinline__ float key_score(uint64_t k) { return unord32(uint32_t(k >> 32)); }
__device__ __forceinline__ uint32_t key_id(uint64_t k) { return 0xffffffffu - uint32_t(k); }

#ifdef __CUDACC__
// ----------------------------------------------------------------------------------
// PTX wrappers (mbarrier, TMA, tcgen05).  Hand-written; no CUTLASS.
// ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// The suspend-time hint lets the hardware park the thread until the phase completes (or the hint
// expires) instead of returning at once: a waiting warp then issues almost nothing and leaves the
// issue slots to the warps that produce what it is waiting for.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (CUDA error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 20000000000LL) {  // ~10 s at 2 GHz
      printf("lrag: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
// Same, polling without the suspend hint.  Which flavour wins is measured per kernel
// (tools/gpu_wait_ab.sh): the dense scan gains 7 % from parked waits (its epilogue warps wait long and
// would otherwise take issue slots from the TMA / MMA warps), the MaxSim gather loses 13 % to the wake-up
// latency (its waits are short and frequent), so maxsim.cu polls.
__device__ __forceinline__ bool mbar_try_wait_now(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_now(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait_now(bar, parity)) {
    if (clock64() - t0 > 20000000000LL) {
      printf("lrag: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0),
      "r"(c1)
      : "memory");
}
// The same with an L2 eviction-priority hint (createpolicy): a stream that is read once (or by a few CTAs within microseconds)
// is marked evict-first so that it does not push longer-lived lines of a kernel running beside it out of L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar,
                                                 int32_t c0, int32_t c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0),
      "r"(c1), "l"(policy)
      : "memory");
}
// 1-D bulk asynchronous copy global -> shared (16-byte aligned source, destination and size); its bytes complete on `bar`
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; bf16 inputs, fp32 accumulate; one thread issues.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets row (lane base + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor for a K-major bf16 tile stored as [rows][64] with the
// 128-byte swizzle TMA writes (8-row x 128 B atoms, 1024 B apart).  Field layout checked
// against cute/arch/mma_sm100_desc.hpp (SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) with SWIZZLE_128B = 2.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3ffff) >> 4);
  d |= uint64_t(1) << 16;                 // LBO (ignored for swizzled K-major)
  d |= uint64_t(1024 >> 4) << 32;         // SBO: next 8-row group
  d |= uint64_t(1) << 46;                 // descriptor version (Blackwell)
  d |= uint64_t(2) << 61;                 // SWIZZLE_128B
  return d;
}
// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32, A and B K-major.
// (InstrDescriptor in mma_sm100_desc.hpp: c_format [4,6), a_format [7,10), b_format [10,13),
//  a_major 15, b_major 16, n>>3 [17,23), m>>4 [24,29).)
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// ----------------------------------------------------------------------------------
// warp-cooperative in-place "keep the k largest keys" over buf[0..n) (n >= k).
// Returns the k-th largest key (the pivot).  All 32 lanes must call.
// Bitwise binary search over the 64-bit key space, skipping bits every key agrees on.
// ----------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t warp_keep_topk(uint64_t* buf, int n, int k, int lane) {
  __syncwarp();
  constexpr int KPL = 8;                       // register-cached keys per lane (n <= 256)
  const bool in_regs = n <= 32 * KPL;
  uint64_t r[KPL];
  uint64_t all_or = 0, all_and = ~uint64_t(0);
  if (in_regs) {
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      int i = j * 32 + lane;
      r[j] = (i < n) ? buf[i] : 0;
      if (i < n) { all_or |= r[j]; all_and &= r[j]; }
    }
  } else {
    for (int i = lane; i < n; i += 32) {
      uint64_t v = buf[i];
      all_or |= v;
      all_and &= v;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    all_or |= __shfl_xor_sync(0xffffffffu, all_or, o);
    all_and &= __shfl_xor_sync(0xffffffffu, all_and, o);
  }
  uint64_t varying = all_or & ~all_and;   // bits where keys differ
  uint64_t prefix = all_and;              // bits every key shares belong to the pivot too
  while (varying) {
    const int bit = 63 - __clzll((long long)varying);
    varying &= ~(uint64_t(1) << bit);
    const uint64_t himask = ~((uint64_t(1) << bit) - 1);            // bits >= bit
    const uint64_t cand = (prefix | (uint64_t(1) << bit)) & himask;  // decided bits + this bit
    int c = 0;
    if (in_regs) {
#pragma unroll
      for (int j = 0; j < KPL; ++j) c += ((j * 32 + lane < n) && ((r[j] & himask) >= cand)) ? 1 : 0;
    } else {
      for (int i = lane; i < n; i += 32) c += ((buf[i] & himask) >= cand) ? 1 : 0;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= k) prefix |= uint64_t(1) << bit;
  }
  // prefix is now exactly the k-th largest key; keep keys >= pivot (exactly k: keys are unique)
  const uint64_t pivot = prefix;
  int out = 0;
  if (in_regs) {
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      if (j * 32 >= n) break;
      const bool keep = (j * 32 + lane < n) && (r[j] >= pivot);
      const uint32_t m = __ballot_sync(0xffffffffu, keep);
      if (keep) buf[out + __popc(m & ((1u << lane) - 1))] = r[j];
      out += __popc(m);
    }
  } else {
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      const uint64_t v = (i < n) ? buf[i] : 0;
      const bool keep = (i < n) && (v >= pivot);
      const uint32_t m = __ballot_sync(0xffffffffu, keep);
      __syncwarp();   // every lane has read its element before any lane overwrites [out, out+32)
      if (keep) buf[out + __popc(m & ((1u << lane) - 1))] = v;
      out += __popc(m);
    }
  }
  __syncwarp();
  return pivot;
}
#endif  // __CUDACC__

}  // namespace lrag
