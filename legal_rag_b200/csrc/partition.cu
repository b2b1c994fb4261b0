// SM partitioning for stages that run side by side (the hybrid step's dense scan and BM25 scan).
//
// Both scans are persistent kernels that size their grids to the machine.  Run one after the other on a power-capped
// B200, the tensor-bound dense scan pulls the SM clock down to ~1.35 GHz at the cap, and the BM25 scan that follows --
// L2-latency bound, ~600 W at full clock -- inherits a clock the governor raises again only slowly: the step averages
// ~800 W of a ~1000 W budget.  Side by side on disjoint SM sets the board draws a steady load at one clock and the
// budget is used all the time.
//
// The two kernels cannot share an SM (192 KB + 2 x 113 KB of shared memory), so a partition is a matter of who gets
// which SMs, and the hardware places CTAs wherever they fit.  lrag_sm_reserve makes the placement deterministic: it
// parks one CTA holding 200 KB of shared memory on `ctas` SMs; the BM25 scan, launched next with a grid of
// 2 x (SMs - ctas) CTAs, can only land on the other SMs, two per SM, and counts its CTAs in as they start; when the
// count reaches the target the reservation exits, and the dense scan (whose stream waits for it) finds exactly those
// `ctas` SMs free -- every other SM is full -- for all of its epochs.
#include "common.cuh"

namespace lrag {

constexpr int RESERVE_SMEM = 200 * 1024;

__global__ void __launch_bounds__(32) sm_reserve_kernel(const unsigned long long* counter, unsigned long long target,
                                                        long long timeout_cycles) {
  extern __shared__ uint8_t hold[];
  if (threadIdx.x == 0) {
    hold[0] = 0;
    const long long t0 = clock64();
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(counter) : "memory");
      if (v >= target) break;
      if (clock64() - t0 > timeout_cycles) break;            // the other kernel never came: give the SMs back
      __nanosleep(500);
    }
  }
}

}  // namespace lrag

using namespace lrag;

extern "C" int lrag_sm_reserve(int ctas, const unsigned long long* counter, unsigned long long target, int timeout_ms,
                               lrag_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LRAG_REQUIRE(initialised(), "lrag_init has not been called");
  LRAG_REQUIRE(ctas >= 0 && ctas <= sm_count() && counter && timeout_ms >= 0 && timeout_ms <= 10000,
               "sm_reserve: need 0 <= ctas <= %d, a counter and a timeout of at most 10 s (ctas=%d timeout=%d ms)", sm_count(), ctas, timeout_ms);
  if (ctas == 0) return LRAG_OK;
  static bool attr_set[LRAG_MAX_DEVICES] = {};      // function attributes are per device
  const int dev = device_slot();
  if (!attr_set[dev]) {
    LRAG_CHECK_CUDA(cudaFuncSetAttribute(sm_reserve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RESERVE_SMEM));
    attr_set[dev] = true;
  }
  sm_reserve_kernel<<<ctas, 32, RESERVE_SMEM, stream>>>(counter, target, (long long)timeout_ms * 2000000LL);
  LRAG_CHECK_CUDA(cudaGetLastError()); note_launch();
  return LRAG_OK;
}
