// Micro-benchmark: throughput of scattered shared-memory read-modify-write, three ways.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_rmw smem_rmw.cu && ./smem_rmw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int SLAB = 16384, ITERS = 2048;
template <int MODE>
__global__ void __launch_bounds__(512, 2) k(const int* __restrict__ idx, int stride_mode, unsigned long long* out_cycles, int* sink) {
  extern __shared__ int acc[];
  for (int i = threadIdx.x; i < SLAB; i += blockDim.x) acc[i] = 0;
  __syncthreads();
  // each thread's doc sequence: pre-generated indices (sorted within a warp-chunk like a posting list)
  int my[8];
  for (int u = 0; u < 8; ++u) my[u] = idx[(blockIdx.x * 8 + u) * 512 + threadIdx.x] & (SLAB - 1);
  int mx = 0;
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int d = (my[u] + it * 33) & (SLAB - 1);
      if (MODE == 0) { float a = __int_as_float(acc[d]); a += 1.5f; acc[d] = __float_as_int(a); mx = max(mx, __float_as_int(a)); }
      if (MODE == 1) { int o = atomicAdd(&acc[d], 3); mx = max(mx, o + 3); }
      if (MODE == 2) { atomicAdd(&acc[d], 3); }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
  if (mx == 123456789) sink[0] = mx;
  __syncthreads();
  if (MODE == 2 && threadIdx.x == 0) sink[1] = acc[5];
}
int main() {
  const int blocks = 296;
  int* idx; unsigned long long* cyc; int* sink;
  cudaMalloc(&idx, blocks * 8 * 512 * 4); cudaMalloc(&cyc, blocks * 8); cudaMalloc(&sink, 16);
  int* h = new int[blocks * 8 * 512];
  for (int pattern = 0; pattern < 3; ++pattern) {
    // 0: dense consecutive docs, 1: every ~10th doc (sorted, jittered), 2: random
    uint32_t r = 12345;
    for (int b = 0; b < blocks * 8; ++b) {
      int base = (b * 977) & (SLAB - 1), cur = base;
      for (int t = 0; t < 512; ++t) {
        r = r * 1664525u + 1013904223u;
        if (pattern == 0) cur = base + t; else if (pattern == 1) cur += 1 + (r >> 16) % 19; else cur = r >> 8;
        h[b * 512 + t] = cur & (SLAB - 1);
      }
    }
    cudaMemcpy(idx, h, blocks * 8 * 512 * 4, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 3; ++mode) {
      auto launch = [&]() {
        if (mode == 0) { cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SLAB * 4); k<0><<<blocks, 512, SLAB * 4>>>(idx, 0, cyc, sink); }
        if (mode == 1) { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SLAB * 4); k<1><<<blocks, 512, SLAB * 4>>>(idx, 0, cyc, sink); }
        if (mode == 2) { cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SLAB * 4); k<2><<<blocks, 512, SLAB * 4>>>(idx, 0, cyc, sink); }
      };
      launch(); cudaDeviceSynchronize(); launch();
      cudaError_t e = cudaDeviceSynchronize();
      unsigned long long hc[296]; cudaMemcpy(hc, cyc, blocks * 8, cudaMemcpyDeviceToHost);
      double avg = 0; for (int b = 0; b < blocks; ++b) avg += hc[b]; avg /= blocks;
      // per SM: 2 CTAs x 16 warps x ITERS x 8 warp-ops in `avg` cycles
      printf("pattern %d mode %d (%s): %.0f cycles, %.2f cycles per warp-RMW per SM (%s)\n", pattern, mode,
             mode == 0 ? "LDS+FADD+STS" : mode == 1 ? "ATOMS.ADD ret" : "ATOMS/RED noret", avg, avg / (2.0 * 16 * ITERS * 8), cudaGetErrorString(e));
    }
  }
  return 0;
}
