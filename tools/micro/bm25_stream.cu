// Micro-benchmark behind the round-2 BM25 consumer loop: warps stream posting segments straight from
// global memory / L2 into registers and add them into a shared-memory slab with integer atomics.
// Question it answers: how many postings per second does that loop sustain, per load flavour and depth,
// against the 0.44-0.71 T postings/s of the round-1 kernel (bulk-copy ring + 16 warps per stage)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bm25_stream bm25_stream.cu && ./bm25_stream
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int SLAB = 24576;          // 96 KB of int32 accumulators
constexpr int WARPS = 16;

__device__ __forceinline__ int ldg_nc_s32(const int* p) {
  int v; asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}
__device__ __forceinline__ float ldg_nc_f32(const float* p) {
  float v; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
}
__device__ __forceinline__ int4 ldg_nc_v4(const void* p) {
  int4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v;
}
__device__ __forceinline__ int atoms_add(uint32_t a, int v) {
  int o; asm volatile("atom.shared.add.s32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory"); return o;
}
__device__ __forceinline__ uint32_t perm(uint32_t i) {   // 4 consecutive docs -> one bank, lanes -> distinct banks
  return (i & ~127u) | ((i >> 2) & 31u) | ((i & 3u) << 5);
}

// MODE 0: LDG.32, SEG postings loaded then added (single buffer)
// MODE 1: LDG.32, double buffered (next segment's loads in flight while this one is added)
// MODE 2: LDG.128 + bank permutation, single buffer
// MODE 3: LDG.128 + bank permutation, double buffered
template <int MODE, int SEG>
__global__ void __launch_bounds__(WARPS * 32, 2)
k(const int* __restrict__ ids, const float* __restrict__ imp, long long nnz, int passes, float ms, int* sink) {
  extern __shared__ int acc[];
  __shared__ int head;
  for (int i = threadIdx.x; i < SLAB; i += blockDim.x) acc[i] = 0;
  if (threadIdx.x == 0) head = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t accb = (uint32_t)__cvta_generic_to_shared(acc);
  const long long nseg = nnz / SEG;
  const long long total = nseg * passes;
  const long long stagger = (long long)blockIdx.x * 977 % nseg;
  int mx = 0;
  constexpr bool V4 = MODE >= 2;
  constexpr int PER = V4 ? SEG / 128 : SEG / 32;     // load instructions per array per segment
  auto claim = [&]() -> long long {
    int s = 0;
    if (lane == 0) s = atomicAdd(&head, 1);
    s = __shfl_sync(0xffffffffu, s, 0);
    return s;
  };
  auto base_of = [&](long long s) -> long long { return ((s + stagger) % nseg) * SEG; };
  if constexpr (!V4) {
    int d[2][PER]; float v[2][PER];
    auto load = [&](int b, long long s) {
      const long long o = base_of(s) + lane;
#pragma unroll
      for (int u = 0; u < PER; ++u) { d[b][u] = ldg_nc_s32(ids + o + u * 32); v[b][u] = ldg_nc_f32(imp + o + u * 32); }
    };
    auto add = [&](int b) {
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int vi = __float2int_rn(v[b][u] * ms);
        mx = max(mx, atoms_add(accb + 4u * (uint32_t)d[b][u], vi) + vi);
      }
    };
    if constexpr (MODE == 0) {
      for (long long s = claim(); s < total; s = claim()) { load(0, s); add(0); }
    } else {
      long long s = claim();
      if (s < total) load(0, s);
      while (s < total) {
        const long long s2 = claim();
        if (s2 < total) load(1, s2);
        add(0);
        s = s2;
        if (s >= total) break;
        const long long s3 = claim();
        if (s3 < total) load(0, s3);
        add(1);
        s = s3;
      }
    }
  } else {
    int4 d[2][PER]; int4 v[2][PER];
    auto load = [&](int b, long long s) {
      const long long o = base_of(s) + lane * 4;
#pragma unroll
      for (int u = 0; u < PER; ++u) { d[b][u] = ldg_nc_v4(ids + o + u * 128); v[b][u] = ldg_nc_v4(imp + o + u * 128); }
    };
    auto add1 = [&](int dd, int vv) {
      const int vi = __float2int_rn(__int_as_float(vv) * ms);
      mx = max(mx, atoms_add(accb + 4u * perm((uint32_t)dd), vi) + vi);
    };
    auto add = [&](int b) {
#pragma unroll
      for (int u = 0; u < PER; ++u) { add1(d[b][u].x, v[b][u].x); add1(d[b][u].y, v[b][u].y); add1(d[b][u].z, v[b][u].z); add1(d[b][u].w, v[b][u].w); }
    };
    if constexpr (MODE == 2) {
      for (long long s = claim(); s < total; s = claim()) { load(0, s); add(0); }
    } else {
      long long s = claim();
      if (s < total) load(0, s);
      while (s < total) {
        const long long s2 = claim();
        if (s2 < total) load(1, s2);
        add(0);
        s = s2;
        if (s >= total) break;
        const long long s3 = claim();
        if (s3 < total) load(0, s3);
        add(1);
        s = s3;
      }
    }
  }
  if (mx == 123456789) sink[0] = mx;
  __syncthreads();
  if (threadIdx.x == 0) sink[1 + (blockIdx.x & 7)] = acc[5];
}

template <int MODE, int SEG>
void run(const char* name, const int* ids, const float* imp, long long nnz, int passes, int* sink, int pattern) {
  cudaFuncSetAttribute(k<MODE, SEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SLAB * 4);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<MODE, SEG>, WARPS * 32, SLAB * 4);
  const int blocks = 148 * occ;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, SEG><<<blocks, WARPS * 32, SLAB * 4>>>(ids, imp, nnz, 1, 3.5f, sink);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE, SEG><<<blocks, WARPS * 32, SLAB * 4>>>(ids, imp, nnz, passes, 3.5f, sink);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double postings = double(nnz / SEG * SEG) * passes * blocks;
  printf("pattern %d %-34s occ %d: %8.3f ms  %.3f T postings/s  %.2f TB/s of (id,impact)  [%s]\n", pattern, name, occ, ms,
         postings / ms / 1e9, postings * 8 / ms / 1e9 / 1e3 * 1e0, cudaGetErrorString(err));
}

int main(int argc, char** argv) {
  const long long nnz = argc > 1 ? atoll(argv[1]) : (8ll << 20);      // 8M postings = 64 MB: L2 resident, like the shared head terms
  const int passes = argc > 2 ? atoi(argv[2]) : 2;
  int* ids; float* imp; int* sink;
  cudaMalloc(&ids, nnz * 4); cudaMalloc(&imp, nnz * 4); cudaMalloc(&sink, 64);
  int* h = (int*)malloc(nnz * 4); float* hf = (float*)malloc(nnz * 4);
  for (int pattern = 0; pattern < 3; ++pattern) {
    uint32_t r = 12345; long long cur = 0;
    for (long long i = 0; i < nnz; ++i) {
      r = r * 1664525u + 1013904223u;
      if (pattern == 0) cur += 1; else if (pattern == 1) cur += 1 + (r >> 16) % 19; else cur += 1 + (r >> 12) % 1999;
      h[i] = int(cur % SLAB); hf[i] = 0.25f + float(r >> 20) * 1e-4f;
    }
    cudaMemcpy(ids, h, nnz * 4, cudaMemcpyHostToDevice); cudaMemcpy(imp, hf, nnz * 4, cudaMemcpyHostToDevice);
    run<0, 128>("LDG.32 seg128 single", ids, imp, nnz, passes, sink, pattern);
    run<0, 256>("LDG.32 seg256 single", ids, imp, nnz, passes, sink, pattern);
    run<1, 128>("LDG.32 seg128 double", ids, imp, nnz, passes, sink, pattern);
    run<1, 256>("LDG.32 seg256 double", ids, imp, nnz, passes, sink, pattern);
    run<2, 256>("LDG.128+perm seg256 single", ids, imp, nnz, passes, sink, pattern);
    run<2, 512>("LDG.128+perm seg512 single", ids, imp, nnz, passes, sink, pattern);
    run<3, 128>("LDG.128+perm seg128 double", ids, imp, nnz, passes, sink, pattern);
    run<3, 256>("LDG.128+perm seg256 double", ids, imp, nnz, passes, sink, pattern);
  }
  return 0;
}
