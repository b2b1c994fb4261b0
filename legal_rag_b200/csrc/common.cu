// liblrag runtime glue: device binding, error reporting, TMA descriptor encoding.
#include "common.cuh"

#include <atomic>
#include <cstdarg>
#include <cstring>

namespace lrag {

static thread_local char g_err[512] = "";
static int g_sm_count = 0;
static int g_device = -1;
static encode_tiled_fn g_encode = nullptr;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int sm_count() { return g_sm_count > 0 ? g_sm_count : 148; }
bool initialised() { return g_device >= 0; }
int device_slot() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0) d = 0;
  return d % LRAG_MAX_DEVICES;
}
encode_tiled_fn tensor_map_encoder() { return g_encode; }

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_cols) {
  if (!g_encode) { set_error("tensor-map encoder unavailable (lrag_init not called?)"); return LRAG_ECUDA; }
  if (box_cols * 2 != 128) { set_error("tensor map: inner box must be 64 bf16 (one 128 B swizzle row)"); return LRAG_EINVAL; }
  if (box_rows == 0 || box_rows > 256) { set_error("tensor map: box rows %u out of range", box_rows); return LRAG_EINVAL; }
  cuuint64_t gdim[2] = {cols, rows};                       // innermost first
  cuuint64_t gstride[1] = {row_stride_elems * 2};          // bytes, dim 1
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box,
                        estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rows=%llu cols=%llu stride=%llu box=%ux%u base=%p", int(r),
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)row_stride_elems, box_rows,
              box_cols, base);
    return LRAG_ECUDA;
  }
  return LRAG_OK;
}

// ---- launch profiler: CUDA event pairs recorded on the launching stream around tagged kernels ----
static cudaEvent_t* g_prof_ev = nullptr;   // [2 * cap]
static int* g_prof_tag = nullptr;
static int g_prof_cap = 0, g_prof_n = 0;
static bool g_prof_open = false;
static std::atomic<long long> g_launches{0};     // entry points may be called from several host threads
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void prof_begin(cudaStream_t stream, int tag) {
  if (!g_prof_cap || g_prof_n >= g_prof_cap) return;
  g_prof_tag[g_prof_n] = tag;
  cudaEventRecord(g_prof_ev[2 * g_prof_n], stream);
  g_prof_open = true;
}
void prof_end(cudaStream_t stream) {
  if (!g_prof_open) return;
  cudaEventRecord(g_prof_ev[2 * g_prof_n + 1], stream);
  ++g_prof_n;
  g_prof_open = false;
}

}  // namespace lrag

using namespace lrag;

extern "C" int lrag_prof_enable(int capacity) {
  for (int i = 0; i < 2 * g_prof_cap; ++i) cudaEventDestroy(g_prof_ev[i]);
  delete[] g_prof_ev; delete[] g_prof_tag;
  g_prof_ev = nullptr; g_prof_tag = nullptr; g_prof_cap = 0; g_prof_n = 0; g_prof_open = false;
  if (capacity <= 0) return LRAG_OK;
  g_prof_ev = new cudaEvent_t[2 * capacity];
  g_prof_tag = new int[capacity];
  for (int i = 0; i < 2 * capacity; ++i) LRAG_CHECK_CUDA(cudaEventCreate(&g_prof_ev[i]));
  g_prof_cap = capacity;
  return LRAG_OK;
}

extern "C" int lrag_prof_collect(float* ms, int* tag, int max_n) {
  int n = g_prof_n < max_n ? g_prof_n : max_n;
  for (int i = 0; i < n; ++i) {
    LRAG_CHECK_CUDA(cudaEventSynchronize(g_prof_ev[2 * i + 1]));
    LRAG_CHECK_CUDA(cudaEventElapsedTime(&ms[i], g_prof_ev[2 * i], g_prof_ev[2 * i + 1]));
    tag[i] = g_prof_tag[i];
  }
  g_prof_n = 0;
  return n;
}

extern "C" long long lrag_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" int lrag_version(void) { return LRAG_VERSION; }
extern "C" const char* lrag_last_error(void) { return g_err; }
extern "C" int lrag_sm_count(void) { return g_sm_count; }

extern "C" int lrag_init(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("lrag_init: no CUDA device visible (%s); this engine has no CPU fallback", cudaGetErrorString(e));
    return LRAG_ECUDA;
  }
  if (device < 0 || device >= count) { set_error("lrag_init: device %d out of range (0..%d)", device, count - 1); return LRAG_EINVAL; }
  cudaDeviceProp prop;
  LRAG_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("lrag_init: device %d is sm_%d%d; liblrag is built for sm_100a only and has no fallback path", device,
              prop.major, prop.minor);
    return LRAG_EARCH;
  }
  LRAG_CHECK_CUDA(cudaSetDevice(device));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  LRAG_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || !fn) { set_error("lrag_init: cuTensorMapEncodeTiled not found in the driver"); return LRAG_ECUDA; }
  g_encode = reinterpret_cast<encode_tiled_fn>(fn);
  g_sm_count = prop.multiProcessorCount;
  g_device = device;
  g_err[0] = 0;
  return LRAG_OK;
}
