// Host runtime of libmindrec_b200.so (not aot): device / pinned memory, copies, streams, events and CUDA-graph capture
// behind a plain C interface, so that the Python host code (mindrec_b200/runtime.py) can drive the aot kernels with
// nothing but ctypes — no PyTorch and no cuda-python in the process (BASELINE north_star: "Python host code calls a .so
// ... with no PyTorch").  Under MindSpore the framework owns these resources and only the aot symbols are used.
#include "common.cuh"

namespace mrec {
__global__ void fill32_kernel(uint32_t* __restrict__ p, uint32_t v, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
static int rt_fail(const char* what, cudaError_t e) {
  cudaGetLastError();
  return fail(ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
}  // namespace mrec
using namespace mrec;

#define RT_CHECK(call, what)                  \
  do {                                        \
    cudaError_t _e = (call);                  \
    if (_e != cudaSuccess) return rt_fail(what, _e); \
  } while (0)

MREC_API int mrec_rt_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
MREC_API int mrec_rt_set_device(int i) { RT_CHECK(cudaSetDevice(i), "mrec_rt_set_device"); return OK; }
MREC_API int mrec_rt_mem_info(size_t* free_b, size_t* total_b) { RT_CHECK(cudaMemGetInfo(free_b, total_b), "mrec_rt_mem_info"); return OK; }
MREC_API void* mrec_rt_malloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { rt_fail("mrec_rt_malloc", cudaGetLastError()); return nullptr; }
  return p;
}
MREC_API int mrec_rt_free(void* p) { RT_CHECK(cudaFree(p), "mrec_rt_free"); return OK; }
MREC_API void* mrec_rt_malloc_host(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { rt_fail("mrec_rt_malloc_host", cudaGetLastError()); return nullptr; }
  return p;
}
MREC_API int mrec_rt_free_host(void* p) { RT_CHECK(cudaFreeHost(p), "mrec_rt_free_host"); return OK; }
// kind: 1 host -> device, 2 device -> host, 3 device -> device; asynchronous on `stream` (pinned host memory for overlap)
MREC_API int mrec_rt_memcpy(void* dst, const void* src, size_t bytes, int kind, void* stream) {
  const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  RT_CHECK(cudaMemcpyAsync(dst, src, bytes, k, (cudaStream_t)stream), "mrec_rt_memcpy");
  return OK;
}
MREC_API int mrec_rt_memset(void* p, int byte, size_t bytes, void* stream) {
  RT_CHECK(cudaMemsetAsync(p, byte, bytes, (cudaStream_t)stream), "mrec_rt_memset");
  return OK;
}
MREC_API int mrec_rt_fill32(void* p, uint32_t pattern, int64_t n, void* stream) {
  if (n <= 0) return OK;
  MREC_LAUNCH(fill32_kernel, grid_for(cdiv(n, 256), 8), 256, 0, (cudaStream_t)stream, reinterpret_cast<uint32_t*>(p), pattern, n);
  return check_launch("mrec_rt_fill32");
}
MREC_API void* mrec_rt_stream_create(void) {
  cudaStream_t s = nullptr;
  if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { rt_fail("mrec_rt_stream_create", cudaGetLastError()); return nullptr; }
  return s;
}
MREC_API int mrec_rt_stream_destroy(void* s) { RT_CHECK(cudaStreamDestroy((cudaStream_t)s), "mrec_rt_stream_destroy"); return OK; }
MREC_API int mrec_rt_stream_sync(void* s) { RT_CHECK(cudaStreamSynchronize((cudaStream_t)s), "mrec_rt_stream_sync"); return OK; }
MREC_API int mrec_rt_device_sync(void) { RT_CHECK(cudaDeviceSynchronize(), "mrec_rt_device_sync"); return OK; }
MREC_API void* mrec_rt_event_create(int timing) {
  cudaEvent_t e = nullptr;
  if (cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming) != cudaSuccess) {
    rt_fail("mrec_rt_event_create", cudaGetLastError());
    return nullptr;
  }
  return e;
}
MREC_API int mrec_rt_event_record(void* e, void* s) { RT_CHECK(cudaEventRecord((cudaEvent_t)e, (cudaStream_t)s), "mrec_rt_event_record"); return OK; }
MREC_API int mrec_rt_event_sync(void* e) { RT_CHECK(cudaEventSynchronize((cudaEvent_t)e), "mrec_rt_event_sync"); return OK; }
MREC_API int mrec_rt_stream_wait_event(void* s, void* e) { RT_CHECK(cudaStreamWaitEvent((cudaStream_t)s, (cudaEvent_t)e, 0), "mrec_rt_stream_wait_event"); return OK; }
MREC_API float mrec_rt_event_elapsed_ms(void* a, void* b) {
  float ms = -1.f;
  if (cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b) != cudaSuccess) { cudaGetLastError(); return -1.f; }
  return ms;
}
MREC_API int mrec_rt_event_destroy(void* e) { RT_CHECK(cudaEventDestroy((cudaEvent_t)e), "mrec_rt_event_destroy"); return OK; }
// Stream capture -> executable graph (the aot kernels only enqueue, so a whole step can be captured)
MREC_API int mrec_rt_graph_begin(void* s) {
  RT_CHECK(cudaStreamBeginCapture((cudaStream_t)s, cudaStreamCaptureModeThreadLocal), "mrec_rt_graph_begin");
  return OK;
}
MREC_API void* mrec_rt_graph_end(void* s) {
  cudaGraph_t g = nullptr;
  if (cudaStreamEndCapture((cudaStream_t)s, &g) != cudaSuccess || !g) { rt_fail("mrec_rt_graph_end", cudaGetLastError()); return nullptr; }
  cudaGraphExec_t x = nullptr;
  const cudaError_t e = cudaGraphInstantiate(&x, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) { rt_fail("mrec_rt_graph_end: instantiate", e); return nullptr; }
  return x;
}
MREC_API int mrec_rt_graph_launch(void* x, void* s) { RT_CHECK(cudaGraphLaunch((cudaGraphExec_t)x, (cudaStream_t)s), "mrec_rt_graph_launch"); return OK; }
MREC_API int mrec_rt_graph_destroy(void* x) { RT_CHECK(cudaGraphExecDestroy((cudaGraphExec_t)x), "mrec_rt_graph_destroy"); return OK; }
