// Non-aot helper entry points of libmindrec_b200.so (see include/mindrec_b200.h).
#include "common.cuh"

namespace mrec {
thread_local char g_last_error[512] = {0};
std::atomic<unsigned long long> g_launches{0};
}  // namespace mrec

MREC_API const char* mrec_version(void) { return "mindrec_b200 0.1.0 (sm_100a)"; }

MREC_API const char* mrec_last_error(void) { return mrec::g_last_error; }

// Number of kernels this library has launched in this process (bench.py's gpu_launches evidence).
MREC_API unsigned long long mrec_launch_count(void) {
  return mrec::g_launches.load(std::memory_order_relaxed);
}
