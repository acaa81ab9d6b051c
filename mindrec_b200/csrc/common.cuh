// Shared device/host helpers for libmindrec_b200.so (sm_100a only).
//
// The library is entered only through the MindSpore ops.Custom(func_type="aot")
// C-ABI declared in include/mindrec_b200.h.  Everything here is internal.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#define MREC_API extern "C" __attribute__((visibility("default")))

namespace mrec {

// ---- error codes returned across the C-ABI (include/mindrec_b200.h mirrors them) ----
enum : int {
  OK = 0,
  ERR_NPARAM = 1,
  ERR_DTYPE = 2,
  ERR_SHAPE = 3,
  ERR_ALIGN = 4,
  ERR_DIM = 5,
  ERR_CUDA = 6,
  ERR_WORKSPACE = 7,
  ERR_NULL = 8,
};

extern thread_local char g_last_error[512];
extern std::atomic<unsigned long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

// ---- aot argument pack ----
struct Aot {
  int nparam;
  void** params;
  int* ndims;
  int64_t** shapes;
  const char** dtypes;
  cudaStream_t stream;

  bool is(int i, const char* dt) const { return dtypes[i] && strcmp(dtypes[i], dt) == 0; }
  bool is_f32(int i) const { return is(i, "float32"); }
  bool is_i32(int i) const { return is(i, "int32"); }
  bool is_i64(int i) const { return is(i, "int64"); }
  int64_t numel(int i) const {
    int64_t n = 1;
    for (int k = 0; k < ndims[i]; ++k) n *= shapes[i][k];
    return n;
  }
  int64_t dim(int i, int k) const { return (k < ndims[i]) ? shapes[i][k] : 1; }
  // product of all dims but the last
  int64_t rows(int i) const {
    int64_t n = 1;
    for (int k = 0; k + 1 < ndims[i]; ++k) n *= shapes[i][k];
    return n;
  }
  int64_t last(int i) const { return ndims[i] > 0 ? shapes[i][ndims[i] - 1] : 1; }
  template <typename T>
  T* ptr(int i) const { return reinterpret_cast<T*>(params[i]); }
  bool aligned(int i, size_t a) const { return (reinterpret_cast<uintptr_t>(params[i]) % a) == 0; }
};

#define MREC_CHECK_NPARAM(a, n)                                                     \
  do {                                                                              \
    if ((a).nparam != (n)) return fail(ERR_NPARAM, "%s: expected %d params, got %d", \
                                       __func__, (n), (a).nparam);                  \
    for (int _i = 0; _i < (n); ++_i)                                                \
      if ((a).params[_i] == nullptr && (a).numel(_i) > 0)                           \
        return fail(ERR_NULL, "%s: param %d is a null pointer", __func__, _i);      \
  } while (0)

#define MREC_REQUIRE(cond, code, ...)          \
  do {                                         \
    if (!(cond)) return fail((code), __VA_ARGS__); \
  } while (0)

inline int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(ERR_CUDA, "%s: CUDA launch error: %s", what, cudaGetErrorString(e));
  }
  return OK;
}

#define MREC_LAUNCH(kernel, grid, block, smem, stream, ...)              \
  do {                                                                   \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);          \
    mrec::g_launches.fetch_add(1, std::memory_order_relaxed);            \
  } while (0)

constexpr int kNumSMs = 148;  // B200

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
// grid size: enough blocks for `work_blocks`, capped at `per_sm` resident blocks per SM
inline int grid_for(int64_t work_blocks, int per_sm) {
  int64_t cap = (int64_t)kNumSMs * per_sm;
  int64_t g = work_blocks < cap ? work_blocks : cap;
  return (int)(g < 1 ? 1 : g);
}

#ifdef __CUDACC__
// ---- 128-bit streaming loads / stores ----
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
// peer-mapped data written by ANOTHER GPU: system-scope relaxed load (never served from a non-coherent cache)
__device__ __forceinline__ float4 ld_volatile_f4(const float4* p) {
  float4 r;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float ld_stream_f1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// ---- 256-bit accesses (sm_100: LDG.E.256 / STG.E.256), L2 evict-first: for rows that are read-modify-written
// once and not needed again soon (optimizer state), so they do not push reusable lines out of the L2 ----
struct alignas(32) F8 { float4 lo, hi; };
__device__ __forceinline__ F8 ld256_evict_first(const F8* p) {
  F8 r;
  asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st256_evict_first(F8* p, const F8& v) {
  asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v.lo.x),
               "f"(v.lo.y), "f"(v.lo.z), "f"(v.lo.w), "f"(v.hi.x), "f"(v.hi.y), "f"(v.hi.z), "f"(v.hi.w)
               : "memory");
}

// Scheduling fence: an empty volatile asm that "rewrites" the registers of a loaded value.  Volatile asms keep
// their program order, so placing these after a batch of (volatile) loads forces every load of the batch to
// be issued before the first use of any of them — without it nvcc interleaves load/use pairs to save
// registers and a thread has only one or two loads in flight (seen in SASS and in ncu's source view).
__device__ __forceinline__ void reg_fence(float4& v) {
  asm volatile("" : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w));
}
__device__ __forceinline__ void reg_fence(float& v) { asm volatile("" : "+f"(v)); }
__device__ __forceinline__ void reg_fence(uint2& v) { asm volatile("" : "+r"(v.x), "+r"(v.y)); }
__device__ __forceinline__ void reg_fence(int& v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void reg_fence(F8& v) {
  asm volatile("" : "+f"(v.lo.x), "+f"(v.lo.y), "+f"(v.lo.z), "+f"(v.lo.w), "+f"(v.hi.x), "+f"(v.hi.y), "+f"(v.hi.z), "+f"(v.hi.w));
}
__device__ __forceinline__ void reg_fence(uint4& v) { asm volatile("" : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w)); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS: UBLKCP) ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(phase)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  while (!mbar_try_wait(bar, phase)) {
  }
}
// global -> shared bulk copy; bytes must be a multiple of 16, both addresses 16-B aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_scale(const float4& a, float s) {
  return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
__device__ __forceinline__ void f4_fma(float4& acc, const float4& a, float s) {
  acc.x = fmaf(a.x, s, acc.x);
  acc.y = fmaf(a.y, s, acc.y);
  acc.z = fmaf(a.z, s, acc.z);
  acc.w = fmaf(a.w, s, acc.w);
}
#endif  // __CUDACC__

}  // namespace mrec
