// DenseLayer GEMMs for the torch-free host path (not aot: a host entry point next to mrec_rt_*).
//
// SURVEY 2b scopes the DenseLayer MLP of the reference (wide_and_deep.py:72-133: MatMul + BiasAdd + ReLU, fp16 under
// use_mixed_precision) as a library call; under MindSpore the framework runs it, in this repository's model cells torch
// does.  `mindrec_b200.runtime` has neither, so the library binds cuBLASLt itself — loaded at run time with dlopen (no
// link-time dependency: a process that only uses the aot kernels never touches it, and a process that already holds a
// cuBLASLt, e.g. torch's, shares that copy).  One entry point:
//
//   mrec_rt_gemm: row-major  C[M,N] = act(alpha * op(A) op(B) + beta * C + bias[N])
//     A, B: fp16 or fp32 (same type), C: fp16 or fp32, fp32 accumulation, epilogue none | bias | relu(bias)
//
// cuBLASLt is column-major: C^T[N,M] = op(B)^T op(A)^T, i.e. the operands swap places and keep their transposes; the
// bias epilogue then runs along the N output features, which is what BiasAdd does.
#include "common.cuh"

#include <cublasLt.h>
#include <dlfcn.h>

#include <map>
#include <mutex>
#include <tuple>

namespace mrec {
namespace {

struct LtApi {
  void* so = nullptr;
  decltype(&cublasLtCreate) Create = nullptr;
  decltype(&cublasLtMatmul) Matmul = nullptr;
  decltype(&cublasLtMatmulDescCreate) DescCreate = nullptr;
  decltype(&cublasLtMatmulDescSetAttribute) DescSet = nullptr;
  decltype(&cublasLtMatmulDescDestroy) DescDestroy = nullptr;
  decltype(&cublasLtMatrixLayoutCreate) LayoutCreate = nullptr;
  decltype(&cublasLtMatrixLayoutDestroy) LayoutDestroy = nullptr;
  decltype(&cublasLtMatmulPreferenceCreate) PrefCreate = nullptr;
  decltype(&cublasLtMatmulPreferenceSetAttribute) PrefSet = nullptr;
  decltype(&cublasLtMatmulPreferenceDestroy) PrefDestroy = nullptr;
  decltype(&cublasLtMatmulAlgoGetHeuristic) Heuristic = nullptr;
  cublasLtHandle_t handle = nullptr;
  void* workspace = nullptr;
  size_t workspace_bytes = 0;
  bool ok = false;
};

constexpr size_t kLtWorkspace = 64u << 20;

LtApi& lt() {
  static LtApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libcublasLt.so.12", "libcublasLt.so.13", "libcublasLt.so"}) {
      api.so = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.so) break;
    }
    if (!api.so) return;
#define MREC_LT_SYM(field, sym)                                          \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.so, sym)); \
  if (!api.field) return;
    MREC_LT_SYM(Create, "cublasLtCreate")
    MREC_LT_SYM(Matmul, "cublasLtMatmul")
    MREC_LT_SYM(DescCreate, "cublasLtMatmulDescCreate")
    MREC_LT_SYM(DescSet, "cublasLtMatmulDescSetAttribute")
    MREC_LT_SYM(DescDestroy, "cublasLtMatmulDescDestroy")
    MREC_LT_SYM(LayoutCreate, "cublasLtMatrixLayoutCreate")
    MREC_LT_SYM(LayoutDestroy, "cublasLtMatrixLayoutDestroy")
    MREC_LT_SYM(PrefCreate, "cublasLtMatmulPreferenceCreate")
    MREC_LT_SYM(PrefSet, "cublasLtMatmulPreferenceSetAttribute")
    MREC_LT_SYM(PrefDestroy, "cublasLtMatmulPreferenceDestroy")
    MREC_LT_SYM(Heuristic, "cublasLtMatmulAlgoGetHeuristic")
#undef MREC_LT_SYM
    if (api.Create(&api.handle) != CUBLAS_STATUS_SUCCESS) return;
    if (cudaMalloc(&api.workspace, kLtWorkspace) != cudaSuccess) {
      cudaGetLastError();
      return;
    }
    api.workspace_bytes = kLtWorkspace;
    api.ok = true;
  });
  return api;
}

// (M, N, K, transA, transB, a16, c16, epilogue, beta != 0) -> the algorithm the heuristic picked for that problem
using AlgoKey = std::tuple<int64_t, int64_t, int64_t, int, int, int, int, int, int>;
std::map<AlgoKey, cublasLtMatmulAlgo_t>& algo_cache() {
  static std::map<AlgoKey, cublasLtMatmulAlgo_t> m;
  return m;
}
std::mutex& algo_mutex() {
  static std::mutex m;
  return m;
}

}  // namespace
}  // namespace mrec
using namespace mrec;

// Row-major C[M,N] = act(alpha * op(A) op(B) + beta * C + bias).  op(A) is [M,K]: A is stored [M,K] (trans_a = 0) or
// [K,M] (trans_a = 1); op(B) is [K,N]: B is stored [K,N] or [N,K].  ab_half / c_half: fp16 (1) or fp32 (0) storage;
// accumulation is always fp32.  epilogue: 0 none, 1 + bias[N], 2 relu(+ bias[N]); bias has C's type.  Asynchronous on
// `stream`, capturable.  Returns 0 or an MREC_ERR_* code (mrec_last_error has the text).
MREC_API int mrec_rt_gemm(const void* a, const void* b, void* c, const void* bias, int64_t m, int64_t n, int64_t k,
                          int trans_a, int trans_b, int ab_half, int c_half, float alpha, float beta, int epilogue,
                          void* stream) {
  if (m <= 0 || n <= 0 || k <= 0) return OK;
  if (!a || !b || !c || (epilogue != 0 && !bias)) return fail(ERR_NULL, "mrec_rt_gemm: null operand");
  if (epilogue < 0 || epilogue > 2) return fail(ERR_SHAPE, "mrec_rt_gemm: epilogue must be 0, 1 or 2");
  LtApi& L = lt();
  if (!L.ok) return fail(ERR_CUDA, "mrec_rt_gemm: cuBLASLt is not available (%s)", L.so ? "initialisation failed" : "dlopen failed");
  const cudaDataType_t ab_t = ab_half ? CUDA_R_16F : CUDA_R_32F;
  const cudaDataType_t c_t = c_half ? CUDA_R_16F : CUDA_R_32F;
  cublasLtMatmulDesc_t desc = nullptr;
  cublasLtMatrixLayout_t la = nullptr, lb = nullptr, lc = nullptr;
  int rc = OK;
  auto bail = [&](const char* what, cublasStatus_t st) {
    rc = fail(ERR_CUDA, "mrec_rt_gemm: %s failed (cublas status %d)", what, (int)st);
  };
  cublasStatus_t st = L.DescCreate(&desc, CUBLAS_COMPUTE_32F, CUDA_R_32F);
  if (st != CUBLAS_STATUS_SUCCESS) return fail(ERR_CUDA, "mrec_rt_gemm: cublasLtMatmulDescCreate failed (%d)", (int)st);
  // column-major problem: C^T[N,M] = op'(B)[N,K] op'(A)[K,M]; a row-major [r,c] array is a column-major [c,r] one
  const cublasOperation_t op_first = trans_b ? CUBLAS_OP_T : CUBLAS_OP_N;   // applied to B as stored
  const cublasOperation_t op_second = trans_a ? CUBLAS_OP_T : CUBLAS_OP_N;  // applied to A as stored
  L.DescSet(desc, CUBLASLT_MATMUL_DESC_TRANSA, &op_first, sizeof(op_first));
  L.DescSet(desc, CUBLASLT_MATMUL_DESC_TRANSB, &op_second, sizeof(op_second));
  if (epilogue) {
    const cublasLtEpilogue_t ep = epilogue == 2 ? CUBLASLT_EPILOGUE_RELU_BIAS : CUBLASLT_EPILOGUE_BIAS;
    L.DescSet(desc, CUBLASLT_MATMUL_DESC_EPILOGUE, &ep, sizeof(ep));
    L.DescSet(desc, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bias, sizeof(bias));
    L.DescSet(desc, CUBLASLT_MATMUL_DESC_BIAS_DATA_TYPE, &c_t, sizeof(c_t));
  }
  // B as stored: row-major [K,N] (or [N,K] when trans_b) = column-major [N,K] (or [K,N]), leading dimension = its row length
  const int64_t b_rows = trans_b ? k : n, b_cols = trans_b ? n : k;
  const int64_t a_rows = trans_a ? m : k, a_cols = trans_a ? k : m;
  if ((st = L.LayoutCreate(&la, ab_t, b_rows, b_cols, b_rows)) != CUBLAS_STATUS_SUCCESS) bail("layout(B)", st);
  if (!rc && (st = L.LayoutCreate(&lb, ab_t, a_rows, a_cols, a_rows)) != CUBLAS_STATUS_SUCCESS) bail("layout(A)", st);
  if (!rc && (st = L.LayoutCreate(&lc, c_t, n, m, n)) != CUBLAS_STATUS_SUCCESS) bail("layout(C)", st);
  if (!rc) {
    const AlgoKey key{m, n, k, trans_a, trans_b, ab_half, c_half, epilogue, beta != 0.f ? 1 : 0};
    cublasLtMatmulAlgo_t algo;
    bool have = false;
    {
      std::lock_guard<std::mutex> g(algo_mutex());
      auto it = algo_cache().find(key);
      if (it != algo_cache().end()) {
        algo = it->second;
        have = true;
      }
    }
    if (!have) {
      cublasLtMatmulPreference_t pref = nullptr;
      cublasLtMatmulHeuristicResult_t res;
      int found = 0;
      st = L.PrefCreate(&pref);
      if (st == CUBLAS_STATUS_SUCCESS) {
        L.PrefSet(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &L.workspace_bytes, sizeof(L.workspace_bytes));
        st = L.Heuristic(L.handle, desc, la, lb, lc, lc, pref, 1, &res, &found);
        L.PrefDestroy(pref);
      }
      if (st != CUBLAS_STATUS_SUCCESS || found == 0) {
        bail("cublasLtMatmulAlgoGetHeuristic", st);
      } else {
        algo = res.algo;
        std::lock_guard<std::mutex> g(algo_mutex());
        algo_cache()[key] = algo;
      }
    }
    if (!rc) {
      st = L.Matmul(L.handle, desc, &alpha, b, la, a, lb, &beta, c, lc, c, lc, &algo, L.workspace, L.workspace_bytes,
                    (cudaStream_t)stream);
      if (st != CUBLAS_STATUS_SUCCESS) bail("cublasLtMatmul", st);      // (library kernels: not in mrec_launch_count)
    }
  }
  if (la) L.LayoutDestroy(la);
  if (lb) L.LayoutDestroy(lb);
  if (lc) L.LayoutDestroy(lc);
  L.DescDestroy(desc);
  return rc;
}

namespace mrec {
__global__ void cast_f32_f16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, int64_t n4, const float* __restrict__ src_tail,
                                    __half* __restrict__ dst_tail, int tail) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = src[i];
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&lo);
    u.y = *reinterpret_cast<const uint32_t*>(&hi);
    dst[i] = u;
  }
  if (blockIdx.x == 0 && threadIdx.x < tail) dst_tail[threadIdx.x] = __float2half_rn(src_tail[threadIdx.x]);
}
}  // namespace mrec

// The per-step Cast(weight, float16) of the mixed-precision DenseLayers (wide_and_deep.py:119-122), for the whole flat
// parameter buffer at once.     in : src[n] f32        out: dst[n] f16
MREC_API int mrec_cast_f32_f16(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes, void* stream,
                               void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  if (a.nparam != 2) return fail(ERR_NPARAM, "mrec_cast_f32_f16: expected 2 params, got %d", a.nparam);
  MREC_REQUIRE(a.is_f32(0) && a.is(1, "float16") && a.numel(0) == a.numel(1), ERR_DTYPE, "mrec_cast_f32_f16: src[n] f32, dst[n] f16");
  MREC_REQUIRE(a.aligned(0, 16) && a.aligned(1, 8), ERR_ALIGN, "mrec_cast_f32_f16: buffers must be 16- / 8-byte aligned");
  const int64_t n = a.numel(0);
  if (n == 0) return OK;
  const int64_t n4 = n / 4;
  MREC_LAUNCH(cast_f32_f16_kernel, grid_for(cdiv(n4 > 0 ? n4 : 1, 256), 8), 256, 0, a.stream, a.ptr<float4>(0), a.ptr<uint2>(1), n4,
              a.ptr<float>(0) + n4 * 4, a.ptr<__half>(1) + n4 * 4, (int)(n - n4 * 4));
  return check_launch("cast_f32_f16");
}
