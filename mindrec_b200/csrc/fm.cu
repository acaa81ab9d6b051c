// K7 — DeepFM second-order interaction, forward and backward, as fused memory-bound reductions.
//
// Replaces the six separate elementwise / reduce ops of models/deepfm/src/deepfm.py:222-228
//     v1 = Square(ReduceSum(vx, 1)); v2 = ReduceSum(Square(vx), 1); fm = 0.5 * ReduceSum(v1 - v2, 1)
// and their autodiff:  d fm / d vx[b,f,d] = g[b] * (S[b,d] - vx[b,f,d]),  S = sum_f vx.
//
// Mapping: a group of D/4 threads owns one sample, each thread one float4 column chunk; it walks the F
// fields with all F loads in flight (39 x 16 B per thread), keeps S and sum(vx^2) for its 4 columns in
// registers, and the D/4 partial results of a sample are added in column order through shared memory
// (fixed order: deterministic).  vx is read once from HBM by the forward (B*F*D*4 bytes); the backward
// re-walks the fields for the store pass out of L1/L2 (the sample's 2.5 KB..12 KB were just read by the
// same threads) and writes dvx once.
#include "common.cuh"
#include <cuda_fp16.h>

namespace mrec {

constexpr int kFmThreads = 256;
constexpr int kFmBatch = 13;  // field loads in flight per thread per trip (39 = 3 trips)

template <typename Vec> struct FmV;
template <> struct FmV<float4> {
  static __device__ __forceinline__ float4 ld(const float* p, int64_t i) {
    return ld_stream_f4(reinterpret_cast<const float4*>(p) + i);
  }
  static __device__ __forceinline__ float4 ld_cached(const float* p, int64_t i) {
    float4 r;
    asm volatile("ld.global.ca.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(reinterpret_cast<const float4*>(p) + i));
    return r;
  }
  static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ void acc(float4& s, float4& s2, const float4& x) {
    s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
    s2.x = fmaf(x.x, x.x, s2.x); s2.y = fmaf(x.y, x.y, s2.y);
    s2.z = fmaf(x.z, x.z, s2.z); s2.w = fmaf(x.w, x.w, s2.w);
  }
  static __device__ __forceinline__ float finish(const float4& s, const float4& s2) {
    return (s.x * s.x - s2.x) + (s.y * s.y - s2.y) + (s.z * s.z - s2.z) + (s.w * s.w - s2.w);
  }
  // dvx = g * (S - vx) (+ addend: the DenseLayer-path gradient of the same vx, fp32 or fp16)
  static __device__ __forceinline__ void st_grad(float* p, int64_t i, float g, const float4& s,
                                                 const float4& x, const void* add, bool add16) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (add) {
      if (add16) {
        const uint2 u = reinterpret_cast<const uint2*>(add)[i];
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
        const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        a = make_float4(lo.x, lo.y, hi.x, hi.y);
      } else {
        a = ld_stream_f4(reinterpret_cast<const float4*>(add) + i);
      }
    }
    st_stream_f4(reinterpret_cast<float4*>(p) + i,
                 make_float4(fmaf(g, s.x - x.x, a.x), fmaf(g, s.y - x.y, a.y), fmaf(g, s.z - x.z, a.z),
                             fmaf(g, s.w - x.w, a.w)));
  }
};
template <> struct FmV<float> {
  static __device__ __forceinline__ float ld(const float* p, int64_t i) { return ld_stream_f1(p + i); }
  static __device__ __forceinline__ float ld_cached(const float* p, int64_t i) { return p[i]; }
  static __device__ __forceinline__ float zero() { return 0.f; }
  static __device__ __forceinline__ void acc(float& s, float& s2, const float& x) {
    s += x;
    s2 = fmaf(x, x, s2);
  }
  static __device__ __forceinline__ float finish(const float& s, const float& s2) { return s * s - s2; }
  static __device__ __forceinline__ void st_grad(float* p, int64_t i, float g, const float& s,
                                                 const float& x, const void* add, bool add16) {
    float a = 0.f;
    if (add) a = add16 ? __half2float(reinterpret_cast<const __half*>(add)[i]) : reinterpret_cast<const float*>(add)[i];
    p[i] = fmaf(g, s - x, a);
  }
};

// groups of `cpr` threads per sample; samples_per_block = kFmThreads / cpr
template <typename Vec, bool BACKWARD>
__global__ void __launch_bounds__(kFmThreads)
fm_kernel(const float* __restrict__ vx, const float* __restrict__ gout, float* __restrict__ out,
          float* __restrict__ dvx, int64_t batch, int fields, int cpr, const void* __restrict__ addend,
          bool add16) {
  __shared__ float s_part[kFmThreads];
  const int spb = kFmThreads / cpr;
  const int gi = threadIdx.x / cpr;
  const int c = threadIdx.x - gi * cpr;
  for (int64_t b0 = (int64_t)blockIdx.x * spb; b0 < batch; b0 += (int64_t)gridDim.x * spb) {
    const int64_t b = b0 + gi;
    const bool live = (gi < spb) && (b < batch);
    Vec s = FmV<Vec>::zero(), s2 = FmV<Vec>::zero();
    const int64_t base = b * fields * cpr + c;
    if (live) {
      for (int f0 = 0; f0 < fields; f0 += kFmBatch) {
        // fields past the end are clamped to the last one (and skipped below): all loads unconditional
        Vec x[kFmBatch];
#pragma unroll
        for (int k = 0; k < kFmBatch; ++k) {
          const int f = min(f0 + k, fields - 1);
          x[k] = BACKWARD ? FmV<Vec>::ld_cached(vx, base + (int64_t)f * cpr)  // keep in L1 for pass 2
                          : FmV<Vec>::ld(vx, base + (int64_t)f * cpr);
        }
#pragma unroll
        for (int k = 0; k < kFmBatch; ++k) reg_fence(x[k]);
#pragma unroll
        for (int k = 0; k < kFmBatch; ++k)
          if (f0 + k < fields) FmV<Vec>::acc(s, s2, x[k]);
      }
    }
    if (!BACKWARD) {
      s_part[threadIdx.x] = live ? FmV<Vec>::finish(s, s2) : 0.f;
      __syncthreads();
      if (live && c == 0) {
        float t = 0.f;
        for (int q = 0; q < cpr; ++q) t += s_part[gi * cpr + q];
        out[b] = 0.5f * t;
      }
      __syncthreads();
    } else if (live) {
      const float g = gout[b];
      for (int f0 = 0; f0 < fields; f0 += kFmBatch) {
        Vec x[kFmBatch];
#pragma unroll
        for (int k = 0; k < kFmBatch; ++k)
          x[k] = FmV<Vec>::ld_cached(vx, base + (int64_t)min(f0 + k, fields - 1) * cpr);
#pragma unroll
        for (int k = 0; k < kFmBatch; ++k) reg_fence(x[k]);
#pragma unroll
        for (int k = 0; k < kFmBatch; ++k)
          if (f0 + k < fields) FmV<Vec>::st_grad(dvx, base + (int64_t)(f0 + k) * cpr, g, s, x[k], addend, add16);
      }
    }
  }
}

static int fm_common(const Aot& a, int vx_i, int64_t* batch, int* fields, int* dim) {
  MREC_REQUIRE(a.is_f32(vx_i), ERR_DTYPE, "mrec_fm: vx must be float32");
  MREC_REQUIRE(a.ndims[vx_i] == 3, ERR_SHAPE, "mrec_fm: vx must be [B,F,D]");
  *batch = a.dim(vx_i, 0);
  *fields = (int)a.dim(vx_i, 1);
  *dim = (int)a.dim(vx_i, 2);
  MREC_REQUIRE(*dim >= 1 && *fields >= 1, ERR_DIM, "mrec_fm: F and D must be >= 1");
  const int cpr = (*dim % 4 == 0) ? *dim / 4 : *dim;
  MREC_REQUIRE(cpr <= kFmThreads, ERR_DIM, "mrec_fm: D too large");
  if (*dim % 4 == 0) MREC_REQUIRE(a.aligned(vx_i, 16), ERR_ALIGN, "mrec_fm: vx must be 16-byte aligned");
  return OK;
}

}  // namespace mrec

using namespace mrec;

// in : vx[B,F,D] f32 (already multiplied by the mask)          out: fm[B] | [B,1] f32
MREC_API int mrec_fm_fwd(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                         void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  MREC_CHECK_NPARAM(a, 2);
  int64_t batch; int fields, dim;
  int rc = fm_common(a, 0, &batch, &fields, &dim);
  if (rc) return rc;
  MREC_REQUIRE(a.is_f32(1) && a.numel(1) == batch, ERR_SHAPE, "mrec_fm_fwd: out must be f32 with B elements");
  if (batch == 0) return OK;
  if (dim % 4 == 0) {
    const int cpr = dim / 4, spb = kFmThreads / cpr;
    MREC_LAUNCH((fm_kernel<float4, false>), grid_for(cdiv(batch, spb), 8), kFmThreads, 0, a.stream,
                a.ptr<float>(0), nullptr, a.ptr<float>(1), nullptr, batch, fields, cpr, nullptr, false);
  } else {
    const int spb = kFmThreads / dim;
    MREC_LAUNCH((fm_kernel<float, false>), grid_for(cdiv(batch, spb), 8), kFmThreads, 0, a.stream,
                a.ptr<float>(0), nullptr, a.ptr<float>(1), nullptr, batch, fields, dim, nullptr, false);
  }
  return check_launch("fm_fwd");
}

// in : vx[B,F,D] f32, gout[B] | [B,1] f32, (addend[B,F,D] f32|f16)     out: dvx[B,F,D] f32
// dvx = gout * (sum_f vx - vx) + addend   (addend = gradient reaching the same vx through the DenseLayers)
MREC_API int mrec_fm_bwd(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                         void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  if (a.nparam != 3 && a.nparam != 4)
    return fail(ERR_NPARAM, "mrec_fm_bwd: expected 3 or 4 params, got %d", a.nparam);
  for (int i = 0; i < a.nparam; ++i)
    if (!a.params[i] && a.numel(i) > 0) return fail(ERR_NULL, "mrec_fm_bwd: param %d is null", i);
  const int o = a.nparam - 1;
  int64_t batch; int fields, dim;
  int rc = fm_common(a, 0, &batch, &fields, &dim);
  if (rc) return rc;
  MREC_REQUIRE(a.is_f32(1) && a.numel(1) == batch, ERR_SHAPE, "mrec_fm_bwd: gout must be f32 with B elements");
  MREC_REQUIRE(a.is_f32(o) && a.numel(o) == a.numel(0), ERR_SHAPE, "mrec_fm_bwd: dvx must match vx");
  const void* addend = nullptr;
  bool add16 = false;
  if (a.nparam == 4) {
    MREC_REQUIRE((a.is_f32(2) || a.is(2, "float16")) && a.numel(2) == a.numel(0), ERR_SHAPE,
                 "mrec_fm_bwd: addend must be f32|f16 with vx's shape");
    addend = a.params[2];
    add16 = a.is(2, "float16");
  }
  if (batch == 0) return OK;
  if (dim % 4 == 0) {
    MREC_REQUIRE(a.aligned(o, 16) && (!addend || a.aligned(2, add16 ? 8 : 16)), ERR_ALIGN,
                 "mrec_fm_bwd: dvx/addend must be 16-byte aligned");
    const int cpr = dim / 4, spb = kFmThreads / cpr;
    MREC_LAUNCH((fm_kernel<float4, true>), grid_for(cdiv(batch, spb), 8), kFmThreads, 0, a.stream,
                a.ptr<float>(0), a.ptr<float>(1), nullptr, a.ptr<float>(o), batch, fields, cpr, addend, add16);
  } else {
    const int spb = kFmThreads / dim;
    MREC_LAUNCH((fm_kernel<float, true>), grid_for(cdiv(batch, spb), 8), kFmThreads, 0, a.stream,
                a.ptr<float>(0), a.ptr<float>(1), nullptr, a.ptr<float>(o), batch, fields, dim, addend, add16);
  }
  return check_launch("fm_bwd");
}
