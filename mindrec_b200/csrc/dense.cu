// DenseLayer backward glue — ReLU gradient mask fused with BiasAddGrad.
//
// Replaces, per DenseLayer of models/wide_deep/src/wide_and_deep.py:72-133 (bprop of BiasAdd + ReLU, also
// deepfm.py:150-170 and deep_and_cross.py:161-200):  gz = g * (y > 0)  (ReluGrad) and  gb = sum_b gz[b, :]
// (BiasAddGrad).  As separate library ops these were an element-wise kernel, a ones-vector GEMM on 8 CTAs
// (13.5 us per layer, ncu r1g) and a cast; here one pass reads g and y once, writes gz in place and leaves the
// fp32 column sums in the flat gradient buffer.
//
// Grid (row blocks, column tiles); a CTA is tx column chunks (16 bytes each) x ty row lanes.  Column sums are
// reduced lane -> CTA -> grid in a fixed order: per-CTA partials go to a workspace and colsum_finish_kernel adds
// them per column (a "last CTA adds them up" ticket variant cost 15-20 us of serial tail per call, r1i kbench),
// so the result is bit-reproducible.
#include "common.cuh"
#include <cstdlib>
#include <cuda_fp16.h>

namespace mrec {

constexpr int kDenseThreads = 256;
constexpr int kDenseMaxRowBlocks = 256;   // row blocks per column tile (partials the last CTA adds up)
constexpr int kDenseTargetCtas = 8 * kNumSMs;

template <typename T> struct DChunk;
template <> struct DChunk<__half> {
  static constexpr int kVec = 8;
  using Raw = uint4;
  static __device__ __forceinline__ void acc(float (&a)[8], Raw& g, const Raw& y, bool mask) {
    __half2* gh = reinterpret_cast<__half2*>(&g);
    const __half2* yh = reinterpret_cast<const __half2*>(&y);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 gf = __half22float2(gh[k]);
      if (mask) {
        const float2 yf = __half22float2(yh[k]);
        if (!(yf.x > 0.f)) gf.x = 0.f;
        if (!(yf.y > 0.f)) gf.y = 0.f;
        gh[k] = __floats2half2_rn(gf.x, gf.y);   // exact: a kept value is unchanged, a dropped one is +0
      }
      a[2 * k] += gf.x;
      a[2 * k + 1] += gf.y;
    }
  }
};
template <> struct DChunk<float> {
  static constexpr int kVec = 4;
  using Raw = float4;
  static __device__ __forceinline__ void acc(float (&a)[4], Raw& g, const Raw& y, bool mask) {
    if (mask) {
      if (!(y.x > 0.f)) g.x = 0.f;
      if (!(y.y > 0.f)) g.y = 0.f;
      if (!(y.z > 0.f)) g.z = 0.f;
      if (!(y.w > 0.f)) g.w = 0.f;
    }
    a[0] += g.x; a[1] += g.y; a[2] += g.z; a[3] += g.w;
  }
};

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
__device__ __forceinline__ void from_f(float& d, float v) { d = v; }
__device__ __forceinline__ void from_f(__half& d, float v) { d = __float2half_rn(v); }

// VEC = true: 16-byte chunks (n_cols % kVec == 0, 16-byte aligned rows); VEC = false: one column per thread.
template <typename T, bool VEC>
__global__ void __launch_bounds__(kDenseThreads)
relu_bwd_bias_kernel(const T* __restrict__ g, const T* __restrict__ y, T* __restrict__ gz, int64_t rows, int n_cols,
                     int rows_per_cta, float* __restrict__ partial) {
  constexpr int V = VEC ? DChunk<T>::kVec : 1;
  __shared__ float s_red[kDenseThreads * V];
  const int tx = threadIdx.x, ty = threadIdx.y, ntx = blockDim.x, nty = blockDim.y;
  const int chunks = n_cols / V;
  const int ch = blockIdx.y * ntx + tx;
  const bool active = ch < chunks;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(rows, r0 + rows_per_cta);
  const bool mask = (y != nullptr);
  float a[V];
#pragma unroll
  for (int k = 0; k < V; ++k) a[k] = 0.f;

  if (active) {
    if constexpr (VEC) {
      using Raw = typename DChunk<T>::Raw;
      const Raw* g4 = reinterpret_cast<const Raw*>(g);
      const Raw* y4 = reinterpret_cast<const Raw*>(y);
      Raw* z4 = reinterpret_cast<Raw*>(gz);
      constexpr int U = 4;                               // rows in flight per thread
      for (int64_t r = r0 + ty; r < r1; r += (int64_t)U * nty) {
        Raw gv[U], yv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t rr = min(r + (int64_t)u * nty, r1 - 1);   // clamped: loads stay unconditional
          gv[u] = g4[rr * chunks + ch];
          yv[u] = mask ? y4[rr * chunks + ch] : gv[u];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t rr = r + (int64_t)u * nty;
          if (rr < r1) {
            DChunk<T>::acc(a, gv[u], yv[u], mask);
            if (mask && z4) z4[rr * chunks + ch] = gv[u];
          }
        }
      }
    } else {
      for (int64_t r = r0 + ty; r < r1; r += nty) {
        float v = to_f(g[r * n_cols + ch]);
        if (mask) {
          if (!(to_f(y[r * n_cols + ch]) > 0.f)) v = 0.f;
          if (gz) from_f(gz[r * n_cols + ch], v);
        }
        a[0] += v;
      }
    }
  }
  // lanes -> CTA, in lane order
  const int slot = (ty * ntx + tx) * V;
#pragma unroll
  for (int k = 0; k < V; ++k) s_red[slot + k] = a[k];
  __syncthreads();
  if (ty == 0 && active) {
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float t = s_red[tx * V + k];
      for (int l = 1; l < nty; ++l) t += s_red[(l * ntx + tx) * V + k];
      partial[(int64_t)blockIdx.x * n_cols + ch * V + k] = t;
    }
  }
}

// partial[plane][rb][n_cols] -> out[plane][n_cols]: CTA = 32 columns x 32 row-block lanes, fixed order.
struct ColsumOut { float* p[3]; int n[3]; };
constexpr int kFinishLanes = 32;

__global__ void __launch_bounds__(32 * kFinishLanes)
colsum_finish_kernel(const float* __restrict__ partial, int row_blocks, int n_cols, ColsumOut out) {
  __shared__ float s_red[kFinishLanes][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int plane = blockIdx.y;
  const int col = blockIdx.x * 32 + tx;
  const int n_out = out.n[plane];                 // columns of this plane that are wanted (1 for the sum(delta) plane)
  const float* src = partial + (int64_t)plane * row_blocks * n_cols;
  float t = 0.f;
  if (col < n_out) {
    int rb = ty;
    for (; rb + 3 * kFinishLanes < row_blocks; rb += 4 * kFinishLanes) {     // four independent loads per round
      const float v0 = __ldcg(src + (int64_t)rb * n_cols + col);
      const float v1 = __ldcg(src + (int64_t)(rb + kFinishLanes) * n_cols + col);
      const float v2 = __ldcg(src + (int64_t)(rb + 2 * kFinishLanes) * n_cols + col);
      const float v3 = __ldcg(src + (int64_t)(rb + 3 * kFinishLanes) * n_cols + col);
      t += v0; t += v1; t += v2; t += v3;
    }
    for (; rb < row_blocks; rb += kFinishLanes) t += __ldcg(src + (int64_t)rb * n_cols + col);
  }
  s_red[ty][tx] = t;
  __syncthreads();
  if (ty == 0 && col < n_out) {
    float r = s_red[0][tx];
#pragma unroll
    for (int l = 1; l < kFinishLanes; ++l) r += s_red[l][tx];
    out.p[plane][col] = r;
  }
}

// ---- output DenseLayer with one unit (the logit head: wide_and_deep.py:293-297 dense_layer_5, deepfm.py:215,
// deep_and_cross.py:309) -------------------------------------------------------------------------------------
// As library GEMMs the N = 1 layer is three skinny calls (GEMV forward, GEMV weight gradient, rank-1 input
// gradient: 8-CTA kernels of 5-13 us each, ncu r1g).  Forward: one warp per row.  Backward: the pass above with
// the incoming gradient generated on the fly, g[b,k] = delta[b] * w[k], so ONE kernel emits the input gradient
// already masked by the previous layer's ReLU, that layer's BiasAddGrad, this layer's weight gradient
// gw[k] = sum_b delta[b] * h[b,k] and its bias gradient sum_b delta[b].
template <typename T>
__global__ void __launch_bounds__(256)
dense_head_fwd_kernel(const T* __restrict__ h, const T* __restrict__ w, const T* __restrict__ bias, int64_t rows,
                      int k_dim, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const T* hr = h + row * k_dim;
  float acc = 0.f;
  for (int k = lane; k < k_dim; k += 32) acc = fmaf(to_f(hr[k]), to_f(w[k]), acc);
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc + to_f(bias[0]);
}

template <>
__global__ void __launch_bounds__(256)
dense_head_fwd_kernel<__half>(const __half* __restrict__ h, const __half* __restrict__ w,
                              const __half* __restrict__ bias, int64_t rows, int k_dim, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const __half* hr = h + row * k_dim;
  float acc = 0.f;
  if ((k_dim & 7) == 0 && (reinterpret_cast<uintptr_t>(h) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0) {
    for (int c = lane; c < (k_dim >> 3); c += 32) {
      const uint4 hv = reinterpret_cast<const uint4*>(hr)[c];
      const uint4 wv = reinterpret_cast<const uint4*>(w)[c];
      const __half2* h2 = reinterpret_cast<const __half2*>(&hv);
      const __half2* w2 = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 a = __half22float2(h2[q]), b = __half22float2(w2[q]);
        acc = fmaf(a.x, b.x, acc);
        acc = fmaf(a.y, b.y, acc);
      }
    }
  } else {
    for (int k = lane; k < k_dim; k += 32) acc = fmaf(__half2float(hr[k]), __half2float(w[k]), acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc + __half2float(bias[0]);
}

// Same grid / reduction scheme as relu_bwd_bias_kernel (16-byte column chunks when VEC, else one column per
// thread).  partial holds three planes per row block: [gb_prev | gw | sum(delta) in column 0].
template <typename T, bool VEC, bool MASK>
__global__ void __launch_bounds__(kDenseThreads)
dense_head_bwd_kernel(const T* __restrict__ delta, const T* __restrict__ h, const T* __restrict__ w, T* __restrict__ gh,
                      int64_t rows, int k_dim, int rows_per_cta, float* __restrict__ partial) {
  constexpr int V = VEC ? DChunk<T>::kVec : 1;
  __shared__ float s_red[2][kDenseThreads * V];
  __shared__ float s_d[kDenseThreads];
  const int tx = threadIdx.x, ty = threadIdx.y, ntx = blockDim.x, nty = blockDim.y;
  const int chunks = k_dim / V;
  const int ch = blockIdx.y * ntx + tx;
  const bool active = ch < chunks;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(rows, r0 + rows_per_cta);
  float wv[V], a_gb[V], a_gw[V];
  float a_d = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    wv[k] = active ? to_f(w[ch * V + k]) : 0.f;
    a_gb[k] = 0.f;
    a_gw[k] = 0.f;
  }
  if (active) {
    constexpr int U = 4;
    for (int64_t r = r0 + ty; r < r1; r += (int64_t)U * nty) {
      float d[U];
      alignas(16) T hv[U][V];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t rr = min(r + (int64_t)u * nty, r1 - 1);
        d[u] = to_f(delta[rr]);
        if constexpr (VEC) {
          using Raw = typename DChunk<T>::Raw;
          *reinterpret_cast<Raw*>(hv[u]) = reinterpret_cast<const Raw*>(h)[rr * chunks + ch];
        } else {
          hv[u][0] = h[rr * k_dim + ch];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t rr = r + (int64_t)u * nty;
        if (rr < r1) {
          alignas(16) T gq[V];
#pragma unroll
          for (int k = 0; k < V; ++k) {
            const float hf = to_f(hv[u][k]);
            from_f(gq[k], d[u] * wv[k]);                   // the rank-1 GEMM's rounding (one product per element)
            if (MASK && !(hf > 0.f)) from_f(gq[k], 0.f);
            a_gb[k] += to_f(gq[k]);
            a_gw[k] = fmaf(d[u], hf, a_gw[k]);
          }
          if constexpr (VEC) {
            using Raw = typename DChunk<T>::Raw;
            reinterpret_cast<Raw*>(gh)[rr * chunks + ch] = *reinterpret_cast<const Raw*>(gq);
          } else {
            gh[rr * k_dim + ch] = gq[0];
          }
          a_d += d[u];
        }
      }
    }
  }
  const int lin = ty * ntx + tx;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    s_red[0][lin * V + k] = a_gb[k];
    s_red[1][lin * V + k] = a_gw[k];
  }
  s_d[lin] = a_d;
  __syncthreads();
  const int64_t plane = (int64_t)gridDim.x * k_dim;
  if (ty == 0 && active) {
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int k = 0; k < V; ++k) {
        float t = s_red[q][tx * V + k];
        for (int l = 1; l < nty; ++l) t += s_red[q][(l * ntx + tx) * V + k];
        partial[q * plane + (int64_t)blockIdx.x * k_dim + ch * V + k] = t;
      }
    if (ch == 0) {
      float t = s_d[0];
      for (int l = 1; l < nty; ++l) t += s_d[l * ntx];
      partial[2 * plane + (int64_t)blockIdx.x * k_dim] = t;   // plane 2, column 0
    }
  }
}

struct DensePlan {
  int ntx, nty, col_tiles, row_blocks, rows_per_cta;
  bool vec;
};

static DensePlan dense_plan(int64_t rows, int n_cols, int elem_bytes, bool aligned) {
  DensePlan p;
  const int v = 16 / elem_bytes;
  p.vec = aligned && (n_cols % v == 0);
  const int chunks = p.vec ? n_cols / v : n_cols;
  int ntx = 1;
  while (ntx < chunks && ntx < 16) ntx <<= 1;
  p.ntx = ntx;
  p.nty = kDenseThreads / ntx;
  p.col_tiles = (int)cdiv(chunks, ntx);
  // ~8 CTAs per SM (r1i: 128-296 CTAs ran at half the memory rate), but at least 4 rows per lane
  static const int per_sm = [] { const char* e = getenv("MREC_DENSE_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
  int rb = (int)cdiv(per_sm > 0 ? per_sm * kNumSMs : kDenseTargetCtas, p.col_tiles);
  if (rb > kDenseMaxRowBlocks) rb = kDenseMaxRowBlocks;
  const int64_t by_rows = cdiv(rows, (int64_t)p.nty * 4);
  if (rb > by_rows) rb = (int)by_rows;
  if (rb < 1) rb = 1;
  p.rows_per_cta = (int)cdiv(rows, rb);
  p.row_blocks = (int)cdiv(rows, p.rows_per_cta);
  return p;
}

}  // namespace mrec

using namespace mrec;

// Plain-C helper: workspace for mrec_relu_bwd_bias (per-row-block partial column sums; no initialisation needed).
MREC_API size_t mrec_relu_bwd_bias_workspace_bytes(int64_t n_cols) {
  return (size_t)kDenseMaxRowBlocks * (size_t)(n_cols > 0 ? n_cols : 1) * sizeof(float);
}

// in : g[B,N] f16|f32, y[B,N] same dtype | numel 0 (no mask: plain BiasAddGrad)
// out: gz[B,N] same dtype (may be the same buffer as g; numel 0 with no mask), gb[N] f32, workspace uint8
MREC_API int mrec_relu_bwd_bias(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                                void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  if (a.nparam != 5) return fail(ERR_NPARAM, "mrec_relu_bwd_bias: expected 5 params, got %d", a.nparam);
  const bool half = a.is(0, "float16");
  MREC_REQUIRE(half || a.is_f32(0), ERR_DTYPE, "mrec_relu_bwd_bias: g must be float16 or float32");
  MREC_REQUIRE(a.ndims[0] == 2, ERR_SHAPE, "mrec_relu_bwd_bias: g must be [B, N]");
  const int64_t rows = a.dim(0, 0), n_cols = a.dim(0, 1);
  const bool mask = a.numel(1) > 0;
  MREC_REQUIRE(!mask || (a.numel(1) == rows * n_cols && a.is(1, half ? "float16" : "float32")), ERR_SHAPE,
               "mrec_relu_bwd_bias: y must match g (or be empty)");
  MREC_REQUIRE(a.numel(2) == 0 || (a.numel(2) == rows * n_cols && a.is(2, half ? "float16" : "float32")), ERR_SHAPE,
               "mrec_relu_bwd_bias: gz must match g (or be empty)");
  MREC_REQUIRE(!mask || a.numel(2) > 0, ERR_SHAPE, "mrec_relu_bwd_bias: gz is required with a mask");
  MREC_REQUIRE(a.is_f32(3) && a.numel(3) == n_cols, ERR_SHAPE, "mrec_relu_bwd_bias: gb must be float32[N]");
  MREC_REQUIRE(n_cols < ((int64_t)1 << 31), ERR_SHAPE, "mrec_relu_bwd_bias: N too large");
  const size_t need = mrec_relu_bwd_bias_workspace_bytes(n_cols);
  if ((size_t)a.numel(4) < need)
    return fail(ERR_WORKSPACE, "mrec_relu_bwd_bias: workspace %lld bytes < required %zu", (long long)a.numel(4), need);
  if (n_cols == 0) return OK;
  if (rows == 0) {
    cudaMemsetAsync(a.params[3], 0, (size_t)n_cols * sizeof(float), a.stream);
    return OK;
  }
  for (int i : {0, 3, 4})
    if (!a.params[i]) return fail(ERR_NULL, "mrec_relu_bwd_bias: param %d is null", i);
  const int eb = half ? 2 : 4;
  bool aligned = reinterpret_cast<uintptr_t>(a.params[0]) % 16 == 0 && (n_cols * eb) % 16 == 0;
  if (mask) aligned = aligned && reinterpret_cast<uintptr_t>(a.params[1]) % 16 == 0 &&
                      reinterpret_cast<uintptr_t>(a.params[2]) % 16 == 0;
  const DensePlan p = dense_plan(rows, (int)n_cols, eb, aligned);
  MREC_REQUIRE(p.col_tiles <= 4096, ERR_SHAPE, "mrec_relu_bwd_bias: too many column tiles");
  float* partial = reinterpret_cast<float*>(a.params[4]);
  const dim3 grid(p.row_blocks, p.col_tiles), block(p.ntx, p.nty);
  const void* y = mask ? a.params[1] : nullptr;
  void* gz = a.numel(2) ? a.params[2] : nullptr;
#define MREC_DENSE(T, V)                                                                                      \
  MREC_LAUNCH((relu_bwd_bias_kernel<T, V>), grid, block, 0, a.stream, reinterpret_cast<const T*>(a.params[0]), \
              reinterpret_cast<const T*>(y), reinterpret_cast<T*>(gz), rows, (int)n_cols, p.rows_per_cta,     \
              partial)
  if (half) { if (p.vec) MREC_DENSE(__half, true); else MREC_DENSE(__half, false); }
  else      { if (p.vec) MREC_DENSE(float, true); else MREC_DENSE(float, false); }
#undef MREC_DENSE
  ColsumOut co{{a.ptr<float>(3), nullptr, nullptr}, {(int)n_cols, 0, 0}};
  MREC_LAUNCH(colsum_finish_kernel, dim3((unsigned)cdiv(n_cols, 32), 1), dim3(32, kFinishLanes), 0, a.stream, partial, p.row_blocks,
              (int)n_cols, co);
  return check_launch("relu_bwd_bias");
}

// Plain-C helper: workspace of mrec_dense_head_bwd (three planes of partials; no initialisation needed).
MREC_API size_t mrec_dense_head_workspace_bytes(int64_t k_dim) {
  return 3 * (size_t)kDenseMaxRowBlocks * (size_t)(k_dim > 0 ? k_dim : 1) * sizeof(float);
}

// in : h[B,K] f16|f32, w[K]|[K,1] same dtype, bias[1] same dtype           out: out[B]|[B,1] f32
MREC_API int mrec_dense_head_fwd(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                                 void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  if (a.nparam != 4) return fail(ERR_NPARAM, "mrec_dense_head_fwd: expected 4 params, got %d", a.nparam);
  const bool half = a.is(0, "float16");
  MREC_REQUIRE((half || a.is_f32(0)) && a.is(1, half ? "float16" : "float32") && a.is(2, half ? "float16" : "float32") &&
                   a.is_f32(3), ERR_DTYPE, "mrec_dense_head_fwd: h/w/bias f16|f32 (same), out f32");
  MREC_REQUIRE(a.ndims[0] == 2, ERR_SHAPE, "mrec_dense_head_fwd: h must be [B, K]");
  const int64_t rows = a.dim(0, 0), k = a.dim(0, 1);
  MREC_REQUIRE(a.numel(1) == k && a.numel(2) >= 1 && a.numel(3) == rows && k < ((int64_t)1 << 31), ERR_SHAPE,
               "mrec_dense_head_fwd: shape mismatch");
  if (rows == 0) return OK;
  for (int i = 0; i < 4; ++i)
    if (!a.params[i]) return fail(ERR_NULL, "mrec_dense_head_fwd: param %d is null", i);
  const int grid = (int)cdiv(rows * 32, 256);
  if (half)
    MREC_LAUNCH(dense_head_fwd_kernel<__half>, grid, 256, 0, a.stream, a.ptr<__half>(0), a.ptr<__half>(1),
                a.ptr<__half>(2), rows, (int)k, a.ptr<float>(3));
  else
    MREC_LAUNCH(dense_head_fwd_kernel<float>, grid, 256, 0, a.stream, a.ptr<float>(0), a.ptr<float>(1),
                a.ptr<float>(2), rows, (int)k, a.ptr<float>(3));
  return check_launch("dense_head_fwd");
}

// in : delta[B]|[B,1] f16|f32, h[B,K] same dtype (the head's input = previous layer's output), w[K]|[K,1] same dtype,
//      relu_like[0|1] (numel 1: mask the input gradient with h > 0, i.e. the previous layer ends in ReLU)
// out: gh[B,K] same dtype, gw[K]|[K,1] f32, gb_head[1] f32, gb_prev[K] f32 | numel 0, workspace uint8
MREC_API int mrec_dense_head_bwd(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                                 void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  if (a.nparam != 9) return fail(ERR_NPARAM, "mrec_dense_head_bwd: expected 9 params, got %d", a.nparam);
  const bool half = a.is(0, "float16");
  const char* dt = half ? "float16" : "float32";
  MREC_REQUIRE((half || a.is_f32(0)) && a.is(1, dt) && a.is(2, dt) && a.is(4, dt) && a.is_f32(5) && a.is_f32(6) &&
                   a.is_f32(7), ERR_DTYPE, "mrec_dense_head_bwd: delta/h/w/gh f16|f32 (same), gw/gb f32");
  MREC_REQUIRE(a.ndims[1] == 2, ERR_SHAPE, "mrec_dense_head_bwd: h must be [B, K]");
  const int64_t rows = a.dim(1, 0), k = a.dim(1, 1);
  MREC_REQUIRE(a.numel(0) == rows && a.numel(2) == k && a.numel(4) == rows * k && a.numel(5) == k && a.numel(6) >= 1 &&
                   (a.numel(7) == 0 || a.numel(7) == k) && k < ((int64_t)1 << 31), ERR_SHAPE,
               "mrec_dense_head_bwd: shape mismatch");
  const size_t need = mrec_dense_head_workspace_bytes(k);
  if ((size_t)a.numel(8) < need)
    return fail(ERR_WORKSPACE, "mrec_dense_head_bwd: workspace %lld bytes < required %zu", (long long)a.numel(8), need);
  if (k == 0) return OK;
  if (rows == 0) {
    cudaMemsetAsync(a.params[5], 0, (size_t)k * sizeof(float), a.stream);
    cudaMemsetAsync(a.params[6], 0, sizeof(float), a.stream);
    if (a.numel(7)) cudaMemsetAsync(a.params[7], 0, (size_t)k * sizeof(float), a.stream);
    return OK;
  }
  for (int i : {0, 1, 2, 4, 5, 6, 8})
    if (!a.params[i]) return fail(ERR_NULL, "mrec_dense_head_bwd: param %d is null", i);
  const int eb = half ? 2 : 4;
  const bool aligned = reinterpret_cast<uintptr_t>(a.params[1]) % 16 == 0 && reinterpret_cast<uintptr_t>(a.params[4]) % 16 == 0 &&
                       (k * eb) % 16 == 0;
  const DensePlan p = dense_plan(rows, (int)k, eb, aligned);
  MREC_REQUIRE(p.col_tiles <= 4096, ERR_SHAPE, "mrec_dense_head_bwd: K too large");
  float* partial = reinterpret_cast<float*>(a.params[8]);
  const dim3 grid(p.row_blocks, p.col_tiles), block(p.ntx, p.nty);
  const bool mask = a.numel(3) > 0;
  float* gb_prev = a.numel(7) ? a.ptr<float>(7) : nullptr;
#define MREC_HEAD(T, V, M)                                                                                       \
  MREC_LAUNCH((dense_head_bwd_kernel<T, V, M>), grid, block, 0, a.stream, a.ptr<T>(0), a.ptr<T>(1), a.ptr<T>(2),  \
              a.ptr<T>(4), rows, (int)k, p.rows_per_cta, partial)
  if (half) {
    if (p.vec) { if (mask) MREC_HEAD(__half, true, true); else MREC_HEAD(__half, true, false); }
    else       { if (mask) MREC_HEAD(__half, false, true); else MREC_HEAD(__half, false, false); }
  } else {
    if (p.vec) { if (mask) MREC_HEAD(float, true, true); else MREC_HEAD(float, true, false); }
    else       { if (mask) MREC_HEAD(float, false, true); else MREC_HEAD(float, false, false); }
  }
#undef MREC_HEAD
  // planes: 0 = previous layer's BiasAddGrad (optional), 1 = gw, 2 = sum(delta) in column 0
  ColsumOut co{{gb_prev, a.ptr<float>(5), a.ptr<float>(6)}, {gb_prev ? (int)k : 0, (int)k, 1}};
  MREC_LAUNCH(colsum_finish_kernel, dim3((unsigned)cdiv(k, 32), 3), dim3(32, kFinishLanes), 0, a.stream, partial, p.row_blocks, (int)k, co);
  return check_launch("dense_head_bwd");
}
