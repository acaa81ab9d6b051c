// K8 — DCN cross-layer stack, forward and backward, fused over all layers.
//
// Replaces, per layer, tensor_dot + batched MatMul([B,D',1] x [B,1,1]) + two adds
// (models/deep_and_cross/src/deep_and_cross.py:139-149) applied six times (:301-306) and its autodiff.
//
// The layer  x_{l+1} = x_0 * (x_l . w_l) + b_l + x_l  keeps every x_l in the plane spanned by x_0 and a
// row-independent vector:  x_l = c_l * x_0 + Bcum_l  with  Bcum_l = sum_{k<l} b_k,  c_0 = 1,
//     s_l = x_l . w_l = c_l * p_l + q_l,   p_l = x_0 . w_l,   q_l = Bcum_l . w_l,   c_{l+1} = c_l + s_l.
// So one pass over a row computes the L dots p_l at once (one block reduction of L values instead of L
// dependent ones), a scalar recurrence, and y = c_L * x_0 + Bcum_L: x_0 is read once, y written once
// (2*B*D'*4 bytes, the fused-stack roofline of SURVEY 8d).  Backward, with r = dy . x_0:
//     ds_l = r + sum_{k>l} ds_k p_k,    dx = c_L * dy + sum_k (ds_k c_k) w_k,
//     dw_l = sum_rows (ds_l c_l) x_0 + (sum_rows ds_l) Bcum_l,   db_l = sum_rows dy + sum_{k>l} (sum_rows ds_k) w_k
// i.e. one block reduction per row, x_0 and dy read once, dx written once (3*B*D'*4 bytes); the column
// accumulators for dw/db live in registers per CTA and are combined over CTAs in a fixed order by a
// second kernel (deterministic, no atomics).  w/Bcum (L*D'*4 = 75 KB at D' = 3120) are read through L1
// (plain cached loads) while x_0 / dy / y / dx bypass it (no_allocate / streaming stores).
#include "common.cuh"
#include <stdlib.h>

namespace mrec {

constexpr int kCrossMaxL = 8;

template <typename Vec> struct CrV;
template <> struct CrV<float4> {
  static constexpr int width = 4;
  static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ float4 ld_stream(const float* p, int64_t i) {
    return ld_stream_f4(reinterpret_cast<const float4*>(p) + i);
  }
  static __device__ __forceinline__ float4 ld(const float* p, int64_t i) {
    return reinterpret_cast<const float4*>(p)[i];
  }
  static __device__ __forceinline__ void st_stream(float* p, int64_t i, const float4& v) {
    st_stream_f4(reinterpret_cast<float4*>(p) + i, v);
  }
  static __device__ __forceinline__ void st(float* p, int64_t i, const float4& v) {
    reinterpret_cast<float4*>(p)[i] = v;
  }
  static __device__ __forceinline__ float dot(const float4& a, const float4& b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
  }
  static __device__ __forceinline__ void axpy(float4& y, float a, const float4& x) { f4_fma(y, x, a); }
  static __device__ __forceinline__ float4 scale_add(const float4& x, float a, const float4& b) {
    return make_float4(fmaf(x.x, a, b.x), fmaf(x.y, a, b.y), fmaf(x.z, a, b.z), fmaf(x.w, a, b.w));
  }
  static __device__ __forceinline__ void add(float4& y, const float4& x) {
    y.x += x.x; y.y += x.y; y.z += x.z; y.w += x.w;
  }
};
template <> struct CrV<float> {
  static constexpr int width = 1;
  static __device__ __forceinline__ float zero() { return 0.f; }
  static __device__ __forceinline__ float ld_stream(const float* p, int64_t i) { return ld_stream_f1(p + i); }
  static __device__ __forceinline__ float ld(const float* p, int64_t i) { return p[i]; }
  static __device__ __forceinline__ void st_stream(float* p, int64_t i, const float& v) { p[i] = v; }
  static __device__ __forceinline__ void st(float* p, int64_t i, const float& v) { p[i] = v; }
  static __device__ __forceinline__ float dot(const float& a, const float& b) { return a * b; }
  static __device__ __forceinline__ void axpy(float& y, float a, const float& x) { y = fmaf(x, a, y); }
  static __device__ __forceinline__ float scale_add(const float& x, float a, const float& b) { return fmaf(x, a, b); }
  static __device__ __forceinline__ void add(float& y, const float& x) { y += x; }
};

// bcum[l][d] = sum_{k<l} b[k][d] (l = 0..L), q[l] = <bcum[l], w[l]>.  One block; fixed reduction order.
__global__ void __launch_bounds__(1024)
cross_prep_kernel(const float* __restrict__ w, const float* __restrict__ b, int layers, int dp,
                  float* __restrict__ bcum, float* __restrict__ q) {
  __shared__ float s_red[32][kCrossMaxL];
  float acc[kCrossMaxL];
#pragma unroll
  for (int l = 0; l < kCrossMaxL; ++l) acc[l] = 0.f;
  for (int d = threadIdx.x; d < dp; d += blockDim.x) {
    float run = 0.f;
#pragma unroll
    for (int l = 0; l < kCrossMaxL; ++l) {
      if (l < layers) {
        bcum[(int64_t)l * dp + d] = run;
        acc[l] = fmaf(run, w[(int64_t)l * dp + d], acc[l]);
        run += b[(int64_t)l * dp + d];
      }
    }
    bcum[(int64_t)layers * dp + d] = run;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int l = 0; l < kCrossMaxL; ++l) {
    const float v = warp_sum(acc[l]);
    if (lane == 0) s_red[warp][l] = v;
  }
  __syncthreads();
  if (threadIdx.x < layers) {
    float t = 0.f;
    for (int wp = 0; wp < (int)(blockDim.x >> 5); ++wp) t += s_red[wp][threadIdx.x];
    q[threadIdx.x] = t;
  }
}

// Block-wide sum of NV values; every thread returns all totals.  Stage 1: warp shuffles, one smem row
// per warp.  Stage 2: in every warp lane i < NV adds column i over the warps (fixed order) and the totals
// are broadcast by shuffle, so there is one __syncthreads per call; callers alternate two smem slabs by
// iteration parity.
template <int NV, int WARPS>
__device__ __forceinline__ void block_sum(float (&v)[NV], float (*buf)[NV]) {
  static_assert(NV <= 32, "block_sum handles at most 32 values");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float t = warp_sum(v[i]);
    if (lane == 0) buf[warp][i] = t;
  }
  __syncthreads();
  float col = 0.f;
  if (lane < NV) {
#pragma unroll
    for (int wp = 0; wp < WARPS; ++wp) col += buf[wp][lane];
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = __shfl_sync(0xffffffffu, col, i);
}

// Software pipeline shared by both kernels: a CTA owns rows row0, row0 + stride, ...; the loads of iteration
// i+1 are issued (into a second register set) before iteration i's reduction, so HBM latency overlaps the
// reduction, the FMA phase and the stores of the current rows.  R rows per iteration, SLOTS chunks per
// thread, THREADS threads per CTA (1024 x 1 chunk for D' <= 4096 keeps the dw/db accumulators of the backward
// at 7 float4 per thread, which leaves room for the second register set).
template <typename Vec, int SLOTS, int R, int THREADS>
struct RowTile {
  Vec v[R][SLOTS];
  __device__ __forceinline__ void load(const float* __restrict__ src, int64_t row0, int64_t batch, int dpv) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int k = 0; k < SLOTS; ++k) {
        const int i = threadIdx.x + k * THREADS;
        // clamp instead of branching: out-of-range chunks / rows re-read a valid element and are masked later
        const int ic = min(i, dpv - 1);
        const int64_t rc = min(row0 + r, batch - 1);
        v[r][k] = CrV<Vec>::ld_stream(src, rc * dpv + ic);
      }
  }
  __device__ __forceinline__ void fence() {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int k = 0; k < SLOTS; ++k) reg_fence(v[r][k]);
  }
};

template <typename Vec, int SLOTS, int L, int R, int THREADS, int MINB, bool PF>
__global__ void __launch_bounds__(THREADS, MINB)
cross_fwd_kernel(const float* __restrict__ x0, const float* __restrict__ w, const float* __restrict__ q,
                 const float* __restrict__ bcum, int64_t batch, int dpv /* D' / vec width */,
                 float* __restrict__ y, float* __restrict__ p_out) {
  constexpr int WARPS = THREADS / 32;
  __shared__ float s_red[2][WARPS][R * L];
  const int tid = threadIdx.x;
  float ql[L];
#pragma unroll
  for (int l = 0; l < L; ++l) ql[l] = q[l];
  const int64_t stride = (int64_t)gridDim.x * R;
  RowTile<Vec, SLOTS, R, THREADS> cur, nxt;
  int64_t row0 = (int64_t)blockIdx.x * R;
  if (row0 < batch) cur.load(x0, row0, batch, dpv);
  int par = 0;
  for (; row0 < batch; row0 += stride, par ^= 1) {
    const bool more = PF && (row0 + stride < batch);
    if (!PF && par >= 0 && row0 != (int64_t)blockIdx.x * R) cur.load(x0, row0, batch, dpv);
    if (more) nxt.load(x0, row0 + stride, batch, dpv);   // in flight during everything below
    cur.fence();
    float p[R * L];
#pragma unroll
    for (int i = 0; i < R * L; ++i) p[i] = 0.f;
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
      const int i = tid + k * THREADS;
      if (i < dpv) {
#pragma unroll
        for (int l = 0; l < L; ++l) {
          const Vec wv = CrV<Vec>::ld(w, (int64_t)l * dpv + i);
#pragma unroll
          for (int r = 0; r < R; ++r) p[r * L + l] += CrV<Vec>::dot(cur.v[r][k], wv);
        }
      }
    }
    block_sum<R * L, WARPS>(p, s_red[par]);
    float c[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      c[r] = 1.f;
#pragma unroll
      for (int l = 0; l < L; ++l) c[r] += fmaf(c[r], p[r * L + l], ql[l]);
    }
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
      const int i = tid + k * THREADS;
      if (i < dpv) {
        const Vec bl = CrV<Vec>::ld(bcum, (int64_t)L * dpv + i);
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (row0 + r < batch) CrV<Vec>::st_stream(y, (row0 + r) * dpv + i, CrV<Vec>::scale_add(cur.v[r][k], c[r], bl));
      }
    }
#pragma unroll
    for (int i = 0; i < R * L; ++i)
      if (tid == i && row0 + i / L < batch) p_out[(row0 + i / L) * L + (i % L)] = p[i];
    if (more) cur = nxt;
  }
}

template <typename Vec, int SLOTS, int L, int R, int THREADS, bool PF>
__global__ void __launch_bounds__(THREADS, 1)
cross_bwd_kernel(const float* __restrict__ x0, const float* __restrict__ dy, const float* __restrict__ w,
                 const float* __restrict__ q, const float* __restrict__ p_in, int64_t batch, int dpv,
                 float* __restrict__ dx, float* __restrict__ part /* [grid][L+1][D'] */,
                 float* __restrict__ sd_part /* [grid][L] */) {
  constexpr int WARPS = THREADS / 32;
  __shared__ float s_red[2][WARPS][R];
  const int tid = threadIdx.x;
  float ql[L];
#pragma unroll
  for (int l = 0; l < L; ++l) ql[l] = q[l];
  Vec A[L][SLOTS], Y[SLOTS];
  float sd[L];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    sd[l] = 0.f;
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) A[l][k] = CrV<Vec>::zero();
  }
#pragma unroll
  for (int k = 0; k < SLOTS; ++k) Y[k] = CrV<Vec>::zero();

  const int64_t stride = (int64_t)gridDim.x * R;
  RowTile<Vec, SLOTS, R, THREADS> x, g, xn, gn;
  int64_t row0 = (int64_t)blockIdx.x * R;
  if (row0 < batch) {
    x.load(x0, row0, batch, dpv);
    g.load(dy, row0, batch, dpv);
  }
  int par = 0;
  for (; row0 < batch; row0 += stride, par ^= 1) {
    const bool more = PF && (row0 + stride < batch);
    if (!PF && row0 != (int64_t)blockIdx.x * R) {
      x.load(x0, row0, batch, dpv);
      g.load(dy, row0, batch, dpv);
    }
    if (more) {
      xn.load(x0, row0 + stride, batch, dpv);
      gn.load(dy, row0 + stride, batch, dpv);
    }
    x.fence();
    g.fence();
    float rr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      rr[r] = 0.f;
#pragma unroll
      for (int k = 0; k < SLOTS; ++k)
        if (tid + k * THREADS < dpv) rr[r] += CrV<Vec>::dot(x.v[r][k], g.v[r][k]);
    }
    block_sum<R, WARPS>(rr, s_red[par]);
    // per row: forward scalars c_l, backward scalars ds_l -> coef_l = ds_l * c_l
    float coef[R][L], cl[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool live = row0 + r < batch;
      float p[L], c[L + 1];
      c[0] = 1.f;
#pragma unroll
      for (int l = 0; l < L; ++l) {
        p[l] = live ? p_in[(row0 + r) * L + l] : 0.f;
        c[l + 1] = c[l] + fmaf(c[l], p[l], ql[l]);
      }
      cl[r] = live ? c[L] : 0.f;
      float t = 0.f;
#pragma unroll
      for (int l = L - 1; l >= 0; --l) {
        const float ds = live ? rr[r] + t : 0.f;
        t = fmaf(ds, p[l], t);
        sd[l] += ds;
        coef[r][l] = ds * c[l];
      }
    }
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
      const int i = tid + k * THREADS;
      if (i < dpv) {
        Vec o[R];
#pragma unroll
        for (int r = 0; r < R; ++r) o[r] = CrV<Vec>::scale_add(g.v[r][k], cl[r], CrV<Vec>::zero());
#pragma unroll
        for (int l = 0; l < L; ++l) {
          const Vec wv = CrV<Vec>::ld(w, (int64_t)l * dpv + i);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            CrV<Vec>::axpy(o[r], coef[r][l], wv);
            CrV<Vec>::axpy(A[l][k], coef[r][l], x.v[r][k]);   // coef is 0 for rows past the end
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (row0 + r < batch) {
            CrV<Vec>::add(Y[k], g.v[r][k]);
            CrV<Vec>::st_stream(dx, (row0 + r) * dpv + i, o[r]);
          }
        }
      }
    }
    if (more) {
      x = xn;
      g = gn;
    }
  }
  float* mine = part + (int64_t)blockIdx.x * (L + 1) * dpv * CrV<Vec>::width;
#pragma unroll
  for (int k = 0; k < SLOTS; ++k) {
    const int i = tid + k * THREADS;
    if (i < dpv) {
#pragma unroll
      for (int l = 0; l < L; ++l) CrV<Vec>::st(mine, (int64_t)l * dpv + i, A[l][k]);
      CrV<Vec>::st(mine, (int64_t)L * dpv + i, Y[k]);
    }
  }
  if (tid == 0) {
#pragma unroll
    for (int l = 0; l < L; ++l) sd_part[blockIdx.x * L + l] = sd[l];
  }
}

// Stage 1 of the finish: column sums of the per-CTA partials in a fixed order.  grid = (ceil(D'/32), L+2):
// blockIdx.y <= L sums array y (A_0..A_{L-1}, Ysum) over the CTAs, blockIdx.y == L+1 sums the ds scalars.
// Block (32 columns x 8 part-groups): group q adds parts q, q+8, ...; the 8 group partials are added in order.
__global__ void __launch_bounds__(256)
cross_bwd_colsum_kernel(const float* __restrict__ part, const float* __restrict__ sd_part, int nparts,
                        int layers, int dp, float* __restrict__ sums /* [L+1][D'] */, float* __restrict__ sdt) {
  __shared__ float s_acc[8][32];
  const int tx = threadIdx.x & 31, q = threadIdx.x >> 5;
  if ((int)blockIdx.y == layers + 1) {
    if (blockIdx.x == 0 && threadIdx.x < layers) {
      float t = 0.f;
      for (int pth = 0; pth < nparts; ++pth) t += sd_part[pth * layers + threadIdx.x];
      sdt[threadIdx.x] = t;
    }
    return;
  }
  const int d = blockIdx.x * 32 + tx;
  const int arr = blockIdx.y;
  float acc = 0.f;
  if (d < dp)
    for (int pth = q; pth < nparts; pth += 8) acc += part[((int64_t)pth * (layers + 1) + arr) * dp + d];
  s_acc[q][tx] = acc;
  __syncthreads();
  if (q == 0 && d < dp) {
    float t = s_acc[0][tx];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += s_acc[k][tx];
    sums[(int64_t)arr * dp + d] = t;
  }
}

// Stage 2: dw_l = A_l + (sum ds_l) Bcum_l ;  db_l = Ysum + sum_{k>l} (sum ds_k) w_k
__global__ void __launch_bounds__(256)
cross_bwd_finish_kernel(const float* __restrict__ sums, const float* __restrict__ sdt,
                        const float* __restrict__ w, const float* __restrict__ bcum, int layers, int dp,
                        float* __restrict__ dw, float* __restrict__ db) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= dp) return;
  const float ysum = sums[(int64_t)layers * dp + d];
  float tail = 0.f;  // sum_{k>l} sdt[k] * w[k][d], built from the last layer down
  for (int l = layers - 1; l >= 0; --l) {
    const float sl = sdt[l];
    dw[(int64_t)l * dp + d] = fmaf(sl, bcum[(int64_t)l * dp + d], sums[(int64_t)l * dp + d]);
    db[(int64_t)l * dp + d] = ysum + tail;
    tail = fmaf(sl, w[(int64_t)l * dp + d], tail);
  }
}

struct CrossWs {
  size_t off_bcum, off_q, off_part, off_sd, off_sums, off_sdt, total;
  int nparts;
};
static CrossWs cross_ws(int layers, int dp, bool backward) {
  CrossWs W;
  size_t o = 0;
  W.off_bcum = o; o = align_up(o + (size_t)(layers + 1) * dp * 4, 256);
  W.off_q = o; o = align_up(o + (size_t)kCrossMaxL * 4, 256);
  W.nparts = kNumSMs;
  W.off_part = o;
  if (backward) o = align_up(o + (size_t)W.nparts * (layers + 1) * dp * 4, 256);
  W.off_sd = o;
  if (backward) o = align_up(o + (size_t)W.nparts * layers * 4, 256);
  W.off_sums = o;
  if (backward) o = align_up(o + (size_t)(layers + 1) * dp * 4, 256);
  W.off_sdt = o;
  if (backward) o = align_up(o + (size_t)kCrossMaxL * 4, 256);
  W.total = o;
  return W;
}

template <typename Vec, int SLOTS, int THREADS, int R, int MINB, bool PF>
static int launch_fwd(int layers, int grid, cudaStream_t st, const float* x0, const float* w, const float* q,
                      const float* bcum, int64_t batch, int dpv, float* y, float* p) {
  switch (layers) {
#define C(LL) case LL: MREC_LAUNCH((cross_fwd_kernel<Vec, SLOTS, LL, R, THREADS, MINB, PF>), grid, THREADS, 0, st, x0, w, q, bcum, batch, dpv, y, p); break;
    C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8)
#undef C
    default: return fail(ERR_DIM, "mrec_cross: 1 <= layers <= %d", kCrossMaxL);
  }
  return OK;
}
template <typename Vec, int SLOTS, int THREADS, int R, bool PF>
static int launch_bwd(int layers, int grid, cudaStream_t st, const float* x0, const float* dy, const float* w,
                      const float* q, const float* p, int64_t batch, int dpv, float* dx, float* part, float* sd) {
  switch (layers) {
#define C(LL) case LL: MREC_LAUNCH((cross_bwd_kernel<Vec, SLOTS, LL, R, THREADS, PF>), grid, THREADS, 0, st, x0, dy, w, q, p, batch, dpv, dx, part, sd); break;
    C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8)
#undef C
    default: return fail(ERR_DIM, "mrec_cross: 1 <= layers <= %d", kCrossMaxL);
  }
  return OK;
}

static int cross_shapes(const Aot& a, int xi, int wi, int bi, int64_t* batch, int* dp, int* layers) {
  MREC_REQUIRE(a.is_f32(xi) && a.is_f32(wi) && a.is_f32(bi), ERR_DTYPE, "mrec_cross: x0/w/b must be float32");
  MREC_REQUIRE(a.ndims[xi] == 2, ERR_SHAPE, "mrec_cross: x0 must be [B, D']");
  *batch = a.dim(xi, 0);
  *dp = (int)a.dim(xi, 1);
  MREC_REQUIRE(a.ndims[wi] >= 2 && a.numel(wi) % *dp == 0, ERR_SHAPE, "mrec_cross: w must be [L, D']");
  *layers = (int)(a.numel(wi) / *dp);
  MREC_REQUIRE(a.numel(bi) == a.numel(wi), ERR_SHAPE, "mrec_cross: b must have w's shape");
  MREC_REQUIRE(*layers >= 1 && *layers <= kCrossMaxL, ERR_DIM, "mrec_cross: 1 <= layers <= %d", kCrossMaxL);
  const bool v4 = (*dp % 4 == 0);
  const int dpv = v4 ? *dp / 4 : *dp;
  MREC_REQUIRE(dpv <= (v4 ? 2048 : 8192), ERR_DIM, "mrec_cross: D' = %d too large (max 8192)", *dp);
  return OK;
}

}  // namespace mrec

using namespace mrec;

MREC_API size_t mrec_cross_workspace_bytes(int64_t layers, int dp) {
  return cross_ws((int)layers, dp, true).total;
}

// in : x0[B,D'] w[L,D'] b[L,D'] f32      out: y[B,D'] f32, p[B,L] f32 (saved dots x0.w_l), workspace
MREC_API int mrec_cross_fwd(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                            void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  MREC_CHECK_NPARAM(a, 6);
  int64_t batch; int dp, layers;
  int rc = cross_shapes(a, 0, 1, 2, &batch, &dp, &layers);
  if (rc) return rc;
  MREC_REQUIRE(a.is_f32(3) && a.numel(3) == batch * dp, ERR_SHAPE, "mrec_cross_fwd: y must be [B, D'] f32");
  MREC_REQUIRE(a.is_f32(4) && a.numel(4) == batch * layers, ERR_SHAPE, "mrec_cross_fwd: p must be [B, L] f32");
  const CrossWs W = cross_ws(layers, dp, false);
  MREC_REQUIRE((size_t)a.numel(5) >= W.total, ERR_WORKSPACE, "mrec_cross_fwd: workspace %lld < %zu",
               (long long)a.numel(5), W.total);
  MREC_REQUIRE(a.aligned(0, 16) && a.aligned(1, 16) && a.aligned(3, 16) && a.aligned(5, 16), ERR_ALIGN,
               "mrec_cross_fwd: buffers must be 16-byte aligned");
  char* ws = a.ptr<char>(5);
  float* bcum = reinterpret_cast<float*>(ws + W.off_bcum);
  float* q = reinterpret_cast<float*>(ws + W.off_q);
  MREC_LAUNCH(cross_prep_kernel, 1, 1024, 0, a.stream, a.ptr<float>(1), a.ptr<float>(2), layers, dp, bcum, q);
  if (batch == 0) return check_launch("cross_prep");
  const float *x0 = a.ptr<float>(0), *w = a.ptr<float>(1);
  float *y = a.ptr<float>(3), *p = a.ptr<float>(4);
  const char* env = getenv("MREC_CROSS_CFG");
  const int cfg = env ? atoi(env) : 0;
  if (dp % 4 == 0) {
    const int dpv = dp / 4;
    // measured r1f (16384 x 3120, L = 6): 512 thr x 2 chunks, 2 rows, 2 CTAs/SM, no prefetch: 0.129 ms;
    // 256 thr x 4 chunks, 4 CTAs/SM: 0.131 ms; 1024 thr x 1 chunk with prefetch, 1 CTA/SM: 0.203 ms.
    // ncu: 56 % issue-slot utilisation, 38 % DRAM — the 12-value block reduction per row pair is the cost.
    if (dpv <= 1024 && cfg == 1) rc = launch_fwd<float4, 4, 256, 1, 4, false>(layers, grid_for(batch, 4), a.stream, x0, w, q, bcum, batch, dpv, y, p);
    else if (dpv <= 1024) rc = launch_fwd<float4, 2, 512, 2, 2, false>(layers, grid_for(cdiv(batch, 2), 2), a.stream, x0, w, q, bcum, batch, dpv, y, p);
    else rc = launch_fwd<float4, 4, 512, 1, 1, false>(layers, grid_for(batch, 1), a.stream, x0, w, q, bcum, batch, dpv, y, p);
  } else {
    if (dp <= 1024) rc = launch_fwd<float, 2, 512, 2, 2, false>(layers, grid_for(cdiv(batch, 2), 2), a.stream, x0, w, q, bcum, batch, dp, y, p);
    else if (dp <= 4096) rc = launch_fwd<float, 8, 512, 1, 1, false>(layers, grid_for(batch, 1), a.stream, x0, w, q, bcum, batch, dp, y, p);
    else rc = launch_fwd<float, 16, 512, 1, 1, false>(layers, grid_for(batch, 1), a.stream, x0, w, q, bcum, batch, dp, y, p);
  }
  if (rc) return rc;
  return check_launch("cross_fwd");
}

// in : x0[B,D'] dy[B,D'] w[L,D'] b[L,D'] p[B,L]     out: dx[B,D'] dw[L,D'] db[L,D'] f32, workspace
MREC_API int mrec_cross_bwd(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                            void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  MREC_CHECK_NPARAM(a, 9);
  int64_t batch; int dp, layers;
  int rc = cross_shapes(a, 0, 2, 3, &batch, &dp, &layers);
  if (rc) return rc;
  MREC_REQUIRE(a.is_f32(1) && a.numel(1) == batch * dp, ERR_SHAPE, "mrec_cross_bwd: dy must be [B, D'] f32");
  MREC_REQUIRE(a.is_f32(4) && a.numel(4) == batch * layers, ERR_SHAPE, "mrec_cross_bwd: p must be [B, L] f32");
  MREC_REQUIRE(a.is_f32(5) && a.numel(5) == batch * dp, ERR_SHAPE, "mrec_cross_bwd: dx must be [B, D'] f32");
  MREC_REQUIRE(a.is_f32(6) && a.is_f32(7) && a.numel(6) == (int64_t)layers * dp && a.numel(7) == a.numel(6),
               ERR_SHAPE, "mrec_cross_bwd: dw/db must be [L, D'] f32");
  const CrossWs W = cross_ws(layers, dp, true);
  MREC_REQUIRE((size_t)a.numel(8) >= W.total, ERR_WORKSPACE, "mrec_cross_bwd: workspace %lld < %zu",
               (long long)a.numel(8), W.total);
  MREC_REQUIRE(a.aligned(0, 16) && a.aligned(1, 16) && a.aligned(2, 16) && a.aligned(5, 16) && a.aligned(8, 16),
               ERR_ALIGN, "mrec_cross_bwd: buffers must be 16-byte aligned");
  char* ws = a.ptr<char>(8);
  float* bcum = reinterpret_cast<float*>(ws + W.off_bcum);
  float* q = reinterpret_cast<float*>(ws + W.off_q);
  float* part = reinterpret_cast<float*>(ws + W.off_part);
  float* sd = reinterpret_cast<float*>(ws + W.off_sd);
  MREC_LAUNCH(cross_prep_kernel, 1, 1024, 0, a.stream, a.ptr<float>(2), a.ptr<float>(3), layers, dp, bcum, q);
  const int grid = W.nparts;  // every CTA writes its (possibly zero) partial: the finish sums all of them
  const float *x0 = a.ptr<float>(0), *dy = a.ptr<float>(1), *w = a.ptr<float>(2), *p = a.ptr<float>(4);
  float* dx = a.ptr<float>(5);
  const char* env = getenv("MREC_CROSS_CFG");
  const int cfg = env ? atoi(env) : 0;
  if (dp % 4 == 0) {
    const int dpv = dp / 4;
    // measured r1f: 512 thr x 2 chunks, 1 row + prefetch of the next: 0.206 ms; 2 rows, no prefetch: 0.219 ms;
    // 1024 thr x 1 chunk: 0.29-0.31 ms (one CTA per SM: every barrier stalls the whole SM)
    if (dpv <= 1024 && cfg != 1) rc = launch_bwd<float4, 2, 512, 1, true>(layers, grid, a.stream, x0, dy, w, q, p, batch, dpv, dx, part, sd);
    else if (dpv <= 1024) rc = launch_bwd<float4, 2, 512, 2, false>(layers, grid, a.stream, x0, dy, w, q, p, batch, dpv, dx, part, sd);
    else rc = launch_bwd<float4, 4, 512, 1, false>(layers, grid, a.stream, x0, dy, w, q, p, batch, dpv, dx, part, sd);
  } else {
    if (dp <= 1024) rc = launch_bwd<float, 2, 512, 2, false>(layers, grid, a.stream, x0, dy, w, q, p, batch, dp, dx, part, sd);
    else if (dp <= 4096) rc = launch_bwd<float, 8, 512, 1, false>(layers, grid, a.stream, x0, dy, w, q, p, batch, dp, dx, part, sd);
    else rc = launch_bwd<float, 16, 512, 1, false>(layers, grid, a.stream, x0, dy, w, q, p, batch, dp, dx, part, sd);
  }
  if (rc) return rc;
  float* sums = reinterpret_cast<float*>(ws + W.off_sums);
  float* sdt = reinterpret_cast<float*>(ws + W.off_sdt);
  MREC_LAUNCH(cross_bwd_colsum_kernel, dim3((unsigned)cdiv(dp, 32), (unsigned)(layers + 2)), 256, 0, a.stream, part,
              sd, W.nparts, layers, dp, sums, sdt);
  MREC_LAUNCH(cross_bwd_finish_kernel, (int)cdiv(dp, 256), 256, 0, a.stream, sums, sdt, w, bcum, layers, dp,
              a.ptr<float>(6), a.ptr<float>(7));
  return check_launch("cross_bwd");
}
