// K1 — embedding gather (+ fused mask multiply, + fused dim-1 masked field reduce).
//
// Replaces the upstream Gather / SparseGatherV2 / EmbeddingLookup kernels reached from
//   models/wide_deep/src/wide_and_deep.py:300-309   (wide + deep lookup, mask mul, wide reduce)
//   models/deepfm/src/deepfm.py:215-222             (W / V lookup, mask mul, linear reduce)
//   models/deep_and_cross/src/deep_and_cross.py:188-203,295-298
//   mindspore_rec/ops/embedding.py:194               (gather_revert by inverse index)
//
// Design (HBM-bound, no tensor cores): the output [N, D] is viewed as a flat array of 16-byte
// chunks; consecutive threads own consecutive chunks, so all 32 lanes are busy for any D % 4 == 0
// (D = 80 -> 20 chunks per row would leave 12 lanes idle with a warp-per-row mapping) and stores are
// perfectly coalesced.  The kernel is persistent (grid = k x 148 CTAs); each CTA walks row tiles and
// stages the tile's ids (+ mask) into shared memory with a double-buffered 1-D bulk async copy
// (TMA engine, cp.async.bulk + mbarrier) so the dependent row loads of tile t never wait on the id
// load of tile t: the id fetch for tile t+1 is in flight while tile t's rows stream.  Every thread
// issues all of its row-chunk loads (ld.global.nc.L1::no_allocate.v4) before the first store.
//
// Out-of-range ids follow the upstream GPU Gather: the output row is zeros; in addition a
// device flag (optional trailing output) is raised so a debug caller can mirror the CPU kernel's error.
#include "common.cuh"
#include <cuda_fp16.h>

namespace mrec {

constexpr int kGatherThreads = 256;

// One 16-byte table chunk becomes 4 outputs: fp32 (16 B) or fp16 (8 B; the Cast(x, float16) at the head
// of the reference's DenseLayer, wide_and_deep.py:119-120, fused into the gather store).
template <typename OutT> struct OutChunk;
template <> struct OutChunk<float> {
  using type = float4;
  static __device__ __forceinline__ void store(float4* p, const float4& v) { st_stream_f4(p, v); }
};
template <> struct OutChunk<__half> {
  using type = uint2;
  static __device__ __forceinline__ void store(uint2* p, const float4& v) {
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&lo);
    u.y = *reinterpret_cast<const uint32_t*>(&hi);
    asm volatile("st.global.cs.v2.b32 [%0], {%1,%2};" ::"l"(p), "r"(u.x), "r"(u.y) : "memory");
  }
};

__host__ __device__ constexpr int tile_rows_for(int cpr) {
  // aim for >= 1024 chunks per tile, tile rows a multiple of 64 (bulk-copy alignment)
  return cpr >= 16 ? 64 : (cpr >= 8 ? 128 : 256);
}

template <int CPR, typename IdT, bool MASKED, typename OutT>
__global__ void __launch_bounds__(kGatherThreads)
gather_rows_kernel(const float4* __restrict__ table, const IdT* __restrict__ ids,
                   const float* __restrict__ mask, typename OutChunk<OutT>::type* __restrict__ out, int64_t n_rows,
                   int64_t vocab, int cpr_rt, int bulk_ok, int* __restrict__ oob, int64_t pitch /* float4 per table row */) {
  constexpr int TR = tile_rows_for(CPR == 0 ? 16 : CPR);
  const int cpr = (CPR == 0) ? cpr_rt : CPR;
  __shared__ __align__(16) IdT s_ids[2][TR];
  __shared__ __align__(16) float s_mask[2][MASKED ? TR : 4];
  __shared__ __align__(8) uint64_t s_bar[2];

  const int tid = threadIdx.x;
  const int64_t n_tiles = (n_rows + TR - 1) / TR;
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // tile t is "bulk" when it is full and the base pointers are 16-B aligned
  auto is_bulk = [&](int64_t t) { return bulk_ok && (t + 1) * TR <= n_rows; };
  auto prefetch = [&](int64_t t, int st) {
    if (is_bulk(t)) {
      if (tid == 0) {
        uint32_t bytes = TR * sizeof(IdT) + (MASKED ? TR * sizeof(float) : 0);
        mbar_expect_tx(&s_bar[st], bytes);
        bulk_g2s(&s_ids[st][0], ids + t * TR, TR * sizeof(IdT), &s_bar[st]);
        if (MASKED) bulk_g2s(&s_mask[st][0], mask + t * TR, TR * sizeof(float), &s_bar[st]);
      }
    }
  };

  uint32_t phase_bits = 0u;  // bit s = parity the next wait on buffer s expects
  int st = 0;
  int64_t tile = blockIdx.x;
  if (tile < n_tiles) prefetch(tile, 0);

  for (; tile < n_tiles; tile += gridDim.x, st ^= 1) {
    const int64_t next = tile + gridDim.x;
    if (next < n_tiles) prefetch(next, st ^ 1);

    const int64_t row0 = tile * TR;
    const int rows_here = (int)min((int64_t)TR, n_rows - row0);
    if (is_bulk(tile)) {
      mbar_wait(&s_bar[st], (phase_bits >> st) & 1u);
      phase_bits ^= (1u << st);
    } else {
      for (int r = tid; r < rows_here; r += kGatherThreads) {
        s_ids[st][r] = ids[row0 + r];
        if (MASKED) s_mask[st][r] = mask[row0 + r];
      }
      __syncthreads();
    }

    const int n_chunks = rows_here * cpr;
    typename OutChunk<OutT>::type* out_tile = out + row0 * cpr;
    if constexpr (CPR != 0) {
      constexpr int CH = (TR * CPR + kGatherThreads - 1) / kGatherThreads;
      float4 v[CH];
      float mk[CH];
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int c = tid + k * kGatherThreads;
        v[k] = f4_zero();
        mk[k] = 1.f;
        if (c < n_chunks) {
          const int r = c / CPR;
          const int col = c - r * CPR;
          const int64_t id = (int64_t)s_ids[st][r];
          if (MASKED) mk[k] = s_mask[st][r];
          if ((uint64_t)id < (uint64_t)vocab) {
            v[k] = ld_stream_f4(table + id * pitch + col);
          } else {
            if (oob) atomicOr(oob, 1);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int c = tid + k * kGatherThreads;
        if (c < n_chunks) {
          float4 o = v[k];
          if (MASKED) o = f4_scale(o, mk[k]);
          OutChunk<OutT>::store(out_tile + c, o);
        }
      }
    } else {
      for (int c0 = tid; c0 < n_chunks; c0 += 4 * kGatherThreads) {
        float4 v[4];
        float mk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = c0 + k * kGatherThreads;
          v[k] = f4_zero();
          mk[k] = 1.f;
          if (c < n_chunks) {
            const int r = c / cpr;
            const int col = c - r * cpr;
            const int64_t id = (int64_t)s_ids[st][r];
            if (MASKED) mk[k] = s_mask[st][r];
            if ((uint64_t)id < (uint64_t)vocab) {
              v[k] = ld_stream_f4(table + id * pitch + col);
            } else {
              if (oob) atomicOr(oob, 1);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = c0 + k * kGatherThreads;
          if (c < n_chunks) {
            float4 o = v[k];
            if (MASKED) o = f4_scale(o, mk[k]);
            OutChunk<OutT>::store(out_tile + c, o);
          }
        }
      }
    }
    __syncthreads();  // buffer `st` is re-filled two iterations from now
  }
}

// D not a multiple of 4 (rows not 16-B aligned): one thread per element.
template <typename IdT, bool MASKED, typename OutT>
__global__ void gather_scalar_kernel(const float* __restrict__ table, const IdT* __restrict__ ids,
                                     const float* __restrict__ mask, OutT* __restrict__ out,
                                     int64_t n_rows, int dim, int64_t vocab, int* __restrict__ oob) {
  const int64_t total = n_rows * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / dim;
    const int col = (int)(e - r * dim);
    const int64_t id = (int64_t)ids[r];
    float v = 0.f;
    if ((uint64_t)id < (uint64_t)vocab) {
      v = table[id * dim + col];
    } else if (oob) {
      atomicOr(oob, 1);
    }
    if (MASKED) v *= mask[r];
    out[e] = (OutT)v;
  }
}

// dim-1 table: out[b] = sum_f table[ids[b,f]] * mask[b,f] + bias   (one warp per sample)
template <typename IdT>
__global__ void __launch_bounds__(256)
gather_reduce_kernel(const float* __restrict__ table, const IdT* __restrict__ ids,
                     const float* __restrict__ mask, const float* __restrict__ bias,
                     float* __restrict__ out, int64_t batch, int fields, int64_t vocab,
                     int* __restrict__ oob) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float b0 = bias ? bias[0] : 0.f;
  for (int64_t b = warp; b < batch; b += n_warps) {
    float acc = 0.f;
    for (int f0 = 0; f0 < fields; f0 += 64) {
      // two lookups per lane in flight (F = 39 -> one trip)
      const int fa = f0 + lane, fb = f0 + 32 + lane;
      int64_t ia = -1, ib = -1;
      float ma = 0.f, mb = 0.f;
      if (fa < fields) { ia = (int64_t)ids[b * fields + fa]; ma = mask[b * fields + fa]; }
      if (fb < fields) { ib = (int64_t)ids[b * fields + fb]; mb = mask[b * fields + fb]; }
      float wa = 0.f, wb = 0.f;
      if ((uint64_t)ia < (uint64_t)vocab) wa = ld_stream_f1(table + ia);
      else if (fa < fields && oob) atomicOr(oob, 1);
      if ((uint64_t)ib < (uint64_t)vocab) wb = ld_stream_f1(table + ib);
      else if (fb < fields && oob) atomicOr(oob, 1);
      acc = fmaf(wa, ma, acc);
      acc = fmaf(wb, mb, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[b] = acc + b0;
  }
}

// Multi-hot pooled lookup: out[b, :] = (1/S) * sum_s table[ids[b,s], :] * mask[b,s]  (mean over ALL S slots,
// masked ones included, like ReduceMean(axis 1) in models/wide_and_deep_multitable/src/wide_and_deep.py:301-346).
// A group of D/4 threads owns one sample and keeps up to 8 slot rows in flight; the [B,S,D] intermediate of the
// reference never exists.  Slots are added in slot order (deterministic).
template <typename IdT>
__global__ void __launch_bounds__(256)
gather_pool_kernel(const float4* __restrict__ table, const IdT* __restrict__ ids, const float* __restrict__ mask,
                   float4* __restrict__ out, int64_t batch, int slots, int cpr, int64_t vocab, float inv_slots,
                   int* __restrict__ oob) {
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t b = gid / cpr;
  if (b >= batch) return;
  const int c = (int)(gid - b * cpr);
  float4 acc = f4_zero();
  for (int s0 = 0; s0 < slots; s0 += 8) {
    float4 v[8];
    float mk[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int sidx = min(s0 + k, slots - 1);
      const int64_t id = (int64_t)ids[b * slots + sidx];
      const bool ok = (uint64_t)id < (uint64_t)vocab;
      if (!ok && oob && s0 + k < slots) atomicOr(oob, 1);
      v[k] = ld_stream_f4(table + (ok ? id : 0) * cpr + c);
      mk[k] = (ok && s0 + k < slots) ? mask[b * slots + sidx] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) reg_fence(v[k]);
#pragma unroll
    for (int k = 0; k < 8; ++k) f4_fma(acc, v[k], mk[k]);
  }
  st_stream_f4(out + b * cpr + c, f4_scale(acc, inv_slots));
}

template <typename IdT, bool MASKED, typename OutT>
static int launch_gather(const float* table, const IdT* ids, const float* mask, OutT* out,
                         int64_t n_rows, int dim, int64_t vocab, int* oob, cudaStream_t stream, int64_t row_floats) {
  if (n_rows == 0) return OK;
  if (dim % 4 != 0) {
    if (row_floats != dim) return fail(ERR_DIM, "mrec_gather: an interleaved table [V,K,D] needs D %% 4 == 0");
    const int64_t total = n_rows * dim;
    int grid = grid_for(cdiv(total, 256), 16);
    MREC_LAUNCH((gather_scalar_kernel<IdT, MASKED, OutT>), grid, 256, 0, stream, table, ids, mask, out,
                n_rows, dim, vocab, oob);
    return check_launch("gather_scalar");
  }
  const int cpr = dim / 4;
  const int64_t pitch = row_floats / 4;
  const int bulk_ok = ((reinterpret_cast<uintptr_t>(ids) % 16) == 0) &&
                      (!MASKED || (reinterpret_cast<uintptr_t>(mask) % 16) == 0);
  const float4* t4 = reinterpret_cast<const float4*>(table);
  auto* o4 = reinterpret_cast<typename OutChunk<OutT>::type*>(out);
#define MREC_GATHER_CASE(C)                                                                   \
  case C: {                                                                                   \
    constexpr int TR = tile_rows_for(C);                                                      \
    int grid = grid_for(cdiv(n_rows, TR), 8);                     \
    MREC_LAUNCH((gather_rows_kernel<C, IdT, MASKED, OutT>), grid, kGatherThreads, 0, stream, t4, ids, \
                mask, o4, n_rows, vocab, cpr, bulk_ok, oob, pitch);                           \
  } break;
  switch (cpr) {
    MREC_GATHER_CASE(4)
    MREC_GATHER_CASE(8)
    MREC_GATHER_CASE(16)
    MREC_GATHER_CASE(20)
    MREC_GATHER_CASE(32)
    default: {
      constexpr int TR = tile_rows_for(16);
      int grid = grid_for(cdiv(n_rows, TR), 8);
      MREC_LAUNCH((gather_rows_kernel<0, IdT, MASKED, OutT>), grid, kGatherThreads, 0, stream, t4, ids,
                  mask, o4, n_rows, vocab, cpr, bulk_ok, oob, pitch);
    } break;
  }
#undef MREC_GATHER_CASE
  return check_launch("gather_rows");
}

static int gather_entry(const Aot& a, bool masked) {
  // inputs: table[V,D] f32, ids[...] i32|i64, (mask[...] f32) ; outputs: out[N*D] f32, (oob[1] i32)
  const int n_in = masked ? 3 : 2;
  if (a.nparam != n_in + 1 && a.nparam != n_in + 2)
    return fail(ERR_NPARAM, "mrec_gather%s: expected %d or %d params, got %d",
                masked ? "_masked" : "", n_in + 1, n_in + 2, a.nparam);
  for (int i = 0; i < a.nparam; ++i)
    if (!a.params[i] && a.numel(i) > 0) return fail(ERR_NULL, "mrec_gather: param %d is null", i);
  const int o = n_in;
  const bool out16 = a.is(o, "float16");
  MREC_REQUIRE(a.is_f32(0) && (a.is_f32(o) || out16), ERR_DTYPE,
               "mrec_gather: table must be float32, out float32 or float16");
  MREC_REQUIRE(a.is_i32(1) || a.is_i64(1), ERR_DTYPE, "mrec_gather: ids must be int32 or int64");
  // table[V,K,D]: K interleaved arrays per row (the w | m | v record of mrec_sparse_lazy_adam); array 0 is read
  MREC_REQUIRE(a.ndims[0] >= 1 && a.ndims[0] <= 3, ERR_SHAPE, "mrec_gather: table must be [V,D], [V] or [V,K,D]");
  const int64_t vocab = a.dim(0, 0);
  const int dim = a.ndims[0] == 3 ? (int)a.dim(0, 2) : (a.ndims[0] == 2 ? (int)a.dim(0, 1) : 1);
  const int64_t row_floats = a.ndims[0] == 3 ? a.dim(0, 1) * a.dim(0, 2) : dim;
  const int64_t n = a.numel(1);
  MREC_REQUIRE(dim >= 1, ERR_DIM, "mrec_gather: embedding dim must be >= 1");
  MREC_REQUIRE(a.numel(o) == n * dim, ERR_SHAPE, "mrec_gather: out numel %lld != N*D %lld",
               (long long)a.numel(o), (long long)(n * dim));
  if (masked) {
    MREC_REQUIRE(a.is_f32(2), ERR_DTYPE, "mrec_gather_masked: mask must be float32");
    MREC_REQUIRE(a.numel(2) == n, ERR_SHAPE, "mrec_gather_masked: mask numel must equal ids numel");
  }
  if (dim % 4 == 0)
    MREC_REQUIRE(a.aligned(0, 16) && a.aligned(o, out16 ? 8 : 16), ERR_ALIGN,
                 "mrec_gather: table/out must be 16-byte aligned");
  int* oob = nullptr;
  if (a.nparam == n_in + 2) {
    MREC_REQUIRE(a.is_i32(o + 1), ERR_DTYPE, "mrec_gather: oob flag must be int32");
    oob = a.ptr<int>(o + 1);
  }
  const float* mask = masked ? a.ptr<float>(2) : nullptr;
#define MREC_GATHER_DISPATCH(IDT, MASKED_, OUTT)                                                   \
  return launch_gather<IDT, MASKED_, OUTT>(a.ptr<float>(0), a.ptr<IDT>(1), mask, a.ptr<OUTT>(o), n, dim, \
                                           vocab, oob, a.stream, row_floats)
  if (a.is_i32(1)) {
    if (masked) { if (out16) MREC_GATHER_DISPATCH(int32_t, true, __half); MREC_GATHER_DISPATCH(int32_t, true, float); }
    if (out16) MREC_GATHER_DISPATCH(int32_t, false, __half);
    MREC_GATHER_DISPATCH(int32_t, false, float);
  }
  if (masked) { if (out16) MREC_GATHER_DISPATCH(int64_t, true, __half); MREC_GATHER_DISPATCH(int64_t, true, float); }
  if (out16) MREC_GATHER_DISPATCH(int64_t, false, __half);
  MREC_GATHER_DISPATCH(int64_t, false, float);
#undef MREC_GATHER_DISPATCH
}

}  // namespace mrec

using namespace mrec;

MREC_API int mrec_gather(int nparam, void** params, int* ndims, int64_t** shapes,
                           const char** dtypes, void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  return gather_entry(a, false);
}

MREC_API int mrec_gather_masked(int nparam, void** params, int* ndims, int64_t** shapes,
                                  const char** dtypes, void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  return gather_entry(a, true);
}

// inputs: table[V] | [V,1] f32, ids[B,F], mask[B,F] f32, bias[1] f32 ; outputs: out[B] | [B,1], (oob[1])
MREC_API int mrec_gather_reduce(int nparam, void** params, int* ndims, int64_t** shapes,
                                  const char** dtypes, void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  if (a.nparam != 5 && a.nparam != 6)
    return fail(ERR_NPARAM, "mrec_gather_reduce: expected 5 or 6 params, got %d", a.nparam);
  for (int i = 0; i < a.nparam; ++i)
    if (!a.params[i] && a.numel(i) > 0) return fail(ERR_NULL, "mrec_gather_reduce: param %d is null", i);
  MREC_REQUIRE(a.is_f32(0) && a.is_f32(2) && a.is_f32(3) && a.is_f32(4), ERR_DTYPE,
               "mrec_gather_reduce: table/mask/bias/out must be float32");
  MREC_REQUIRE(a.is_i32(1) || a.is_i64(1), ERR_DTYPE, "mrec_gather_reduce: ids must be int32|int64");
  MREC_REQUIRE(a.ndims[1] == 2, ERR_SHAPE, "mrec_gather_reduce: ids must be [B,F]");
  const int64_t vocab = a.dim(0, 0);
  MREC_REQUIRE(a.numel(0) == vocab, ERR_DIM, "mrec_gather_reduce: table must have dim 1");
  const int64_t batch = a.dim(1, 0);
  const int fields = (int)a.dim(1, 1);
  MREC_REQUIRE(a.numel(2) == batch * fields, ERR_SHAPE, "mrec_gather_reduce: mask shape != ids shape");
  MREC_REQUIRE(a.numel(4) == batch, ERR_SHAPE, "mrec_gather_reduce: out must have B elements");
  int* oob = a.nparam == 6 ? a.ptr<int>(5) : nullptr;
  if (batch == 0) return OK;
  int grid = grid_for(cdiv(batch, 8), 32);
  if (a.is_i32(1)) {
    MREC_LAUNCH(gather_reduce_kernel<int32_t>, grid, 256, 0, a.stream, a.ptr<float>(0),
                a.ptr<int32_t>(1), a.ptr<float>(2), a.ptr<float>(3), a.ptr<float>(4), batch, fields,
                vocab, oob);
  } else {
    MREC_LAUNCH(gather_reduce_kernel<int64_t>, grid, 256, 0, a.stream, a.ptr<float>(0),
                a.ptr<int64_t>(1), a.ptr<float>(2), a.ptr<float>(3), a.ptr<float>(4), batch, fields,
                vocab, oob);
  }
  return check_launch("gather_reduce");
}

// in : table[V,D] f32 (D % 4 == 0), ids[B,S] i32|i64, mask[B,S] f32      out: out[B,D] f32, (oob[1] i32)
MREC_API int mrec_gather_pool(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                              void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  if (a.nparam != 4 && a.nparam != 5)
    return fail(ERR_NPARAM, "mrec_gather_pool: expected 4 or 5 params, got %d", a.nparam);
  for (int i = 0; i < a.nparam; ++i)
    if (!a.params[i] && a.numel(i) > 0) return fail(ERR_NULL, "mrec_gather_pool: param %d is null", i);
  MREC_REQUIRE(a.is_f32(0) && a.is_f32(2) && a.is_f32(3), ERR_DTYPE, "mrec_gather_pool: table/mask/out must be float32");
  MREC_REQUIRE(a.is_i32(1) || a.is_i64(1), ERR_DTYPE, "mrec_gather_pool: ids must be int32|int64");
  MREC_REQUIRE(a.ndims[0] == 2 && a.ndims[1] == 2, ERR_SHAPE, "mrec_gather_pool: table[V,D], ids[B,S] expected");
  const int64_t vocab = a.dim(0, 0), batch = a.dim(1, 0);
  const int dim = (int)a.dim(0, 1), slots = (int)a.dim(1, 1);
  MREC_REQUIRE(dim % 4 == 0 && dim >= 4, ERR_DIM, "mrec_gather_pool: D must be a multiple of 4");
  MREC_REQUIRE(slots >= 1 && vocab >= 1, ERR_SHAPE, "mrec_gather_pool: S and V must be >= 1");
  MREC_REQUIRE(a.numel(2) == batch * slots && a.numel(3) == batch * dim, ERR_SHAPE,
               "mrec_gather_pool: mask must be [B,S] and out [B,D]");
  MREC_REQUIRE(a.aligned(0, 16) && a.aligned(3, 16), ERR_ALIGN, "mrec_gather_pool: table/out must be 16-byte aligned");
  int* oob = a.nparam == 5 ? a.ptr<int>(4) : nullptr;
  if (batch == 0) return OK;
  const int cpr = dim / 4;
  const int grid = (int)cdiv(batch * cpr, 256);
  if (a.is_i32(1)) {
    MREC_LAUNCH(gather_pool_kernel<int32_t>, grid, 256, 0, a.stream, a.ptr<float4>(0), a.ptr<int32_t>(1),
                a.ptr<float>(2), a.ptr<float4>(3), batch, slots, cpr, vocab, 1.f / (float)slots, oob);
  } else {
    MREC_LAUNCH(gather_pool_kernel<int64_t>, grid, 256, 0, a.stream, a.ptr<float4>(0), a.ptr<int64_t>(1),
                a.ptr<float>(2), a.ptr<float4>(3), batch, slots, cpr, vocab, 1.f / (float)slots, oob);
  }
  return check_launch("gather_pool");
}
