// K3/K4/K5 — deterministic segment-sum fused with sparse LazyAdam / FTRL row updates; dense Adam/FTRL.
//
// Replaces, for the embedding backward + optimizer step of
//   models/wide_deep/src/wide_and_deep.py:420-445,479-492  (LazyAdam deep, FTRL wide, loss_scale=sens)
//   models/wide_and_deep_multitable/src/wide_and_deep.py:525-535
//   models/deepfm/src/deepfm.py:272, models/deep_and_cross/src/deep_and_cross.py:342-344 (Adam)
// the upstream chain  Gather-bprop -> RowTensor -> Unique -> UnsortedSegmentSum(atomicAdd) ->
// FusedSparseLazyAdam / FusedSparseFtrl  (SURVEY B4-B7).
//
// Input is the stable sort of the lookup ids (mrec_unique: perm / seg_of / seg_start / uniq).
// Phase 1 (segment sum).  The sorted positions are cut into fixed tiles of 32; a group of D/4 threads
// (one float4 column chunk each) walks a tile in order, accumulating mask[p] * g[p] for each run of equal
// keys, 8 independent row loads in flight per thread:
//   * a run that starts and ends inside the tile is complete: its sum is stored to gsum[segment];
//   * a run that crosses a tile edge writes a partial (at most 2 per tile).
// Phase 2 (finish + row update, one launch).  One thread group per unique row reads gsum[u] (still
// L2-resident: U*D*4 bytes, 39 MB at BASELINE config 2 against a 126 MB L2) — or adds the partials of a
// short chain in tile order — and w/m/v[uniq[u]] as independent 16-byte loads, and writes the LazyAdam /
// FTRL result.  Chains longer than kLongChain tiles (Zipf head keys, the 13 dense Criteo fields that
// appear once per sample) are listed by phase 1 and reduced by a few dedicated CTAs of the same launch
// with a fixed-shape tree.  Splitting the update from the tile walk removes the dependent
// load chain that an in-line update puts after every run (ncu r1a: 110 regs, 22 % warps active, 17 %
// of DRAM peak when fused) and lets both phases run at full memory-level parallelism.
// No atomics touch floating-point data: the summation order depends only on the sorted order, so the
// result is bit-reproducible run to run.
#include "common.cuh"
#include "kernels.h"
#include <cuda_fp16.h>
#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace mrec {

#ifndef MREC_SEG_TILE
#define MREC_SEG_TILE 32
#endif
constexpr int kSegTile = MREC_SEG_TILE;  // sorted positions per tile
constexpr int kSegBatch = 8;      // independent row loads in flight per thread
constexpr int kLongChain = 16;    // partial chains longer than this go to the CTA kernel
constexpr int kSegThreads = 256;

// ---- vector helpers ----
template <typename Vec> struct VOps;
template <> struct VOps<float4> {
  static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ void fma(float4& a, const float4& x, float s) { f4_fma(a, x, s); }
  static __device__ __forceinline__ void add(float4& a, const float4& x) {
    a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
  }
  // gradient rows: fp32, or fp16 (the DenseLayer backward emits fp16 when use_mixed_precision; the
  // Cast-to-fp32 of wide_and_deep.py:119 bprop is fused into this load)
  // Raw loads are kept apart from the conversion so that a batch of loads can be issued back to back
  // (ncu r1c: with load + convert inside one conditional the compiler serialised all 8 row loads).
  using RawF = float4;
  using RawH = uint2;
  static __device__ __forceinline__ float4 ld_raw(const float* g, int64_t chunk) {
    return ld_stream_f4(reinterpret_cast<const float4*>(g) + chunk);
  }
  static __device__ __forceinline__ uint2 ld_raw(const __half* g, int64_t chunk) {
    uint2 u;
    asm volatile("ld.global.nc.L1::no_allocate.v2.b32 {%0,%1}, [%2];"
                 : "=r"(u.x), "=r"(u.y) : "l"(reinterpret_cast<const uint2*>(g) + chunk));
    return u;
  }
  static __device__ __forceinline__ float4 cvt(const float4& r) { return r; }
  static __device__ __forceinline__ float4 cvt(const uint2& u) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
};
template <> struct VOps<F8> {
  static __device__ __forceinline__ F8 zero() { F8 z; z.lo = make_float4(0.f, 0.f, 0.f, 0.f); z.hi = z.lo; return z; }
  static __device__ __forceinline__ void add(F8& a, const F8& x) {
    a.lo.x += x.lo.x; a.lo.y += x.lo.y; a.lo.z += x.lo.z; a.lo.w += x.lo.w;
    a.hi.x += x.hi.x; a.hi.y += x.hi.y; a.hi.z += x.hi.z; a.hi.w += x.hi.w;
  }
};
template <> struct VOps<float> {
  static __device__ __forceinline__ float zero() { return 0.f; }
  static __device__ __forceinline__ void fma(float& a, const float& x, float s) { a = fmaf(x, s, a); }
  static __device__ __forceinline__ void add(float& a, const float& x) { a += x; }
  static __device__ __forceinline__ float ld_raw(const float* g, int64_t i) { return ld_stream_f1(g + i); }
  static __device__ __forceinline__ float ld_raw(const __half* g, int64_t i) { return __half2float(g[i]); }
  static __device__ __forceinline__ float cvt(const float& r) { return r; }
};

// ---- hyper-parameter blocks (device f32 tensors, so schedules / bias-correction never sync the host)
// Adam : [0] lr [1] beta1 [2] beta2 [3] eps [4] beta1_power [5] beta2_power [6] lr_t [7] grad_scale
//        [8] l2 (dense-mode table regulariser: g += l2 * w after scaling, wide_and_deep.py:359-360)
// FTRL : [0] lr [1] l1 [2] l2 [3] lr_power [4] grad_scale
constexpr int kHyperLen = 16;

__device__ __forceinline__ void adam_elem(float& w, float& m, float& v, float g, float b1, float b2,
                                          float eps, float lr_t) {
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  w = w - lr_t * m / (sqrtf(v) + eps);
}

__device__ __forceinline__ void ftrl_elem(float& w, float& a, float& lin, float g, float lr, float l1,
                                          float l2, float lr_power) {
  const float a_new = a + g * g;
  float pa_new, pa_old;
  if (lr_power == -0.5f) {
    pa_new = sqrtf(a_new);
    pa_old = sqrtf(a);
  } else {
    pa_new = powf(a_new, -lr_power);
    pa_old = powf(a, -lr_power);
  }
  const float sigma = (pa_new - pa_old) / lr;
  lin = lin + g - sigma * w;
  const float q = pa_new / lr + 2.f * l2;
  const float sgn = lin > 0.f ? 1.f : (lin < 0.f ? -1.f : 0.f);
  w = (fabsf(lin) > l1) ? (sgn * l1 - lin) / q : 0.f;
  a = a_new;
}

// ---- per-segment sinks ----
// A sink owns the per-row state of an optimizer.  The row pass drives it in three steps so that the loads of
// several rows can be in flight together: row_of(seg) (one index load), load(offset, state) (unconditional loads
// of the state chunks; the caller clamps the offset of skipped rows to a valid one), finish(offset, state, gsum)
// (math + stores).  apply() is the one-row form used by the chain paths.
template <typename T> struct State3 { T a, b, c; };
struct State0 {};

template <typename Vec, typename IdT> struct LazyAdamSink;
template <typename IdT>
struct LazyAdamSink<float4, IdT> {
  static constexpr bool kApplyComplete = true;
  using State = State3<float4>;
  float4* w; float4* m; float4* v;
  const IdT* uniq;
  const float* hyper;
  int64_t vocab;
  int cpr;
  int64_t pitch;   // Vec units between consecutive rows: cpr, or 3 * cpr for the interleaved w|m|v record
  __device__ __forceinline__ int64_t offset(int64_t row, int c) const { return row * pitch + c; }
  __device__ __forceinline__ int64_t row_of(int seg) const { return (int64_t)uniq[seg]; }
  __device__ __forceinline__ bool in_range(int64_t row) const { return (uint64_t)row < (uint64_t)vocab; }
  __device__ __forceinline__ void load(int64_t o, State& s) const { s.a = w[o]; s.b = m[o]; s.c = v[o]; }
  __device__ __forceinline__ void fence(State& s) const { reg_fence(s.a); reg_fence(s.b); reg_fence(s.c); }
  __device__ __forceinline__ void finish(int64_t o, State& s, const float4& gs) const {
    const float b1 = hyper[1], b2 = hyper[2], eps = hyper[3], lr_t = hyper[6], sc = hyper[7];
    const float l2 = hyper[8];
    float4 W = s.a, M = s.b, V = s.c;
    adam_elem(W.x, M.x, V.x, fmaf(l2, W.x, gs.x * sc), b1, b2, eps, lr_t);
    adam_elem(W.y, M.y, V.y, fmaf(l2, W.y, gs.y * sc), b1, b2, eps, lr_t);
    adam_elem(W.z, M.z, V.z, fmaf(l2, W.z, gs.z * sc), b1, b2, eps, lr_t);
    adam_elem(W.w, M.w, V.w, fmaf(l2, W.w, gs.w * sc), b1, b2, eps, lr_t);
    w[o] = W; m[o] = M; v[o] = V;
  }
  __device__ __forceinline__ void apply(int seg, int c, const float4& gs) const {
    const int64_t row = row_of(seg);
    if (!in_range(row)) return;                    // out-of-range ids carry no row
    State s;
    load(offset(row, c), s);
    finish(offset(row, c), s, gs);
  }
};
// 32-byte form of the LazyAdam row update: D/8 threads per row, LDG.256 / STG.256 with L2 evict-first on w, m, v
// (each touched once per step), so the L2 keeps gsum and the hot Zipf rows instead.  Same per-element math.
template <typename IdT>
struct LazyAdamSink<F8, IdT> {
  static constexpr bool kApplyComplete = true;
  using State = State3<F8>;
  F8* w; F8* m; F8* v;
  const IdT* uniq;
  const float* hyper;
  int64_t vocab;
  int cpr;
  int64_t pitch;   // Vec units between consecutive rows: cpr, or 3 * cpr for the interleaved w|m|v record
  __device__ __forceinline__ int64_t offset(int64_t row, int c) const { return row * pitch + c; }
  __device__ __forceinline__ int64_t row_of(int seg) const { return (int64_t)uniq[seg]; }
  __device__ __forceinline__ bool in_range(int64_t row) const { return (uint64_t)row < (uint64_t)vocab; }
  __device__ __forceinline__ void load(int64_t o, State& s) const {
    s.a = ld256_evict_first(w + o); s.b = ld256_evict_first(m + o); s.c = ld256_evict_first(v + o);
  }
  __device__ __forceinline__ void fence(State& s) const { reg_fence(s.a); reg_fence(s.b); reg_fence(s.c); }
  __device__ __forceinline__ void finish(int64_t o, State& s, const F8& gs) const {
    const float b1 = hyper[1], b2 = hyper[2], eps = hyper[3], lr_t = hyper[6], sc = hyper[7];
    const float l2 = hyper[8];
    F8 W = s.a, M = s.b, V = s.c;
    adam_elem(W.lo.x, M.lo.x, V.lo.x, fmaf(l2, W.lo.x, gs.lo.x * sc), b1, b2, eps, lr_t);
    adam_elem(W.lo.y, M.lo.y, V.lo.y, fmaf(l2, W.lo.y, gs.lo.y * sc), b1, b2, eps, lr_t);
    adam_elem(W.lo.z, M.lo.z, V.lo.z, fmaf(l2, W.lo.z, gs.lo.z * sc), b1, b2, eps, lr_t);
    adam_elem(W.lo.w, M.lo.w, V.lo.w, fmaf(l2, W.lo.w, gs.lo.w * sc), b1, b2, eps, lr_t);
    adam_elem(W.hi.x, M.hi.x, V.hi.x, fmaf(l2, W.hi.x, gs.hi.x * sc), b1, b2, eps, lr_t);
    adam_elem(W.hi.y, M.hi.y, V.hi.y, fmaf(l2, W.hi.y, gs.hi.y * sc), b1, b2, eps, lr_t);
    adam_elem(W.hi.z, M.hi.z, V.hi.z, fmaf(l2, W.hi.z, gs.hi.z * sc), b1, b2, eps, lr_t);
    adam_elem(W.hi.w, M.hi.w, V.hi.w, fmaf(l2, W.hi.w, gs.hi.w * sc), b1, b2, eps, lr_t);
    st256_evict_first(w + o, W); st256_evict_first(m + o, M); st256_evict_first(v + o, V);
  }
  __device__ __forceinline__ void apply(int seg, int c, const F8& gs) const {
    const int64_t row = row_of(seg);
    if (!in_range(row)) return;
    State s;
    load(offset(row, c), s);
    finish(offset(row, c), s, gs);
  }
};
template <typename IdT>
struct LazyAdamSink<float, IdT> {
  static constexpr bool kApplyComplete = true;
  using State = State3<float>;
  float* w; float* m; float* v;
  const IdT* uniq;
  const float* hyper;
  int64_t vocab;
  int cpr;
  int64_t pitch;   // Vec units between consecutive rows: cpr, or 3 * cpr for the interleaved w|m|v record
  __device__ __forceinline__ int64_t offset(int64_t row, int c) const { return row * pitch + c; }
  __device__ __forceinline__ int64_t row_of(int seg) const { return (int64_t)uniq[seg]; }
  __device__ __forceinline__ bool in_range(int64_t row) const { return (uint64_t)row < (uint64_t)vocab; }
  __device__ __forceinline__ void load(int64_t o, State& s) const { s.a = w[o]; s.b = m[o]; s.c = v[o]; }
  __device__ __forceinline__ void fence(State& s) const { reg_fence(s.a); reg_fence(s.b); reg_fence(s.c); }
  __device__ __forceinline__ void finish(int64_t o, State& s, const float& gs) const {
    float W = s.a, M = s.b, V = s.c;
    adam_elem(W, M, V, fmaf(hyper[8], W, gs * hyper[7]), hyper[1], hyper[2], hyper[3], hyper[6]);
    w[o] = W; m[o] = M; v[o] = V;
  }
  __device__ __forceinline__ void apply(int seg, int c, const float& gs) const {
    const int64_t row = row_of(seg);
    if (!in_range(row)) return;
    State s;
    load(offset(row, c), s);
    finish(offset(row, c), s, gs);
  }
};

template <typename Vec, typename IdT> struct FtrlSink;
template <typename IdT>
struct FtrlSink<float4, IdT> {
  static constexpr bool kApplyComplete = true;
  using State = State3<float4>;
  float4* w; float4* acc; float4* lin;
  const IdT* uniq;
  const float* hyper;
  int64_t vocab;
  int cpr;
  __device__ __forceinline__ int64_t offset(int64_t row, int c) const { return row * cpr + c; }
  __device__ __forceinline__ int64_t row_of(int seg) const { return (int64_t)uniq[seg]; }
  __device__ __forceinline__ bool in_range(int64_t row) const { return (uint64_t)row < (uint64_t)vocab; }
  __device__ __forceinline__ void load(int64_t o, State& s) const { s.a = w[o]; s.b = acc[o]; s.c = lin[o]; }
  __device__ __forceinline__ void fence(State& s) const { reg_fence(s.a); reg_fence(s.b); reg_fence(s.c); }
  __device__ __forceinline__ void finish(int64_t o, State& s, const float4& gs) const {
    const float lr = hyper[0], l1 = hyper[1], l2 = hyper[2], p = hyper[3], sc = hyper[4];
    float4 W = s.a, A = s.b, L = s.c;
    ftrl_elem(W.x, A.x, L.x, gs.x * sc, lr, l1, l2, p);
    ftrl_elem(W.y, A.y, L.y, gs.y * sc, lr, l1, l2, p);
    ftrl_elem(W.z, A.z, L.z, gs.z * sc, lr, l1, l2, p);
    ftrl_elem(W.w, A.w, L.w, gs.w * sc, lr, l1, l2, p);
    w[o] = W; acc[o] = A; lin[o] = L;
  }
  __device__ __forceinline__ void apply(int seg, int c, const float4& gs) const {
    const int64_t row = row_of(seg);
    if (!in_range(row)) return;
    State s;
    load(row * cpr + c, s);
    finish(row * cpr + c, s, gs);
  }
};
template <typename IdT>
struct FtrlSink<float, IdT> {
  static constexpr bool kApplyComplete = true;
  using State = State3<float>;
  float* w; float* acc; float* lin;
  const IdT* uniq;
  const float* hyper;
  int64_t vocab;
  int cpr;
  __device__ __forceinline__ int64_t offset(int64_t row, int c) const { return row * cpr + c; }
  __device__ __forceinline__ int64_t row_of(int seg) const { return (int64_t)uniq[seg]; }
  __device__ __forceinline__ bool in_range(int64_t row) const { return (uint64_t)row < (uint64_t)vocab; }
  __device__ __forceinline__ void load(int64_t o, State& s) const { s.a = w[o]; s.b = acc[o]; s.c = lin[o]; }
  __device__ __forceinline__ void fence(State& s) const { reg_fence(s.a); reg_fence(s.b); reg_fence(s.c); }
  __device__ __forceinline__ void finish(int64_t o, State& s, const float& gs) const {
    float W = s.a, A = s.b, L = s.c;
    ftrl_elem(W, A, L, gs * hyper[4], hyper[0], hyper[1], hyper[2], hyper[3]);
    w[o] = W; acc[o] = A; lin[o] = L;
  }
  __device__ __forceinline__ void apply(int seg, int c, const float& gs) const {
    const int64_t row = row_of(seg);
    if (!in_range(row)) return;
    State s;
    load(row * cpr + c, s);
    finish(row * cpr + c, s, gs);
  }
};

// Dense-gradient accumulation (UnsortedSegmentSum into a [V, D] gradient table, the bprop of a non-sparse
// `Gather`): table[uniq[seg]] += sum.  Rows of one call are distinct, so there are no atomics and the result is
// deterministic; several calls (several inputs looking up the same table) accumulate in call order.
template <typename Vec, typename IdT>
struct ScatterAddSink {
  static constexpr bool kApplyComplete = true;
  struct State { Vec a; };
  Vec* table;
  const IdT* uniq;
  int64_t vocab;
  int cpr;
  __device__ __forceinline__ int64_t offset(int64_t row, int c) const { return row * cpr + c; }
  __device__ __forceinline__ int64_t row_of(int seg) const { return (int64_t)uniq[seg]; }
  __device__ __forceinline__ bool in_range(int64_t row) const { return (uint64_t)row < (uint64_t)vocab; }
  __device__ __forceinline__ void load(int64_t o, State& s) const { s.a = table[o]; }
  __device__ __forceinline__ void fence(State& s) const { reg_fence(s.a); }
  __device__ __forceinline__ void finish(int64_t o, State& s, const Vec& gs) const {
    Vec t = s.a;
    VOps<Vec>::add(t, gs);
    table[o] = t;
  }
  __device__ __forceinline__ void apply(int seg, int c, const Vec& gs) const {
    const int64_t row = row_of(seg);
    if (!in_range(row)) return;
    State s;
    load(row * cpr + c, s);
    finish(row * cpr + c, s, gs);
  }
};

// ---- kernel A: walk tiles ----
template <typename Vec, typename GT, bool HAS_MASK>
__global__ void __launch_bounds__(kSegThreads, 3)
segsum_tiles_kernel(const GT* __restrict__ g, int cpr, int div, const float* __restrict__ mask,
                    const int32_t* __restrict__ perm, const int32_t* __restrict__ seg_of,
                    const int32_t* __restrict__ seg_start, int64_t n, int64_t n_tiles, Vec* __restrict__ part,
                    Vec* __restrict__ gsum, int32_t* __restrict__ long_list, int32_t* __restrict__ long_count,
                    const int32_t* __restrict__ n_valid) {
  // n_valid (optional): only the first *n_valid sorted positions are real, the rest is padding of a static buffer
  if (n_valid) n = min(n, (int64_t)n_valid[0]);
  const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t j = gid / cpr;
  if (j >= n_tiles) return;
  const int c = (int)(gid - j * cpr);
  const int64_t pos0 = j * kSegTile;
  if (pos0 >= n) return;
  const int64_t pos1 = min(n, pos0 + kSegTile);

  int cur = seg_of[pos0];
  bool enters = (pos0 > 0) && (seg_of[pos0 - 1] == cur);
  Vec acc = VOps<Vec>::zero();

  for (int64_t b0 = pos0; b0 < pos1; b0 += kSegBatch) {
    // Positions past the end of the tile are clamped to its last position with a zero weight: every load
    // of the batch is unconditional, so all kSegBatch row loads are in flight together.
    int seg[kSegBatch];
    int32_t p[kSegBatch];
    float mk[kSegBatch];
    decltype(VOps<Vec>::ld_raw(g, 0)) raw[kSegBatch];
#pragma unroll
    for (int k = 0; k < kSegBatch; ++k) {
      const int64_t i = min(b0 + k, pos1 - 1);
      seg[k] = seg_of[i];
      p[k] = perm[i];
    }
#pragma unroll
    for (int k = 0; k < kSegBatch; ++k) {
      const int64_t grow = (div == 1) ? (int64_t)p[k] : (int64_t)(p[k] / div);
      raw[k] = VOps<Vec>::ld_raw(g, grow * cpr + c);
      mk[k] = HAS_MASK ? mask[p[k]] : 1.f;
    }
#pragma unroll
    for (int k = 0; k < kSegBatch; ++k) reg_fence(raw[k]);
#pragma unroll
    for (int k = 0; k < kSegBatch; ++k) {
      if (b0 + k >= pos1) mk[k] = 0.f;
      if (seg[k] != cur) {
        // run [.., here) ended inside the tile
        if (enters) part[(j * 2 + 0) * cpr + c] = acc;
        else gsum[(int64_t)cur * cpr + c] = acc;
        cur = seg[k];
        enters = false;
        acc = VOps<Vec>::zero();
      }
      VOps<Vec>::fma(acc, VOps<Vec>::cvt(raw[k]), mk[k]);
    }
  }
  const bool leaves = (pos1 < n) && (seg_of[pos1] == cur);
  if (enters) {
    part[(j * 2 + 0) * cpr + c] = acc;                   // entered (and maybe also leaves): slot 0
  } else if (leaves) {
    part[(j * 2 + 1) * cpr + c] = acc;                   // starts here, continues right: slot 1
    if (c == 0) {                                        // chains of more than kLongChain tiles get a CTA each
      const int64_t j_last = ((int64_t)seg_start[cur + 1] - 1) / kSegTile;
      if (j_last - j > kLongChain) long_list[atomicAdd(long_count, 1)] = cur;
    }
  } else {
    gsum[(int64_t)cur * cpr + c] = acc;
  }
}


// ---- kernel A, staged form: rows through shared memory with cp.async --------------------------------------
// The register-resident walks above keep a thread alive for 32 positions with only a handful of row loads in
// flight behind a perm -> row dependent chain (ncu r1g: 33 % warps active, 34 % of DRAM peak).  Here a CTA owns
// P consecutive sorted positions: (1) perm / seg_of of the range go to shared memory, (2) EVERY 16-byte chunk of
// the P gradient rows is requested at once with cp.async (LDGSTS: no registers, no scoreboard, the whole range
// in flight), (3) the segmented sums are then walked out of shared memory, one thread per (32-position tile,
// float4 chunk), in exactly the order and with exactly the outputs of segsum_tiles_kernel (bit-identical).
// Many small CTAs per SM overlap one CTA's walk with the others' loads.
// Fused gradient push of the row-sharded path (backward all-to-all): a segment sum is not parked in a
// local gsum[U, D] buffer and copied again, it is stored straight into the inbox of the rank that OWNS the segment's
// key, through the peer-mapped pointer (NVLink).  With keys owner-major, segment `seg` of my sorted unique keys
// belongs to owner o = #{r >= 1 : seg >= bounds[r]} and lands at inbox row inbox_off[o] + seg - bounds[o].
struct PeerDest {
  const int32_t* bounds;      // [G+1] start of every owner's bucket among my sorted unique keys (bounds[G] = valid keys)
  const int32_t* inbox_off;   // [G]   row of my bucket's first key inside owner o's inbox
  const int64_t* peer_ptrs;   // [G]   base address of owner o's inbox as mapped in this process
  int world;
  int64_t cap;                // rows an inbox holds
  int32_t* err;               // bit 1 is raised when a row does not fit (it is dropped)
};
template <typename Vec>
__device__ __forceinline__ Vec* peer_row(const PeerDest& pd, int seg, int cpr) {
  if (seg >= pd.bounds[pd.world]) return nullptr;          // the segment of out-of-range ids carries no row
  int o = 0;
  for (int r = 1; r < pd.world; ++r) o += (seg >= pd.bounds[r]) ? 1 : 0;
  const int64_t slot = (int64_t)pd.inbox_off[o] + (seg - pd.bounds[o]);
  if (slot >= pd.cap) {
    atomicOr(pd.err, 2);
    return nullptr;
  }
  return reinterpret_cast<Vec*>(pd.peer_ptrs[o]) + slot * cpr;
}

template <typename GT>
__device__ __forceinline__ float4 seg_smem_f4(const GT* row, int c);
template <>
__device__ __forceinline__ float4 seg_smem_f4<float>(const float* row, int c) {
  return reinterpret_cast<const float4*>(row)[c];
}
template <>
__device__ __forceinline__ float4 seg_smem_f4<__half>(const __half* row, int c) {
  const uint2 u = reinterpret_cast<const uint2*>(row)[c];
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

template <typename GT, bool HAS_MASK, bool PEER = false>
__global__ void __launch_bounds__(kSegThreads)
segsum_stage_kernel(const GT* __restrict__ g, int dim, int div, const float* __restrict__ mask,
                    const int32_t* __restrict__ perm, const int32_t* __restrict__ seg_of,
                    const int32_t* __restrict__ seg_start, int64_t n64, int P, float4* __restrict__ part,
                    float4* __restrict__ gsum, int32_t* __restrict__ long_list, int32_t* __restrict__ long_count,
                    const int32_t* __restrict__ n_valid, PeerDest pd = PeerDest{}) {
  extern __shared__ uint4 seg_smem[];
  int n = (int)n64;
  if (n_valid) n = min(n, n_valid[0]);
  const int pos0 = blockIdx.x * P;
  if (pos0 >= n) return;
  const int cnt = min(P, n - pos0);
  const int row_bytes = dim * (int)sizeof(GT);
  char* s_rows = reinterpret_cast<char*>(seg_smem);
  int32_t* s_perm = reinterpret_cast<int32_t*>(s_rows + (size_t)P * row_bytes);
  int32_t* s_seg = s_perm + P;
  float* s_mask = reinterpret_cast<float*>(s_seg + P);
  const int tid = threadIdx.x;

  for (int i = tid; i < cnt; i += kSegThreads) {
    s_perm[i] = perm[pos0 + i];
    s_seg[i] = seg_of[pos0 + i];
  }
  __syncthreads();
  const int cpr16 = row_bytes >> 4;
  const uint32_t s_base = smem_u32(s_rows);
  for (int q = tid; q < cnt * cpr16; q += kSegThreads) {
    const int r = q / cpr16;
    const int c = q - r * cpr16;
    const int p = s_perm[r];
    const int64_t grow = (div == 1) ? (int64_t)p : (int64_t)(p / div);
    const char* src = reinterpret_cast<const char*>(g) + (grow * cpr16 + c) * 16;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_base + (uint32_t)q * 16u), "l"(src) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = tid; i < cnt; i += kSegThreads) s_mask[i] = HAS_MASK ? __ldg(mask + s_perm[i]) : 1.f;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int cpr = dim >> 2;
  // a complete segment sum goes to the local gsum row, or (PEER) straight into its owner's inbox
  auto store_sum = [&](int seg, int c, const float4& a) {
    if constexpr (PEER) {
      float4* r = peer_row<float4>(pd, seg, cpr);
      if (r) r[c] = a;
    } else {
      gsum[(int64_t)seg * cpr + c] = a;
    }
  };
  const int tiles = (cnt + kSegTile - 1) / kSegTile;
  for (int t = tid; t < tiles * cpr; t += kSegThreads) {
    const int tl = t / cpr;
    const int c = t - tl * cpr;
    const int j = pos0 / kSegTile + tl;              // global tile index (P is a multiple of kSegTile)
    const int l0 = tl * kSegTile;
    const int l1 = min(cnt, l0 + kSegTile);
    int cur = s_seg[l0];
    bool enters;
    if (l0 > 0) enters = (s_seg[l0 - 1] == cur);
    else enters = (pos0 > 0) && (seg_of[pos0 - 1] == cur);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int l = l0; l < l1; ++l) {
      const int sg = s_seg[l];
      if (sg != cur) {
        if (enters) part[(int64_t)(j * 2 + 0) * cpr + c] = acc;
        else store_sum(cur, c, acc);
        cur = sg;
        enters = false;
        acc = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const float4 v = seg_smem_f4<GT>(reinterpret_cast<const GT*>(s_rows + (size_t)l * row_bytes), c);
      f4_fma(acc, v, s_mask[l]);
    }
    bool leaves = false;
    if (pos0 + l1 < n) leaves = ((l1 < cnt) ? s_seg[l1] : seg_of[pos0 + l1]) == cur;
    if (enters) {
      part[(int64_t)(j * 2 + 0) * cpr + c] = acc;
    } else if (leaves) {
      part[(int64_t)(j * 2 + 1) * cpr + c] = acc;
      if (c == 0) {
        const int j_last = (seg_start[cur + 1] - 1) / kSegTile;
        if (j_last - j > kLongChain) long_list[atomicAdd(long_count, 1)] = cur;
      }
    } else {
      store_sum(cur, c, acc);
    }
  }
}

// ---- kernel B: finish every segment and hand its sum to the sink -------------------------------------
// A segment [s, e) of the sorted order is "complete" when it lies inside one tile (its sum is already in
// gsum), else it is a chain of partials: slot 1 of tile s/32, then slot 0 of every following tile up to
// (e-1)/32, added in tile order.
//   * blocks >= kChainBlocks: one D/4-thread group per unique row; complete rows and chains of up to
//     kLongChain tiles are summed in line (independent loads, fixed order) and passed to sink.apply();
//   * blocks <  kChainBlocks: one CTA per entry of the long-chain list (Zipf head keys, the dense Criteo
//     fields): group q adds pieces q, q+G, ... with 4 loads in flight, the G group partials are added in
//     group order, then sink.apply().  These CTAs run concurrently with the row groups of the same launch.
// Sink::kApplyComplete = false (plain segment-sum into gsum) skips rows whose sum is already stored.
constexpr int kChainBlocks = 64;

template <typename Vec, typename Sink, int ROWS>
__global__ void __launch_bounds__(kSegThreads, (sizeof(Vec) == 16 ? (ROWS == 2 ? 3 : 4) : (sizeof(Vec) == 32 ? 3 : 4)))
rows_update_kernel(int cpr, const int32_t* __restrict__ seg_of, const int32_t* __restrict__ seg_start,
                   int64_t n, const Vec* __restrict__ gsum, const Vec* __restrict__ part,
                   const int32_t* __restrict__ long_list, const int32_t* __restrict__ long_count, Sink sink,
                   const int32_t* __restrict__ n_valid) {
  __shared__ Vec s_part[kSegThreads];
  if (n_valid) n = min(n, (int64_t)n_valid[0]);
  if (n <= 0) return;
  const int gi = threadIdx.x / cpr;
  const int c = threadIdx.x - gi * cpr;
  const int groups = kSegThreads / cpr;
  if (blockIdx.x < kChainBlocks) {
    const int n_long = *long_count;
    const int lg = min(groups, 32);
    for (int idx = blockIdx.x; idx < n_long; idx += kChainBlocks) {
      const int u = long_list[idx];
      const int64_t j = seg_start[u] / kSegTile;
      const int64_t pieces = ((int64_t)seg_start[u + 1] - 1) / kSegTile - j + 1;
      if (gi < lg) {
        Vec acc = VOps<Vec>::zero();
        for (int64_t k0 = gi; k0 < pieces; k0 += 4 * (int64_t)lg) {
          Vec v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int64_t k = k0 + (int64_t)q * lg;
            v[q] = VOps<Vec>::zero();
            if (k < pieces) v[q] = part[((k == 0) ? (j * 2 + 1) : ((j + k) * 2)) * cpr + c];
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) VOps<Vec>::add(acc, v[q]);
        }
        s_part[gi * cpr + c] = acc;
      }
      __syncthreads();
      if (gi == 0) {
        Vec acc = s_part[c];
        for (int q = 1; q < lg; ++q) VOps<Vec>::add(acc, s_part[q * cpr + c]);
        sink.apply(u, c, acc);
      }
      __syncthreads();
    }
    return;
  }
  if (gi >= groups) return;
  const int n_seg = seg_of[n - 1] + 1;  // U: segment id of the last sorted position + 1
  const int64_t n_groups = (int64_t)(gridDim.x - kChainBlocks) * groups;
  // ROWS rows per thread and iteration: the index loads of all of them, then all their state / gsum chunk loads,
  // are issued before any is used (ncu r1h: with one row at a time the chain seg_start -> uniq -> state left
  // ~16 bytes in flight per thread and 26 % issue activity).  Rows past the end are clamped and discarded.
  for (int64_t u0 = (int64_t)(blockIdx.x - kChainBlocks) * groups + gi; u0 < n_seg; u0 += ROWS * n_groups) {
    int s0[ROWS], s1[ROWS];
    int64_t row[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      const int uc = (int)min(u0 + k * n_groups, (int64_t)n_seg - 1);
      s0[k] = seg_start[uc];
      s1[k] = seg_start[uc + 1];
      row[k] = sink.row_of(uc);
    }
    Vec gs[ROWS];
    typename Sink::State st[ROWS];
    int64_t off[ROWS];
    bool ok[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      const int64_t u = u0 + k * n_groups;
      const int uc = (int)min(u, (int64_t)n_seg - 1);
      ok[k] = (u < n_seg) && sink.in_range(row[k]);
      off[k] = sink.offset(ok[k] ? row[k] : 0, c);
      if (Sink::kApplyComplete) gs[k] = gsum[(int64_t)uc * cpr + c];
      sink.load(off[k], st[k]);
    }
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      if (Sink::kApplyComplete) reg_fence(gs[k]);
      sink.fence(st[k]);
    }
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      const int64_t u = u0 + k * n_groups;
      if (u >= n_seg) continue;
      const int64_t j0 = s0[k] / kSegTile;
      const int64_t j1 = ((int64_t)s1[k] - 1) / kSegTile;
      if (j0 == j1) {
        if (Sink::kApplyComplete && ok[k]) sink.finish(off[k], st[k], gs[k]);
      } else if (j1 - j0 <= kLongChain) {
        Vec acc = part[(j0 * 2 + 1) * cpr + c];
        for (int64_t jj = j0 + 1; jj <= j1; ++jj) VOps<Vec>::add(acc, part[(jj * 2) * cpr + c]);
        if (ok[k]) sink.finish(off[k], st[k], acc);
      }
    }
  }
}

// plain segment-sum: gsum[u] = total (complete rows are already in place)
template <typename Vec>
struct StoreSink {
  static constexpr bool kApplyComplete = false;
  using State = State0;
  Vec* out;
  int cpr;
  __device__ __forceinline__ int64_t offset(int64_t row, int c) const { return row * cpr + c; }
  __device__ __forceinline__ int64_t row_of(int seg) const { return seg; }
  __device__ __forceinline__ bool in_range(int64_t) const { return true; }
  __device__ __forceinline__ void load(int64_t, State&) const {}
  __device__ __forceinline__ void fence(State&) const {}
  __device__ __forceinline__ void finish(int64_t o, State&, const Vec& gs) const { out[o] = gs; }
  __device__ __forceinline__ void apply(int seg, int c, const Vec& gs) const { out[(int64_t)seg * cpr + c] = gs; }
};

// chain totals of the fused gradient push: rows_update_kernel hands every multi-tile segment's sum to this sink
template <typename Vec>
struct PeerStoreSink {
  static constexpr bool kApplyComplete = false;
  using State = State0;
  PeerDest pd;
  int cpr;
  __device__ __forceinline__ int64_t offset(int64_t row, int c) const { return row * cpr + c; }
  __device__ __forceinline__ int64_t row_of(int seg) const { return seg; }
  __device__ __forceinline__ bool in_range(int64_t seg) const { return seg < pd.bounds[pd.world]; }
  __device__ __forceinline__ void load(int64_t, State&) const {}
  __device__ __forceinline__ void fence(State&) const {}
  __device__ __forceinline__ void finish(int64_t o, State&, const Vec& gs) const {
    const int seg = (int)(o / cpr);
    apply(seg, (int)(o - (int64_t)seg * cpr), gs);
  }
  __device__ __forceinline__ void apply(int seg, int c, const Vec& gs) const {
    Vec* r = peer_row<Vec>(pd, seg, cpr);
    if (r) r[c] = gs;
  }
};

struct SegWorkspace {
  size_t off_part, off_list, off_count, off_gsum, total;
};
static SegWorkspace seg_ws(int64_t n, int dim, bool with_gsum) {
  const int64_t n_tiles = cdiv(n > 0 ? n : 1, kSegTile);
  SegWorkspace w;
  size_t o = 0;
  w.off_part = o; o = align_up(o + (size_t)n_tiles * 2 * dim * 4, 256);
  w.off_list = o; o = align_up(o + (size_t)n_tiles * 4, 256);
  w.off_count = o; o = align_up(o + 16, 256);
  w.off_gsum = o;
  if (with_gsum) o = align_up(o + (size_t)(n > 0 ? n : 1) * dim * 4, 256);
  w.total = o;
  return w;
}
size_t sparse_opt_workspace_bytes(int64_t n, int dim) { return seg_ws(n, dim, true).total; }
size_t segment_sum_workspace_bytes(int64_t n, int dim) { return seg_ws(n, dim, false).total; }

struct NoSink {};
template <typename S> struct IsLazyAdamF4 { static constexpr bool value = false; using Id = int; };
template <typename IdT> struct IsLazyAdamF4<LazyAdamSink<float4, IdT>> { static constexpr bool value = true; using Id = IdT; };

// Sink = NoSink: stand-alone segment sum into `out`; otherwise the sums go to the workspace and the sink
// (optimizer update) is applied per unique row.  Two launches: tile walk, then finish + sink.
template <typename Vec, typename GT, typename Sink>
static int run_segsum_t(const GT* g, int dim, int div, const float* mask, const int32_t* perm,
                        const int32_t* seg_start, const int32_t* seg_of, int64_t n, void* ws,
                        size_t ws_bytes, Sink sink, Vec* out, cudaStream_t stream, const int32_t* n_valid) {
  if (n == 0) return OK;
  constexpr bool kStandalone = std::is_same<Sink, NoSink>::value;
  const SegWorkspace W = seg_ws(n, dim, !kStandalone);
  if (ws_bytes < W.total)
    return fail(ERR_WORKSPACE, "sparse update: workspace %zu bytes < required %zu", ws_bytes, W.total);
  if (reinterpret_cast<uintptr_t>(ws) % 16 != 0)
    return fail(ERR_ALIGN, "sparse update: workspace must be 16-byte aligned");
  char* w = reinterpret_cast<char*>(ws);
  Vec* gsum = kStandalone ? out : reinterpret_cast<Vec*>(w + W.off_gsum);
  const int cpr = dim / (int)(sizeof(Vec) / 4);
  Vec* part = reinterpret_cast<Vec*>(w + W.off_part);
  int32_t* long_list = reinterpret_cast<int32_t*>(w + W.off_list);
  int32_t* long_count = reinterpret_cast<int32_t*>(w + W.off_count);
  const int64_t n_tiles = cdiv(n, kSegTile);
  const int grid = (int)cdiv(n_tiles * cpr, kSegThreads);
  cudaMemsetAsync(long_count, 0, sizeof(int32_t), stream);
  // MREC_SEG_MODE: -1 (default) staged cp.async kernel | 0 register tile walk
  const char* seg_env = getenv("MREC_SEG_MODE");
  const int seg_batch = seg_env ? atoi(seg_env) : -1;
  const int row_bytes = dim * (int)sizeof(GT);
  int stage_p = 0;                                  // positions per CTA of the staged kernel (0: not applicable)
  if (std::is_same<Vec, float4>::value && seg_batch < 0 && row_bytes % 16 == 0 && reinterpret_cast<uintptr_t>(g) % 16 == 0 &&
      n < ((int64_t)1 << 31)) {
    const char* pe = getenv("MREC_SEG_P");
    stage_p = pe ? atoi(pe) : 128;
    while (stage_p > 32 && (size_t)stage_p * (row_bytes + 12) > 40 * 1024) stage_p >>= 1;
    if ((size_t)stage_p * (row_bytes + 12) > 40 * 1024) stage_p = 0;
  }
  if (stage_p) {
    if constexpr (std::is_same<Vec, float4>::value) {
      const size_t smem = (size_t)stage_p * (row_bytes + 12);
      const int grid_s = (int)cdiv(n, stage_p);
      if (mask) {
        MREC_LAUNCH((segsum_stage_kernel<GT, true>), grid_s, kSegThreads, smem, stream, g, dim, div, mask, perm, seg_of,
                    seg_start, n, stage_p, part, gsum, long_list, long_count, n_valid);
      } else {
        MREC_LAUNCH((segsum_stage_kernel<GT, false>), grid_s, kSegThreads, smem, stream, g, dim, div, mask, perm, seg_of,
                    seg_start, n, stage_p, part, gsum, long_list, long_count, n_valid);
      }
    }
  } else if (mask) {
    MREC_LAUNCH((segsum_tiles_kernel<Vec, GT, true>), grid, kSegThreads, 0, stream, g, cpr, div, mask, perm,
                seg_of, seg_start, n, n_tiles, part, gsum, long_list, long_count, n_valid);
  } else {
    MREC_LAUNCH((segsum_tiles_kernel<Vec, GT, false>), grid, kSegThreads, 0, stream, g, cpr, div, mask, perm,
                seg_of, seg_start, n, n_tiles, part, gsum, long_list, long_count, n_valid);
  }
  const int groups = kSegThreads / cpr;
  int row_blocks = grid_for(cdiv(n, groups), 8);
  if constexpr (kStandalone) {
    if (n_tiles > 1) {
      StoreSink<Vec> store{gsum, cpr};
      MREC_LAUNCH((rows_update_kernel<Vec, StoreSink<Vec>, 1>), kChainBlocks + row_blocks, kSegThreads, 0, stream,
                  cpr, seg_of, seg_start, n, gsum, part, long_list, long_count, store, n_valid);
    }
  } else {
    // MREC_ROWS_R: rows in flight per thread of the row pass (1 default | 2).  Measured r1h (U = 243 k rows of
    // 320 B x {w, m, v}): 1 row x 8 CTAs/SM 0.170 ms, 2 rows x 3/SM 0.176, 4 rows x 2/SM 0.182 — the pass is
    // not bound by loads in flight but by DRAM efficiency on random 320-byte read-modify-write rows.
    const char* re = getenv("MREC_ROWS_R");
    const int rows_r = re ? atoi(re) : 1;
    const bool v4 = sizeof(Vec) == 16;
    const char* pse = getenv("MREC_ROWS_PER_SM");
    const int per_sm_env = pse ? atoi(pse) : 0;
#define MREC_ROWS(R, PER_SM)                                                                                     \
  do {                                                                                                           \
    row_blocks = (int)std::min<int64_t>(cdiv(n, (int64_t)groups * R), (int64_t)kNumSMs * (per_sm_env ? per_sm_env : PER_SM)); \
    if (row_blocks < 1) row_blocks = 1;                                                                          \
    MREC_LAUNCH((rows_update_kernel<Vec, Sink, R>), kChainBlocks + row_blocks, kSegThreads, 0, stream, cpr, seg_of, \
                seg_start, n, gsum, part, long_list, long_count, sink, n_valid);                                 \
  } while (0)
    bool done256 = false;
    if constexpr (IsLazyAdamF4<Sink>::value) {
      // MREC_ROWS_256 (default on): 32-byte accesses with L2 evict-first for the LazyAdam rows when D % 8 == 0
      const char* e256 = getenv("MREC_ROWS_256");
      const bool want = e256 ? atoi(e256) != 0 : true;
      const bool al = ((reinterpret_cast<uintptr_t>(sink.w) | reinterpret_cast<uintptr_t>(sink.m) |
                        reinterpret_cast<uintptr_t>(sink.v) | reinterpret_cast<uintptr_t>(gsum) |
                        reinterpret_cast<uintptr_t>(part)) % 32) == 0;
      if (want && al && dim % 8 == 0) {
        using S8 = LazyAdamSink<F8, typename IsLazyAdamF4<Sink>::Id>;
        const int cpr8 = dim / 8, groups8 = kSegThreads / cpr8;
        S8 s8{reinterpret_cast<F8*>(sink.w), reinterpret_cast<F8*>(sink.m), reinterpret_cast<F8*>(sink.v), sink.uniq,
              sink.hyper, sink.vocab, cpr8, sink.pitch / 2};
        int rb8 = (int)std::min<int64_t>(cdiv(n, (int64_t)groups8), (int64_t)kNumSMs * (per_sm_env ? per_sm_env : 8));
        if (rb8 < 1) rb8 = 1;
        MREC_LAUNCH((rows_update_kernel<F8, S8, 1>), kChainBlocks + rb8, kSegThreads, 0, stream, cpr8, seg_of, seg_start, n,
                    reinterpret_cast<const F8*>(gsum), reinterpret_cast<const F8*>(part), long_list, long_count, s8, n_valid);
        done256 = true;
      }
    }
    if (done256) {
    } else if (rows_r == 2) MREC_ROWS(2, (v4 ? 3 : 4));
    else MREC_ROWS(1, 8);
#undef MREC_ROWS
  }
  return check_launch("segment_sum");
}

// g is float32 unless g16 (float16 gradient rows)
template <typename Vec, typename Sink>
static int run_segsum(const void* g, bool g16, int dim, int div, const float* mask, const int32_t* perm,
                      const int32_t* seg_start, const int32_t* seg_of, int64_t n, void* ws, size_t ws_bytes,
                      Sink sink, cudaStream_t stream, Vec* out = nullptr, const int32_t* n_valid = nullptr) {
  if (g16)
    return run_segsum_t<Vec, __half, Sink>(reinterpret_cast<const __half*>(g), dim, div, mask, perm, seg_start,
                                           seg_of, n, ws, ws_bytes, sink, out, stream, n_valid);
  return run_segsum_t<Vec, float, Sink>(reinterpret_cast<const float*>(g), dim, div, mask, perm, seg_start,
                                        seg_of, n, ws, ws_bytes, sink, out, stream, n_valid);
}

// ---- dense optimizers (MLP parameters, Wide_b: SURVEY a5/a7) ----
__global__ void adam_begin_step_kernel(float* hyper) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float b1p = hyper[4] * hyper[1];
    const float b2p = hyper[5] * hyper[2];
    hyper[4] = b1p;
    hyper[5] = b2p;
    hyper[6] = hyper[0] * sqrtf(1.f - b2p) / (1.f - b1p);
  }
}

__global__ void __launch_bounds__(256)
adam_dense_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v,
                  const float* __restrict__ g, const float* __restrict__ hyper, int64_t n) {
  const float b1 = hyper[1], b2 = hyper[2], eps = hyper[3], lr_t = hyper[6], sc = hyper[7];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(g)) % 16 == 0)
                         ? n / 4 : 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 W = reinterpret_cast<float4*>(w)[i], M = reinterpret_cast<float4*>(m)[i],
           V = reinterpret_cast<float4*>(v)[i];
    const float4 G = reinterpret_cast<const float4*>(g)[i];
    adam_elem(W.x, M.x, V.x, G.x * sc, b1, b2, eps, lr_t);
    adam_elem(W.y, M.y, V.y, G.y * sc, b1, b2, eps, lr_t);
    adam_elem(W.z, M.z, V.z, G.z * sc, b1, b2, eps, lr_t);
    adam_elem(W.w, M.w, V.w, G.w * sc, b1, b2, eps, lr_t);
    reinterpret_cast<float4*>(w)[i] = W;
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
  }
  for (int64_t i = n4 * 4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float W = w[i], M = m[i], V = v[i];
    adam_elem(W, M, V, g[i] * sc, b1, b2, eps, lr_t);
    w[i] = W; m[i] = M; v[i] = V;
  }
}

__global__ void __launch_bounds__(256)
ftrl_dense_kernel(float* __restrict__ w, float* __restrict__ acc, float* __restrict__ lin,
                  const float* __restrict__ g, const float* __restrict__ hyper, int64_t n) {
  const float lr = hyper[0], l1 = hyper[1], l2 = hyper[2], p = hyper[3], sc = hyper[4];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float W = w[i], A = acc[i], L = lin[i];
    ftrl_elem(W, A, L, g[i] * sc, lr, l1, l2, p);
    w[i] = W; acc[i] = A; lin[i] = L;
  }
}


// ---- nn.Adam with a RowTensor gradient on GPU = dense-equivalent: every row's moments
// decay and every row moves; looked-up rows additionally receive their summed gradient.  Rows are
// flagged by the looked-up keys, the streaming kernel handles the unflagged rows (g = l2 * w), the fused
// segment-sum kernel handles the flagged ones, and the flags are cleared again.
template <typename IdT>
__global__ void mark_rows_kernel(const IdT* __restrict__ uniq, const int32_t* __restrict__ seg_of,
                                 int64_t n, int64_t vocab, uint8_t* __restrict__ flags, uint8_t value) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t row = (int64_t)uniq[seg_of[i]];
  if ((uint64_t)row < (uint64_t)vocab) flags[row] = value;
}

__global__ void __launch_bounds__(256)
adam_untouched_rows_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v,
                           const float* __restrict__ hyper, const uint8_t* __restrict__ flags,
                           int64_t vocab, int dim) {
  const float b1 = hyper[1], b2 = hyper[2], eps = hyper[3], lr_t = hyper[6], l2 = hyper[8];
  const int64_t total = vocab * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e / dim;
    if (flags[row]) continue;
    float W = w[e], M = m[e], V = v[e];
    adam_elem(W, M, V, l2 * W, b1, b2, eps, lr_t);
    w[e] = W; m[e] = M; v[e] = V;
  }
}

}  // namespace mrec

using namespace mrec;

MREC_API size_t mrec_sparse_opt_workspace_bytes(int64_t n, int dim) {
  return sparse_opt_workspace_bytes(n, dim);
}
MREC_API size_t mrec_segment_sum_workspace_bytes(int64_t n, int dim) {
  return segment_sum_workspace_bytes(n, dim);
}

// Shared validation of (g, mask, uniq, perm, seg_start, seg_of) starting at param index `b`.
struct SegArgs {
  const void* g; bool g16; const float* mask; const void* uniq; bool uniq64;
  const int32_t* perm; const int32_t* seg_start; const int32_t* seg_of;
  int64_t n; int div;
};
static int parse_seg_args(const Aot& a, int b, int dim, bool has_uniq, SegArgs* s, const char* who) {
  int i = b;
  MREC_REQUIRE(a.is_f32(i) || a.is(i, "float16"), ERR_DTYPE, "%s: grad rows must be float32|float16", who);
  s->g16 = a.is(i, "float16");
  MREC_REQUIRE(dim == 1 || a.last(i) == dim, ERR_SHAPE, "%s: grad rows must have last dim %d", who, dim);
  const int64_t g_rows = a.numel(i) / dim;
  s->g = a.params[i++];
  MREC_REQUIRE(a.is_f32(i), ERR_DTYPE, "%s: mask must be float32 (numel 0 = no mask)", who);
  const int64_t mask_n = a.numel(i);
  s->mask = mask_n ? a.ptr<float>(i) : nullptr;
  i++;
  if (has_uniq) {
    MREC_REQUIRE(a.is_i32(i) || a.is_i64(i), ERR_DTYPE, "%s: uniq must be int32|int64", who);
    s->uniq64 = a.is_i64(i);
    s->uniq = a.params[i++];
  }
  MREC_REQUIRE(a.is_i32(i) && a.is_i32(i + 1) && a.is_i32(i + 2), ERR_DTYPE,
               "%s: perm/seg_start/seg_of must be int32", who);
  s->n = a.numel(i);
  s->perm = a.ptr<int32_t>(i++);
  MREC_REQUIRE(a.numel(i) >= s->n + 1, ERR_SHAPE, "%s: seg_start must have N+1 entries", who);
  s->seg_start = a.ptr<int32_t>(i++);
  MREC_REQUIRE(a.numel(i) >= s->n, ERR_SHAPE, "%s: seg_of must have N entries", who);
  s->seg_of = a.ptr<int32_t>(i++);
  MREC_REQUIRE(mask_n == 0 || mask_n == s->n, ERR_SHAPE, "%s: mask numel must be 0 or N", who);
  MREC_REQUIRE(g_rows > 0 ? (s->n % g_rows == 0) : s->n == 0, ERR_SHAPE,
               "%s: N (%lld) must be a multiple of the number of grad rows (%lld)", who, (long long)s->n,
               (long long)g_rows);
  s->div = g_rows > 0 ? (int)(s->n / g_rows) : 1;
  if (dim % 4 == 0)
    MREC_REQUIRE(reinterpret_cast<uintptr_t>(s->g) % (s->g16 ? 8 : 16) == 0, ERR_ALIGN,
                 "%s: grad rows must be 16-B aligned", who);
  return OK;
}

// inputs : w[V,D] m[V,D] v[V,D] hyper[8] g[N/div,D] mask[N|0] uniq[N] perm[N] seg_start[N+1] seg_of[N]
// outputs: dummy[1] i32, workspace[bytes]
MREC_API int mrec_sparse_lazy_adam(int nparam, void** params, int* ndims, int64_t** shapes,
                                   const char** dtypes, void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  // optional input 10: n_valid[1] i32 (device-side count of real sorted positions; outputs shift by one)
  const bool has_nv = (a.nparam == 13);
  if (a.nparam != 12 && a.nparam != 13)
    return fail(ERR_NPARAM, "mrec_sparse_lazy_adam: expected 12 or 13 params, got %d", a.nparam);
  for (int i = 0; i < a.nparam; ++i)
    if (!a.params[i] && a.numel(i) > 0) return fail(ERR_NULL, "mrec_sparse_lazy_adam: param %d is null", i);
  const int32_t* n_valid = nullptr;
  if (has_nv) {
    MREC_REQUIRE(a.is_i32(10) && a.numel(10) >= 1, ERR_DTYPE, "mrec_sparse_lazy_adam: n_valid must be int32[1]");
    n_valid = a.ptr<int32_t>(10);
  }
  const int ws_i = has_nv ? 12 : 11;
  MREC_REQUIRE(a.is_f32(0) && a.is_f32(1) && a.is_f32(2) && a.is_f32(3), ERR_DTYPE,
               "mrec_sparse_lazy_adam: w/m/v/hyper must be float32");
  MREC_REQUIRE(a.numel(3) >= kHyperLen, ERR_SHAPE, "mrec_sparse_lazy_adam: hyper needs %d floats", kHyperLen);
  const int64_t vocab = a.dim(0, 0);
  // interleaved layout: ONE array wmv[V,3,D] holds each row's w | m | v back to back (a 3*D*4-byte record: one DRAM
  // burst per row update instead of three); m and v are then passed with numel 0
  const bool packed = a.ndims[0] == 3 && a.dim(0, 1) == 3 && a.numel(1) == 0 && a.numel(2) == 0;
  const int dim = packed ? (int)a.dim(0, 2) : (a.ndims[0] >= 2 ? (int)a.dim(0, 1) : 1);
  MREC_REQUIRE(packed || (a.numel(1) == a.numel(0) && a.numel(2) == a.numel(0)), ERR_SHAPE,
               "mrec_sparse_lazy_adam: m/v shape must equal w shape (or w = wmv[V,3,D] with empty m, v)");
  MREC_REQUIRE(!packed || dim % 4 == 0, ERR_DIM, "mrec_sparse_lazy_adam: the interleaved layout needs D % 4 == 0");
  float* const w_p = a.ptr<float>(0);
  float* const m_p = packed ? w_p + dim : a.ptr<float>(1);
  float* const v_p = packed ? w_p + 2 * dim : a.ptr<float>(2);
  const int64_t rows_pitch = packed ? 3 : 1;      // in units of one row of D floats
  SegArgs s;
  int rc = parse_seg_args(a, 4, dim, true, &s, "mrec_sparse_lazy_adam");
  if (rc) return rc;
  const size_t ws_bytes = (size_t)a.numel(ws_i);
  void* ws = a.params[ws_i];
  const float* hyper = a.ptr<float>(3);
  if (dim % 4 == 0) {
    MREC_REQUIRE(a.aligned(0, 16) && (packed || (a.aligned(1, 16) && a.aligned(2, 16))), ERR_ALIGN,
                 "mrec_sparse_lazy_adam: w/m/v must be 16-byte aligned");
    const int cpr = dim / 4;
    MREC_REQUIRE(cpr <= kSegThreads, ERR_DIM, "mrec_sparse_lazy_adam: D too large");
    if (s.uniq64) {
      LazyAdamSink<float4, int64_t> sink{(float4*)w_p, (float4*)m_p, (float4*)v_p,
                                         (const int64_t*)s.uniq, hyper, vocab, cpr, rows_pitch * cpr};
      return run_segsum<float4>(s.g, s.g16, dim, s.div, s.mask, s.perm, s.seg_start, s.seg_of, s.n, ws, ws_bytes, sink, a.stream, nullptr, n_valid);
    }
    LazyAdamSink<float4, int32_t> sink{(float4*)w_p, (float4*)m_p, (float4*)v_p,
                                       (const int32_t*)s.uniq, hyper, vocab, cpr, rows_pitch * cpr};
    return run_segsum<float4>(s.g, s.g16, dim, s.div, s.mask, s.perm, s.seg_start, s.seg_of, s.n, ws, ws_bytes, sink, a.stream, nullptr, n_valid);
  }
  MREC_REQUIRE(dim <= kSegThreads, ERR_DIM, "mrec_sparse_lazy_adam: D too large");
  if (s.uniq64) {
    LazyAdamSink<float, int64_t> sink{w_p, m_p, v_p,
                                      (const int64_t*)s.uniq, hyper, vocab, dim, (int64_t)dim};
    return run_segsum<float>(s.g, s.g16, dim, s.div, s.mask, s.perm, s.seg_start, s.seg_of, s.n, ws, ws_bytes, sink, a.stream, nullptr, n_valid);
  }
  LazyAdamSink<float, int32_t> sink{w_p, m_p, v_p,
                                    (const int32_t*)s.uniq, hyper, vocab, dim, (int64_t)dim};
  return run_segsum<float>(s.g, s.g16, dim, s.div, s.mask, s.perm, s.seg_start, s.seg_of, s.n, ws, ws_bytes, sink, a.stream, nullptr, n_valid);
}

// inputs : w[V,D] accum[V,D] linear[V,D] hyper[8] g[N/div,D] mask[N|0] uniq[N] perm[N] seg_start[N+1] seg_of[N]
// outputs: dummy[1] i32, workspace[bytes]
MREC_API int mrec_sparse_ftrl(int nparam, void** params, int* ndims, int64_t** shapes,
                              const char** dtypes, void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  // optional input 10: n_valid[1] i32 (device-side count of real sorted positions; outputs shift by one)
  const bool has_nv = (a.nparam == 13);
  if (a.nparam != 12 && a.nparam != 13)
    return fail(ERR_NPARAM, "mrec_sparse_ftrl: expected 12 or 13 params, got %d", a.nparam);
  for (int i = 0; i < a.nparam; ++i)
    if (!a.params[i] && a.numel(i) > 0) return fail(ERR_NULL, "mrec_sparse_ftrl: param %d is null", i);
  const int32_t* n_valid = nullptr;
  if (has_nv) {
    MREC_REQUIRE(a.is_i32(10) && a.numel(10) >= 1, ERR_DTYPE, "mrec_sparse_ftrl: n_valid must be int32[1]");
    n_valid = a.ptr<int32_t>(10);
  }
  const int ws_i = has_nv ? 12 : 11;
  MREC_REQUIRE(a.is_f32(0) && a.is_f32(1) && a.is_f32(2) && a.is_f32(3), ERR_DTYPE,
               "mrec_sparse_ftrl: w/accum/linear/hyper must be float32");
  MREC_REQUIRE(a.numel(3) >= kHyperLen, ERR_SHAPE, "mrec_sparse_ftrl: hyper needs %d floats", kHyperLen);
  const int64_t vocab = a.dim(0, 0);
  const int dim = a.ndims[0] >= 2 ? (int)a.dim(0, 1) : 1;
  MREC_REQUIRE(a.numel(1) == a.numel(0) && a.numel(2) == a.numel(0), ERR_SHAPE,
               "mrec_sparse_ftrl: accum/linear shape must equal w shape");
  SegArgs s;
  int rc = parse_seg_args(a, 4, dim, true, &s, "mrec_sparse_ftrl");
  if (rc) return rc;
  const size_t ws_bytes = (size_t)a.numel(ws_i);
  void* ws = a.params[ws_i];
  const float* hyper = a.ptr<float>(3);
  if (dim % 4 == 0) {
    MREC_REQUIRE(a.aligned(0, 16) && a.aligned(1, 16) && a.aligned(2, 16), ERR_ALIGN,
                 "mrec_sparse_ftrl: w/accum/linear must be 16-byte aligned");
    const int cpr = dim / 4;
    MREC_REQUIRE(cpr <= kSegThreads, ERR_DIM, "mrec_sparse_ftrl: D too large");
    if (s.uniq64) {
      FtrlSink<float4, int64_t> sink{a.ptr<float4>(0), a.ptr<float4>(1), a.ptr<float4>(2),
                                     (const int64_t*)s.uniq, hyper, vocab, cpr};
      return run_segsum<float4>(s.g, s.g16, dim, s.div, s.mask, s.perm, s.seg_start, s.seg_of, s.n, ws, ws_bytes, sink, a.stream, nullptr, n_valid);
    }
    FtrlSink<float4, int32_t> sink{a.ptr<float4>(0), a.ptr<float4>(1), a.ptr<float4>(2),
                                   (const int32_t*)s.uniq, hyper, vocab, cpr};
    return run_segsum<float4>(s.g, s.g16, dim, s.div, s.mask, s.perm, s.seg_start, s.seg_of, s.n, ws, ws_bytes, sink, a.stream, nullptr, n_valid);
  }
  MREC_REQUIRE(dim <= kSegThreads, ERR_DIM, "mrec_sparse_ftrl: D too large");
  if (s.uniq64) {
    FtrlSink<float, int64_t> sink{a.ptr<float>(0), a.ptr<float>(1), a.ptr<float>(2),
                                  (const int64_t*)s.uniq, hyper, vocab, dim};
    return run_segsum<float>(s.g, s.g16, dim, s.div, s.mask, s.perm, s.seg_start, s.seg_of, s.n, ws, ws_bytes, sink, a.stream, nullptr, n_valid);
  }
  FtrlSink<float, int32_t> sink{a.ptr<float>(0), a.ptr<float>(1), a.ptr<float>(2),
                                (const int32_t*)s.uniq, hyper, vocab, dim};
  return run_segsum<float>(s.g, s.g16, dim, s.div, s.mask, s.perm, s.seg_start, s.seg_of, s.n, ws, ws_bytes, sink, a.stream, nullptr, n_valid);
}

// Standalone deterministic segment-sum (the UnsortedSegmentSum of SURVEY a4, in sorted-segment order).
// inputs : g[N/div,D] mask[N|0] perm[N] seg_start[N+1] seg_of[N] ; outputs: gsum[N,D] (rows >= U untouched),
//          workspace[bytes]
MREC_API int mrec_segment_sum(int nparam, void** params, int* ndims, int64_t** shapes,
                              const char** dtypes, void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  MREC_CHECK_NPARAM(a, 7);
  MREC_REQUIRE(a.is_f32(5), ERR_DTYPE, "mrec_segment_sum: gsum must be float32");
  const int dim = a.ndims[5] >= 2 ? (int)a.last(5) : 1;
  SegArgs s;
  int rc = parse_seg_args(a, 0, dim, false, &s, "mrec_segment_sum");
  if (rc) return rc;
  MREC_REQUIRE(a.numel(5) >= s.n * dim, ERR_SHAPE, "mrec_segment_sum: gsum must be padded to [N,D]");
  const size_t ws_bytes = (size_t)a.numel(6);
  if (dim % 4 == 0) {
    MREC_REQUIRE(a.aligned(5, 16), ERR_ALIGN, "mrec_segment_sum: gsum must be 16-byte aligned");
    MREC_REQUIRE(dim / 4 <= kSegThreads, ERR_DIM, "mrec_segment_sum: D too large");
    return run_segsum<float4>(s.g, s.g16, dim, s.div, s.mask, s.perm, s.seg_start, s.seg_of, s.n, a.params[6],
                              ws_bytes, NoSink{}, a.stream, a.ptr<float4>(5));
  }
  MREC_REQUIRE(dim <= kSegThreads, ERR_DIM, "mrec_segment_sum: D too large");
  return run_segsum<float>(s.g, s.g16, dim, s.div, s.mask, s.perm, s.seg_start, s.seg_of, s.n, a.params[6],
                           ws_bytes, NoSink{}, a.stream, a.ptr<float>(5));
}


// Dense gradient of a (non-sparse) Gather, accumulated: table[uniq[u], :] += sum of segment u.
// inputs : g[N/div,D] mask[N|0] uniq[N] perm[N] seg_start[N+1] seg_of[N]
// outputs: table[V,D] f32 (accumulated in place), workspace[mrec_sparse_opt_workspace_bytes]
MREC_API int mrec_segment_sum_scatter_add(int nparam, void** params, int* ndims, int64_t** shapes,
                                          const char** dtypes, void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  MREC_CHECK_NPARAM(a, 8);
  MREC_REQUIRE(a.is_f32(6), ERR_DTYPE, "mrec_segment_sum_scatter_add: table must be float32");
  const int64_t vocab = a.dim(6, 0);
  const int dim = a.ndims[6] >= 2 ? (int)a.dim(6, 1) : 1;
  SegArgs s;
  int rc = parse_seg_args(a, 0, dim, true, &s, "mrec_segment_sum_scatter_add");
  if (rc) return rc;
  const size_t ws_bytes = (size_t)a.numel(7);
  void* ws = a.params[7];
  if (s.n == 0 || vocab == 0) return OK;
  if (dim % 4 == 0) {
    MREC_REQUIRE(a.aligned(6, 16), ERR_ALIGN, "mrec_segment_sum_scatter_add: table must be 16-byte aligned");
    MREC_REQUIRE(dim / 4 <= kSegThreads, ERR_DIM, "mrec_segment_sum_scatter_add: D too large");
    const int cpr = dim / 4;
    if (s.uniq64) {
      ScatterAddSink<float4, int64_t> sink{a.ptr<float4>(6), reinterpret_cast<const int64_t*>(s.uniq), vocab, cpr};
      return run_segsum<float4>(s.g, s.g16, dim, s.div, s.mask, s.perm, s.seg_start, s.seg_of, s.n, ws, ws_bytes, sink, a.stream);
    }
    ScatterAddSink<float4, int32_t> sink{a.ptr<float4>(6), reinterpret_cast<const int32_t*>(s.uniq), vocab, cpr};
    return run_segsum<float4>(s.g, s.g16, dim, s.div, s.mask, s.perm, s.seg_start, s.seg_of, s.n, ws, ws_bytes, sink, a.stream);
  }
  MREC_REQUIRE(dim <= kSegThreads, ERR_DIM, "mrec_segment_sum_scatter_add: D too large");
  if (s.uniq64) {
    ScatterAddSink<float, int64_t> sink{a.ptr<float>(6), reinterpret_cast<const int64_t*>(s.uniq), vocab, dim};
    return run_segsum<float>(s.g, s.g16, dim, s.div, s.mask, s.perm, s.seg_start, s.seg_of, s.n, ws, ws_bytes, sink, a.stream);
  }
  ScatterAddSink<float, int32_t> sink{a.ptr<float>(6), reinterpret_cast<const int32_t*>(s.uniq), vocab, dim};
  return run_segsum<float>(s.g, s.g16, dim, s.div, s.mask, s.perm, s.seg_start, s.seg_of, s.n, ws, ws_bytes, sink, a.stream);
}


// nn.Adam (not Lazy) with a RowTensor gradient: dense-equivalent update of the whole table.
// inputs : w m v hyper[16] g mask uniq perm seg_start seg_of row_flags[V] u8 (all zero on entry and exit)
// outputs: dummy[1] i32, workspace[mrec_sparse_opt_workspace_bytes]
MREC_API int mrec_adam_rowsparse(int nparam, void** params, int* ndims, int64_t** shapes,
                                 const char** dtypes, void* stream, void* extra) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  MREC_CHECK_NPARAM(a, 13);
  MREC_REQUIRE(a.is(10, "uint8") || a.is(10, "int8") || a.is(10, "bool"), ERR_DTYPE,
               "mrec_adam_rowsparse: row_flags must be a 1-byte type");
  const int64_t vocab = a.dim(0, 0);
  const int dim = a.ndims[0] >= 2 ? (int)a.dim(0, 1) : 1;
  MREC_REQUIRE(a.numel(10) >= vocab, ERR_SHAPE, "mrec_adam_rowsparse: row_flags needs V entries");
  MREC_REQUIRE(a.is_i32(6) || a.is_i64(6), ERR_DTYPE, "mrec_adam_rowsparse: uniq must be int32|int64");
  MREC_REQUIRE(a.is_i32(9), ERR_DTYPE, "mrec_adam_rowsparse: seg_of must be int32");
  const int64_t n = a.numel(7);
  uint8_t* flags = a.ptr<uint8_t>(10);
  const int grid_n = (int)cdiv(n > 0 ? n : 1, 256);
  if (n > 0) {
    if (a.is_i64(6)) MREC_LAUNCH(mark_rows_kernel<int64_t>, grid_n, 256, 0, a.stream, a.ptr<int64_t>(6), a.ptr<int32_t>(9), n, vocab, flags, (uint8_t)1);
    else MREC_LAUNCH(mark_rows_kernel<int32_t>, grid_n, 256, 0, a.stream, a.ptr<int32_t>(6), a.ptr<int32_t>(9), n, vocab, flags, (uint8_t)1);
  }
  MREC_REQUIRE(a.is_f32(0) && a.is_f32(1) && a.is_f32(2) && a.is_f32(3) && a.numel(3) >= kHyperLen, ERR_DTYPE,
               "mrec_adam_rowsparse: w/m/v/hyper must be float32 (hyper[16])");
  MREC_LAUNCH(adam_untouched_rows_kernel, grid_for(cdiv(vocab * dim, 1024), 8), 256, 0, a.stream,
              a.ptr<float>(0), a.ptr<float>(1), a.ptr<float>(2), a.ptr<float>(3), flags, vocab, dim);
  // touched rows: the LazyAdam path (same formula, g = gsum/scale + l2*w)
  void* p2[12]; int nd2[12]; int64_t* sh2[12]; const char* dt2[12];
  const int map[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 12};
  for (int i = 0; i < 12; ++i) { p2[i] = params[map[i]]; nd2[i] = ndims[map[i]]; sh2[i] = shapes[map[i]]; dt2[i] = dtypes[map[i]]; }
  int rc = mrec_sparse_lazy_adam(12, p2, nd2, sh2, dt2, stream, extra);
  if (rc) return rc;
  if (n > 0) {
    if (a.is_i64(6)) MREC_LAUNCH(mark_rows_kernel<int64_t>, grid_n, 256, 0, a.stream, a.ptr<int64_t>(6), a.ptr<int32_t>(9), n, vocab, flags, (uint8_t)0);
    else MREC_LAUNCH(mark_rows_kernel<int32_t>, grid_n, 256, 0, a.stream, a.ptr<int32_t>(6), a.ptr<int32_t>(9), n, vocab, flags, (uint8_t)0);
  }
  return check_launch("adam_rowsparse");
}

// inputs: hyper[16] ; outputs: dummy[1].  beta powers advance, lr_t = lr*sqrt(1-b2^t)/(1-b1^t).
MREC_API int mrec_adam_begin_step(int nparam, void** params, int* ndims, int64_t** shapes,
                                  const char** dtypes, void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  MREC_CHECK_NPARAM(a, 2);
  MREC_REQUIRE(a.is_f32(0) && a.numel(0) >= kHyperLen, ERR_SHAPE, "mrec_adam_begin_step: hyper must be f32[16]");
  MREC_LAUNCH(adam_begin_step_kernel, 1, 32, 0, a.stream, a.ptr<float>(0));
  return check_launch("adam_begin_step");
}

// inputs: w m v hyper[8] g (all same numel) ; outputs: dummy[1]
MREC_API int mrec_adam_dense(int nparam, void** params, int* ndims, int64_t** shapes,
                             const char** dtypes, void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  MREC_CHECK_NPARAM(a, 6);
  for (int i = 0; i < 5; ++i) MREC_REQUIRE(a.is_f32(i), ERR_DTYPE, "mrec_adam_dense: param %d must be float32", i);
  const int64_t n = a.numel(0);
  MREC_REQUIRE(a.numel(1) == n && a.numel(2) == n && a.numel(4) == n && a.numel(3) >= kHyperLen, ERR_SHAPE,
               "mrec_adam_dense: w/m/v/g numel mismatch");
  if (n == 0) return OK;
  MREC_LAUNCH(adam_dense_kernel, grid_for(cdiv(n, 1024), 8), 256, 0, a.stream, a.ptr<float>(0),
              a.ptr<float>(1), a.ptr<float>(2), a.ptr<float>(4), a.ptr<float>(3), n);
  return check_launch("adam_dense");
}

// inputs: w accum linear hyper[8] g ; outputs: dummy[1]
MREC_API int mrec_ftrl_dense(int nparam, void** params, int* ndims, int64_t** shapes,
                             const char** dtypes, void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  MREC_CHECK_NPARAM(a, 6);
  for (int i = 0; i < 5; ++i) MREC_REQUIRE(a.is_f32(i), ERR_DTYPE, "mrec_ftrl_dense: param %d must be float32", i);
  const int64_t n = a.numel(0);
  MREC_REQUIRE(a.numel(1) == n && a.numel(2) == n && a.numel(4) == n && a.numel(3) >= kHyperLen, ERR_SHAPE,
               "mrec_ftrl_dense: w/accum/linear/g numel mismatch");
  if (n == 0) return OK;
  MREC_LAUNCH(ftrl_dense_kernel, grid_for(cdiv(n, 256), 8), 256, 0, a.stream, a.ptr<float>(0),
              a.ptr<float>(1), a.ptr<float>(2), a.ptr<float>(4), a.ptr<float>(3), n);
  return check_launch("ftrl_dense");
}

// Fused UnsortedSegmentSum + gradient push (backward all-to-all of row-sharded tables): segment u of the local dedup
// is summed and stored directly into the inbox of the rank that owns its key (see PeerDest).
// inputs : g[N/div,D] f32|f16 (D % 4 == 0), mask[N|0], perm[N], seg_start[N+1], seg_of[N], my_bounds[G+1] i32,
//          inbox_off[G] i32, peer_ptrs[G] i64, cap_like[cap_rows, ..]
// outputs: err[1] i32 (bit 1: an inbox overflowed), workspace[mrec_segment_sum_workspace_bytes(N, D)]
MREC_API int mrec_segment_sum_to_peers(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                                       void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  MREC_CHECK_NPARAM(a, 11);
  const int dim = (int)a.last(0);
  MREC_REQUIRE(dim >= 4 && dim % 4 == 0 && dim / 4 <= kSegThreads, ERR_DIM, "mrec_segment_sum_to_peers: D must be a multiple of 4");
  SegArgs s;
  int rc = parse_seg_args(a, 0, dim, false, &s, "mrec_segment_sum_to_peers");
  if (rc) return rc;
  MREC_REQUIRE(a.is_i32(5) && a.is_i32(6) && a.is_i64(7) && a.is_i32(9), ERR_DTYPE,
               "mrec_segment_sum_to_peers: bounds/inbox_off/err int32, peer_ptrs int64");
  const int world = (int)a.numel(6);
  MREC_REQUIRE(world >= 1 && a.numel(5) == world + 1 && a.numel(7) == world && a.numel(9) >= 1, ERR_SHAPE,
               "mrec_segment_sum_to_peers: bounds[G+1], inbox_off[G], peer_ptrs[G], err[1]");
  const int64_t n = s.n;
  if (n == 0) return OK;
  MREC_REQUIRE(n < ((int64_t)1 << 31), ERR_SHAPE, "mrec_segment_sum_to_peers: N must be < 2^31");
  const SegWorkspace W = seg_ws(n, dim, false);
  MREC_REQUIRE((size_t)a.numel(10) >= W.total, ERR_WORKSPACE, "mrec_segment_sum_to_peers: workspace %lld < %zu",
               (long long)a.numel(10), W.total);
  MREC_REQUIRE(a.aligned(10, 16), ERR_ALIGN, "mrec_segment_sum_to_peers: workspace must be 16-byte aligned");
  char* w = a.ptr<char>(10);
  float4* part = reinterpret_cast<float4*>(w + W.off_part);
  int32_t* long_list = reinterpret_cast<int32_t*>(w + W.off_list);
  int32_t* long_count = reinterpret_cast<int32_t*>(w + W.off_count);
  const PeerDest pd{a.ptr<int32_t>(5), a.ptr<int32_t>(6), a.ptr<int64_t>(7), world, a.dim(8, 0), a.ptr<int32_t>(9)};
  const int row_bytes = dim * (s.g16 ? 2 : 4);
  int stage_p = 128;
  while (stage_p > 32 && (size_t)stage_p * (row_bytes + 12) > 40 * 1024) stage_p >>= 1;
  MREC_REQUIRE((size_t)stage_p * (row_bytes + 12) <= 40 * 1024, ERR_DIM, "mrec_segment_sum_to_peers: D too large");
  const size_t smem = (size_t)stage_p * (row_bytes + 12);
  const int grid_s = (int)cdiv(n, stage_p);
  cudaMemsetAsync(long_count, 0, sizeof(int32_t), a.stream);
#define MREC_STAGE_PEER(GT, HM)                                                                                        \
  MREC_LAUNCH((segsum_stage_kernel<GT, HM, true>), grid_s, kSegThreads, smem, a.stream, reinterpret_cast<const GT*>(s.g), \
              dim, s.div, s.mask, s.perm, s.seg_of, s.seg_start, n, stage_p, part, (float4*)nullptr, long_list, long_count, \
              (const int32_t*)nullptr, pd)
  if (s.g16) { if (s.mask) MREC_STAGE_PEER(__half, true); else MREC_STAGE_PEER(__half, false); }
  else { if (s.mask) MREC_STAGE_PEER(float, true); else MREC_STAGE_PEER(float, false); }
#undef MREC_STAGE_PEER
  const int cpr = dim / 4;
  const int groups = kSegThreads / cpr;
  if (cdiv(n, kSegTile) > 1) {
    PeerStoreSink<float4> sink{pd, cpr};
    const int row_blocks = grid_for(cdiv(n, groups), 8);
    MREC_LAUNCH((rows_update_kernel<float4, PeerStoreSink<float4>, 1>), kChainBlocks + row_blocks, kSegThreads, 0, a.stream,
                cpr, s.seg_of, s.seg_start, n, (const float4*)nullptr, part, long_list, long_count, sink, (const int32_t*)nullptr);
  }
  return check_launch("segment_sum_to_peers");
}
