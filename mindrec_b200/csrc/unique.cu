// K2 — sparse-gradient / lookup dedup: stable LSD radix sort of (key, position) + unique.
//
// Replaces the upstream Unique kernel (thrust stable_sort_by_key + unique + scan) reached from
//   mindspore_rec/ops/embedding.py:191-194            (HashEmbeddingLookup forward dedup)
//   models/wide_deep/src/wide_and_deep.py:420-445     (optimizer-side RowTensor dedup)
// and produces, in one call, everything the deterministic segment-sum needs:
//   uniq[U]      ascending unique keys (GPU Unique order)
//   inverse[N]   uniq[inverse[i]] == ids[i]
//   count[1]     U (device scalar: aot outputs are statically shaped, so uniq is padded to N)
//   perm[N]      stable sort permutation (sorted position -> original position)
//   seg_start[N+1]  first sorted position of each segment, seg_start[U] = N
//   seg_of[N]    segment id of each sorted position
//
// The sort is an LSD radix sort over only the significant key bits with digits of up to 9 bits (a table
// with V rows needs ceil(log2(V+1)) bits: 26 bits -> 3 passes of 9 bits for the 33.8 M-row Criteo table).  Each pass is
// histogram (pass 0 only: later passes get theirs from the previous scatter) -> column scan -> stable scatter; ranking inside a tile uses warp match-any so equal digits
// keep their input order (stability is what makes perm's first entry of a segment the first occurrence).
// All buffers (ping-pong keys/values, histograms) come from the caller-provided workspace: the library
// never allocates (aot contract).
#include "common.cuh"
#include "kernels.h"
#include <algorithm>
#include <cstdlib>

namespace mrec {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 2048
constexpr int MAX_RADIX_BITS = 9;             // digit width is chosen per sort: ceil(bits / passes) <= 9
constexpr int MAX_RADIX = 1 << MAX_RADIX_BITS;  // shared-memory tables are sized for 512 bins
// Accumulating the next pass's histogram with global atomics inside the scatter was measured (r1e) to cost as
// much as the separate histogram kernel it saves; kept as a switch.
constexpr bool kFuseNextHist = false;

template <typename KeyT> struct UKeyOf;
template <> struct UKeyOf<int32_t> { using type = uint32_t; static constexpr uint32_t sign = 0x80000000u; };
template <> struct UKeyOf<int64_t> { using type = uint64_t; static constexpr uint64_t sign = 0x8000000000000000ull; };

// bound > 0: keys outside [0, bound) collapse onto `bound` (they sort last and form one segment that
// downstream kernels skip); bound == 0: full-width signed order (sign bit flipped).
template <typename KeyT>
__device__ __forceinline__ typename UKeyOf<KeyT>::type encode_key(KeyT k, uint64_t bound) {
  using U = typename UKeyOf<KeyT>::type;
  if (bound) return ((uint64_t)k < bound) ? (U)k : (U)bound;
  return (U)k ^ UKeyOf<KeyT>::sign;
}
template <typename KeyT>
__device__ __forceinline__ KeyT decode_key(typename UKeyOf<KeyT>::type u, uint64_t bound) {
  if (bound) return (KeyT)u;
  return (KeyT)(u ^ UKeyOf<KeyT>::sign);
}

// n_valid (optional): the input is a statically sized buffer of which only the first *n_valid entries are
// real (device-side count).  Every kernel of the sort then works on n_eff = min(n, *n_valid): CTAs past the valid
// prefix exit at once, so the cost follows the valid count, not the capacity; outputs past n_eff are not written.
__device__ __forceinline__ int64_t eff_n(int64_t n, const int32_t* n_valid) {
  if (!n_valid) return n;
  const int64_t v = n_valid[0];
  return v < 0 ? 0 : (v < n ? v : n);
}

template <typename KeyT, bool RAW>
__device__ __forceinline__ typename UKeyOf<KeyT>::type load_key(const void* keys, int64_t i, uint64_t bound) {
  using U = typename UKeyOf<KeyT>::type;
  if (RAW) {
    return encode_key<KeyT>(reinterpret_cast<const KeyT*>(keys)[i], bound);
  }
  return reinterpret_cast<const U*>(keys)[i];
}

// ---- pass kernel 1: per-block digit histogram -> hist[block][digit] ----
template <typename KeyT, bool RAW>
__global__ void __launch_bounds__(RS_THREADS)
radix_hist_kernel(const void* __restrict__ keys, int64_t n, int shift, int radix, uint64_t bound,
                  int tiles_per_block, uint32_t* __restrict__ hist, const int32_t* __restrict__ n_valid) {
  __shared__ uint32_t s_hist[MAX_RADIX];
  const int RADIX = radix;
  n = eff_n(n, n_valid);
  for (int d = threadIdx.x; d < RADIX; d += RS_THREADS) s_hist[d] = 0;
  __syncthreads();
  const int64_t begin = (int64_t)blockIdx.x * tiles_per_block * RS_TILE;
  const int64_t end = min(n, begin + (int64_t)tiles_per_block * RS_TILE);
  for (int64_t i = begin + threadIdx.x; i < end; i += RS_THREADS) {
    const auto k = load_key<KeyT, RAW>(keys, i, bound);
    atomicAdd(&s_hist[(uint32_t)(k >> shift) & (RADIX - 1)], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < RADIX; d += RS_THREADS)
    hist[(int64_t)blockIdx.x * RADIX + d] = s_hist[d];
}

// ---- pass kernel 2: hist[b][d] -> exclusive prefix over blocks (per digit), digit totals ----
// One CTA (32 warps) per group of 32 digits.  The [blocks x 32 digits] slab is walked in tiles of 32 blocks:
// warp w reads row b0+w (32 consecutive digits = one 128-byte line, coalesced), the tile is transposed through
// padded shared memory, warp w scans digit column w across the 32 blocks with shuffles, adds the running
// carry, and the tile is written back row-wise (coalesced).  (The first version let each warp walk a digit
// column directly: 32 cache lines per warp load, 21 us per pass on 16 SMs — ncu r1e.)  The cross-digit base
// is a <= 512-entry scan that every scatter block redoes in shared memory from the totals row.
__global__ void __launch_bounds__(1024)
radix_scan_kernel(uint32_t* __restrict__ hist, int n_blocks, int radix) {
  __shared__ uint32_t s_tile[32][33];
  __shared__ uint32_t s_carry[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d0 = blockIdx.x * 32;             // first digit of this CTA's group
  const bool d_ok = (d0 + lane) < radix;
  if (warp == 0) s_carry[lane] = 0;
  uint32_t next = 0;
  if (warp < n_blocks && d_ok) next = hist[(int64_t)warp * radix + d0 + lane];
  for (int b0 = 0; b0 < n_blocks; b0 += 32) {
    const int b = b0 + warp;
    const uint32_t cur = next;
    // prefetch the next tile's row while this one is processed
    next = 0;
    if (b + 32 < n_blocks && d_ok) next = hist[(int64_t)(b + 32) * radix + d0 + lane];
    s_tile[warp][lane] = (b < n_blocks) ? cur : 0u;
    __syncthreads();
    // warp w owns digit d0+w: lane = block index inside the tile
    const uint32_t c = s_tile[lane][warp];
    uint32_t v = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    const uint32_t carry = s_carry[warp];
    __syncthreads();                            // everyone has read s_tile / s_carry
    s_tile[lane][warp] = carry + v - c;         // exclusive prefix over blocks
    if (lane == 31) s_carry[warp] = carry + v;
    __syncthreads();
    if (b < n_blocks && d_ok) hist[(int64_t)b * radix + d0 + lane] = s_tile[warp][lane];
    __syncthreads();
  }
  if (warp == 0 && d_ok) hist[(int64_t)n_blocks * radix + d0 + lane] = s_carry[lane];  // digit totals
}

// ---- pass kernel 3: stable scatter ----
template <typename KeyT, bool RAW>
__global__ void __launch_bounds__(RS_THREADS)
radix_scatter_kernel(const void* __restrict__ keys_in, const int32_t* __restrict__ vals_in,
                     typename UKeyOf<KeyT>::type* __restrict__ keys_out, int32_t* __restrict__ vals_out,
                     int64_t n, int shift, int radix, uint64_t bound, int tiles_per_block,
                     const uint32_t* __restrict__ hist, int n_blocks, uint32_t* __restrict__ hist_next,
                     int next_shift, const int32_t* __restrict__ n_valid) {
  using U = typename UKeyOf<KeyT>::type;
  __shared__ uint32_t s_cnt[RS_WARPS][MAX_RADIX];
  __shared__ uint32_t s_base[MAX_RADIX];
  __shared__ uint32_t s_wsum[RS_WARPS];
  const int RADIX = radix;
  const int rbits = __ffs(radix) - 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  n = eff_n(n, n_valid);
  if ((int64_t)blockIdx.x * tiles_per_block * RS_TILE >= n) return;
  const uint32_t lt_mask = (1u << lane) - 1u;
  {
    // exclusive scan of the digit totals (<= 512 entries, 2 per thread) -> digit base
    const uint32_t* tot = hist + (int64_t)n_blocks * RADIX;
    const int d0 = 2 * tid, d1 = 2 * tid + 1;
    const uint32_t t0 = d0 < RADIX ? tot[d0] : 0u, t1 = d1 < RADIX ? tot[d1] : 0u;
    uint32_t v = t0 + t1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) s_wsum[warp] = v;
    __syncthreads();
    uint32_t off = 0;
    for (int w = 0; w < warp; ++w) off += s_wsum[w];
    const uint32_t excl = off + v - (t0 + t1);
    if (d0 < RADIX) s_base[d0] = excl + hist[(int64_t)blockIdx.x * RADIX + d0];
    if (d1 < RADIX) s_base[d1] = excl + t0 + hist[(int64_t)blockIdx.x * RADIX + d1];
  }

  const int64_t begin = (int64_t)blockIdx.x * tiles_per_block * RS_TILE;
  for (int t = 0; t < tiles_per_block; ++t) {
    const int64_t tile0 = begin + (int64_t)t * RS_TILE;
    if (tile0 >= n) break;
    for (int d = tid; d < RS_WARPS * RADIX; d += RS_THREADS) s_cnt[d >> rbits][d & (RADIX - 1)] = 0;
    __syncthreads();

    U key[RS_ITEMS];
    int32_t val[RS_ITEMS];
    uint32_t dig[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
      const int64_t idx = tile0 + warp * (RS_ITEMS * 32) + i * 32 + lane;
      if (idx < n) {
        key[i] = load_key<KeyT, RAW>(keys_in, idx, bound);
        val[i] = RAW ? (int32_t)idx : vals_in[idx];
        dig[i] = (uint32_t)(key[i] >> shift) & (RADIX - 1);
      } else {
        key[i] = 0;
        val[i] = 0;
        dig[i] = RADIX;  // invalid marker
      }
    }
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
      const uint32_t peers = __match_any_sync(0xffffffffu, dig[i]);
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (lane == leader && dig[i] < RADIX) {
        old = s_cnt[warp][dig[i]];
        s_cnt[warp][dig[i]] = old + __popc(peers);
      }
      old = __shfl_sync(0xffffffffu, old, leader);
      rank[i] = old + __popc(peers & lt_mask);
      __syncwarp();
    }
    __syncthreads();
    // per-digit exclusive prefix over warps, plus the running block base
    for (int d = tid; d < RADIX; d += RS_THREADS) {
      uint32_t run = s_base[d];
#pragma unroll
      for (int w = 0; w < RS_WARPS; ++w) {
        const uint32_t c = s_cnt[w][d];
        s_cnt[w][d] = run;
        run += c;
      }
      s_base[d] = run;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
      if (dig[i] < RADIX) {
        const uint32_t pos = s_cnt[warp][dig[i]] + rank[i];
        keys_out[pos] = key[i];
        vals_out[pos] = val[i];
        // histogram of the NEXT pass, binned by the block that will own this element's new position:
        // saves that pass's histogram kernel (hist_next was zeroed by a memset node)
        if (hist_next)
          atomicAdd(&hist_next[(int64_t)(pos / (uint32_t)(tiles_per_block * RS_TILE)) * RADIX +
                               ((uint32_t)(key[i] >> next_shift) & (RADIX - 1))], 1u);
      }
    }
    __syncthreads();
  }
}

// ---- unique from sorted keys ----
template <typename UKey>
__global__ void __launch_bounds__(RS_THREADS)
seg_count_kernel(const UKey* __restrict__ sorted, int64_t n, uint32_t* __restrict__ tile_heads,
                 const int32_t* __restrict__ n_valid) {
  __shared__ uint32_t s_warp[RS_WARPS];
  n = eff_n(n, n_valid);
  if ((int64_t)blockIdx.x * RS_TILE >= n) {
    if (threadIdx.x == 0) tile_heads[blockIdx.x] = 0;
    return;
  }
  const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)threadIdx.x * RS_ITEMS;
  uint32_t c = 0;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const int64_t i = base + k;
    if (i < n) c += (i == 0 || sorted[i] != sorted[i - 1]) ? 1u : 0u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < RS_WARPS; ++w) t += s_warp[w];
    tile_heads[blockIdx.x] = t;
  }
}

// single block: exclusive scan of tile_heads (in place), total -> count[0] and seg_start[total] = n
__global__ void __launch_bounds__(1024)
seg_scan_kernel(uint32_t* __restrict__ tile_heads, int n_tiles, int32_t* __restrict__ count,
                int32_t* __restrict__ seg_start, int64_t n, const int32_t* __restrict__ n_valid) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  n = eff_n(n, n_valid);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n_tiles; base += 1024) {
    const int i = base + tid;
    const uint32_t x = (i < n_tiles) ? tile_heads[i] : 0u;
    uint32_t v = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t warp_off = warp ? s_warp[warp - 1] : 0u;
    const uint32_t carry = s_carry;
    if (i < n_tiles) tile_heads[i] = carry + warp_off + v - x;
    __syncthreads();
    if (tid == 1023) s_carry = carry + warp_off + v;
    __syncthreads();
  }
  if (tid == 0) {
    count[0] = (int32_t)s_carry;
    seg_start[s_carry] = (int32_t)n;
  }
}

template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS)
seg_emit_kernel(const typename UKeyOf<KeyT>::type* __restrict__ sorted,
                const int32_t* __restrict__ perm, int64_t n, uint64_t bound,
                const uint32_t* __restrict__ tile_offset, KeyT* __restrict__ uniq,
                int32_t* __restrict__ inverse, int32_t* __restrict__ seg_start,
                int32_t* __restrict__ seg_of, const int32_t* __restrict__ n_valid) {
  __shared__ uint32_t s_warp[RS_WARPS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  n = eff_n(n, n_valid);
  if ((int64_t)blockIdx.x * RS_TILE >= n) return;
  const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)tid * RS_ITEMS;
  uint32_t head[RS_ITEMS];
  uint32_t c = 0;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const int64_t i = base + k;
    head[k] = (i < n && (i == 0 || sorted[i] != sorted[i - 1])) ? 1u : 0u;
    c += head[k];
  }
  uint32_t v = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) s_warp[warp] = v;
  __syncthreads();
  uint32_t off = tile_offset[blockIdx.x];
  for (int w = 0; w < warp; ++w) off += s_warp[w];
  uint32_t seg = off + v - c;  // number of heads strictly before this thread's first item
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const int64_t i = base + k;
    if (i < n) {
      seg += head[k];
      const int32_t s = (int32_t)seg - 1;
      seg_of[i] = s;
      inverse[perm[i]] = s;
      if (head[k]) {
        uniq[s] = decode_key<KeyT>(sorted[i], bound);
        seg_start[s] = (int32_t)i;
      }
    }
  }
}

// =============================================================================================
// One-sweep form (round 2): 1 memset + (1 + passes + 1) launches instead of 3 * passes + 3.
//
//   onesweep_hist_kernel   digit totals of EVERY pass in one read of the keys (they do not depend on the order)
//   onesweep_pass_kernel   one launch per digit: a CTA takes the next tile (atomic ticket = tile index, so tile t only
//                          ever waits for tiles that already run), ranks its keys stably (warp match-any + per-warp
//                          counters, as the LSD kernels above), publishes its per-digit counts, obtains the counts
//                          of all earlier tiles by decoupled look-back (per digit: walk back over the published
//                          aggregates until a tile with an inclusive prefix is met) and scatters
//   onesweep_seg_kernel    head flags -> segment ids -> uniq / inverse / seg_start / seg_of / count, tile offsets by
//                          decoupled look-back too (one warp looks at 32 predecessors at a time)
// The problem is latency bound (N = 624 000 keys = 2.5 MB, everything lives in the L2): tiles are LARGE (16 384 int32
// keys, 8 192 int64 keys: 39 tiles at N = 624 000) so that the look-back chain is a few batched reads, not hundreds.
// A look-back word is (flag << 30) | count: flag and payload travel in one 32-bit store, no fence is needed.
constexpr int OS_THREADS = 512;
constexpr int OS_WARPS = OS_THREADS / 32;
constexpr uint32_t LB_AGG = 1u << 30, LB_INC = 2u << 30, LB_MASK = (1u << 30) - 1u;
constexpr int OS_MAX_PASSES = 8;
constexpr int kSpinLimit = 1 << 22;   // ~3 s of L2 round trips
// Keys per thread of a pass tile: 8, 12 or 16 (4096 / 6144 / 8192 keys per tile), the smallest that fits the sort into
// ONE wave of CTAs (one CTA per SM: 148 tiles) — a 149th tile would wait for a free SM and double the pass.
constexpr int OS_MAX_ITEMS = 16;
static int os_items_for(int64_t n) {
  for (int it = 8; it < OS_MAX_ITEMS; it += 4)
    if (cdiv(n > 0 ? n : 1, (int64_t)OS_THREADS * it) <= kNumSMs) return it;
  return OS_MAX_ITEMS;
}

__device__ __forceinline__ uint32_t ld_vol_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_vol_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ghist[pass][digit] += occurrences, through a shared-memory histogram per CTA.  (A warp-aggregated form built on
// MATCH.ANY was measured 10x slower here, r2d: 45 us — the instruction's cost grows with the number of DISTINCT values in
// the warp, and random digits are all distinct.)
constexpr int OS_HIST_THREADS = 1024;
template <typename KeyT>
__global__ void __launch_bounds__(OS_HIST_THREADS)
onesweep_hist_kernel(const KeyT* __restrict__ keys, int64_t n, uint64_t bound, int passes, int digit_bits,
                     uint32_t* __restrict__ ghist, const int32_t* __restrict__ n_valid) {
  __shared__ uint32_t s_hist[OS_MAX_PASSES * MAX_RADIX];
  const int radix = 1 << digit_bits;
  n = eff_n(n, n_valid);
  for (int d = threadIdx.x; d < passes * radix; d += OS_HIST_THREADS) s_hist[d] = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * OS_HIST_THREADS;
  for (int64_t i = (int64_t)blockIdx.x * OS_HIST_THREADS + threadIdx.x; i < n; i += stride) {
    const auto k = encode_key<KeyT>(keys[i], bound);
    for (int p = 0; p < passes; ++p) atomicAdd(&s_hist[p * radix + ((uint32_t)(k >> (p * digit_bits)) & (radix - 1))], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < passes * radix; d += OS_HIST_THREADS) {
    const uint32_t c = s_hist[d];
    if (c) atomicAdd(&ghist[(d / radix) * MAX_RADIX + (d & (radix - 1))], c);
  }
}

// Lanes of the warp that hold the same digit, from one ballot per digit bit: constant cost, whereas MATCH.ANY iterates
// over the distinct values of the warp (32 for random digits; measured r2d: 18-24 us per pass with it).
__device__ __forceinline__ uint32_t same_digit_lanes(uint32_t dig, int bits) {
  uint32_t peers = 0xffffffffu;
  for (int b = 0; b < bits; ++b) {
    const bool one = (dig >> b) & 1u;
    const uint32_t vote = __ballot_sync(0xffffffffu, one);
    peers &= one ? vote : ~vote;
  }
  return peers;
}

template <typename KeyT, bool RAW, int ITEMS>
__global__ void __launch_bounds__(OS_THREADS, 1)
onesweep_pass_kernel(const void* __restrict__ keys_in, const int32_t* __restrict__ vals_in,
                     typename UKeyOf<KeyT>::type* __restrict__ keys_out, int32_t* __restrict__ vals_out, int64_t n,
                     int shift, int radix, uint64_t bound, const uint32_t* __restrict__ ghist /* this pass */,
                     uint32_t* __restrict__ lookback /* [tiles][radix], zeroed */, uint32_t* __restrict__ ticket,
                     const int32_t* __restrict__ n_valid) {
  using U = typename UKeyOf<KeyT>::type;
  constexpr int TILE = OS_THREADS * ITEMS;
  // dynamic shared memory: first the per-warp look-back partials [OS_WARPS][radix] u32, later (aliased) the tile's
  // keys and values in their sorted order
  extern __shared__ uint4 os_dyn[];
  uint32_t* s_pref = reinterpret_cast<uint32_t*>(os_dyn);
  U* s_key = reinterpret_cast<U*>(os_dyn);
  int32_t* s_val = reinterpret_cast<int32_t*>(s_key + TILE);
  __shared__ uint16_t s_cnt[OS_WARPS][MAX_RADIX];   // per-warp digit counts (<= 32 * ITEMS), then exclusive over warps
  __shared__ uint32_t s_goff[MAX_RADIX];            // global position of sorted-tile index 0 of each digit (mod 2^32)
  __shared__ uint32_t s_dstart[MAX_RADIX];          // first sorted-tile index of each digit
  __shared__ uint32_t s_wsum[OS_WARPS];
  __shared__ int s_tile;
  const int RADIX = radix;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  n = eff_n(n, n_valid);
  if (tid == 0) s_tile = (int)atomicAdd(ticket, 1u);
  for (int d = tid; d < OS_WARPS * MAX_RADIX / 2; d += OS_THREADS) reinterpret_cast<uint32_t*>(&s_cnt[0][0])[d] = 0;
  __syncthreads();
  const int tile = s_tile;
  const int64_t tile0 = (int64_t)tile * TILE;
  if (tile0 >= n) return;
  const int tile_n = (int)min((int64_t)TILE, n - tile0);
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int digit_bits1 = __ffs(RADIX);                 // log2(RADIX) + 1: the extra bit tells the invalid marker apart

  U key[ITEMS];
  int32_t val[ITEMS];
  uint32_t rk[ITEMS / 2];                           // rank inside (warp, digit), two 16-bit ranks per register
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int li = warp * (ITEMS * 32) + i * 32 + lane;
    if (li < tile_n) {
      key[i] = load_key<KeyT, RAW>(keys_in, tile0 + li, bound);
      val[i] = RAW ? (int32_t)(tile0 + li) : vals_in[tile0 + li];
    } else {
      key[i] = 0;
      val[i] = 0;
    }
  }
  // digit base = exclusive scan of this pass's digit totals (one digit per thread); overlaps the key loads
  uint32_t base_d = 0;
  {
    const uint32_t t0 = tid < RADIX ? ghist[tid] : 0u;
    uint32_t v = t0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) s_wsum[warp] = v;
    __syncthreads();
    uint32_t off = 0;
    for (int w = 0; w < warp; ++w) off += s_wsum[w];
    base_d = off + v - t0;
  }
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int li = warp * (ITEMS * 32) + i * 32 + lane;
    const uint32_t dig = (li < tile_n) ? ((uint32_t)(key[i] >> shift) & (RADIX - 1)) : (uint32_t)RADIX;   // RADIX = invalid
    const uint32_t peers = same_digit_lanes(dig, digit_bits1);
    const int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (lane == leader && dig < (uint32_t)RADIX) {
      old = s_cnt[warp][dig];
      s_cnt[warp][dig] = (uint16_t)(old + __popc(peers));
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    const uint32_t rank = old + __popc(peers & lt_mask);
    if (i & 1) rk[i >> 1] |= rank << 16;
    else rk[i >> 1] = rank;
    __syncwarp();
  }
  __syncthreads();
  // thread = digit: exclusive prefix over the warps, tile count -> published for the later tiles
  uint32_t run = 0;
  if (tid < RADIX) {
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w) {
      const uint32_t c = s_cnt[w][tid];
      s_cnt[w][tid] = (uint16_t)run;
      run += c;
    }
    st_vol_u32(lookback + (int64_t)tile * RADIX + tid, run | LB_AGG);
  }
  // look back: the counts of ALL earlier tiles, summed.  Warp w takes tiles w, w + 16, ...; a lane reads 4 digits
  // per 16-byte load, every load of the warp's rows is issued before the first is used.  Earlier tiles hold earlier
  // tickets, so they run and publish without waiting for anyone; the spin bound only turns a bug into wrong output.
  if ((RADIX & 127) == 0) {
    const int chunks = RADIX >> 7;                  // 16-byte loads per lane and row (<= 4)
    uint4 acc[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = make_uint4(0u, 0u, 0u, 0u);
    constexpr int RB = ITEMS >= 16 ? 2 : 4;         // rows per batch: RB * chunks 16-byte loads in flight per lane
    for (int t0 = warp; t0 < tile; t0 += RB * OS_WARPS) {
      uint4 v[RB][4];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int t = t0 + r * OS_WARPS;
        const uint4* row = reinterpret_cast<const uint4*>(lookback + (int64_t)min(t, tile - 1) * RADIX);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < chunks)
            asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v[r][c].x), "=r"(v[r][c].y), "=r"(v[r][c].z), "=r"(v[r][c].w) : "l"(row + c * 32 + lane) : "memory");
      }
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int t = t0 + r * OS_WARPS;
        if (t >= tile) break;
        const uint4* row = reinterpret_cast<const uint4*>(lookback + (int64_t)t * RADIX);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < chunks) {
            uint4 x = v[r][c];
            for (int spin = 0; ((x.x & x.y & x.z & x.w) >> 30) == 0u && spin < kSpinLimit; ++spin)
              asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                           : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(row + c * 32 + lane) : "memory");
            acc[c].x += x.x & LB_MASK; acc[c].y += x.y & LB_MASK; acc[c].z += x.z & LB_MASK; acc[c].w += x.w & LB_MASK;
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < chunks) reinterpret_cast<uint4*>(s_pref + warp * RADIX)[c * 32 + lane] = acc[c];
  } else {
    for (int d = lane; d < RADIX; d += 32) {
      uint32_t a = 0;
      for (int t = warp; t < tile; t += OS_WARPS) {
        uint32_t v = ld_vol_u32(lookback + (int64_t)t * RADIX + d);
        for (int spin = 0; (v >> 30) == 0u && spin < kSpinLimit; ++spin) v = ld_vol_u32(lookback + (int64_t)t * RADIX + d);
        a += v & LB_MASK;
      }
      s_pref[warp * RADIX + d] = a;
    }
  }
  // tile-local exclusive scan of the digit counts (first sorted-tile index of each digit)
  {
    uint32_t v = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) s_wsum[warp] = v;
    __syncthreads();                                // also: s_pref complete
    uint32_t off = 0;
    for (int w = 0; w < warp; ++w) off += s_wsum[w];
    if (tid < RADIX) {
      const uint32_t dstart = off + v - run;
      uint32_t excl = 0;
#pragma unroll
      for (int w = 0; w < OS_WARPS; ++w) excl += s_pref[w * RADIX + tid];
      s_dstart[tid] = dstart;
      s_goff[tid] = base_d + excl - dstart;
    }
  }
  __syncthreads();                                  // s_pref is dead: the same memory now takes the sorted tile
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int li = warp * (ITEMS * 32) + i * 32 + lane;
    if (li < tile_n) {
      const uint32_t dig = (uint32_t)(key[i] >> shift) & (RADIX - 1);
      const uint32_t si = s_dstart[dig] + s_cnt[warp][dig] + ((rk[i >> 1] >> (16 * (i & 1))) & 0xffffu);
      s_key[si] = key[i];
      s_val[si] = val[i];
    }
  }
  __syncthreads();
  // write-out: consecutive threads hold consecutive sorted-tile indices = consecutive addresses inside a digit run
  for (int j = tid; j < tile_n; j += OS_THREADS) {
    const U k = s_key[j];
    const uint32_t pos = s_goff[(uint32_t)(k >> shift) & (RADIX - 1)] + (uint32_t)j;
    keys_out[pos] = k;
    vals_out[pos] = s_val[j];
  }
}

// count + scan + emit in one launch.  Global accesses are striped (thread t touches tile0 + k * 512 + t: coalesced);
// the position-ordered scan works on a blocked view of the head flags in shared memory.
constexpr int SEG_ITEMS = 8;
constexpr int SEG_TILE = OS_THREADS * SEG_ITEMS;   // 4096
template <typename KeyT>
__global__ void __launch_bounds__(OS_THREADS)
onesweep_seg_kernel(const typename UKeyOf<KeyT>::type* __restrict__ sorted, const int32_t* __restrict__ perm, int64_t n,
                    uint64_t bound, uint32_t* __restrict__ lookback /* [tiles], zeroed */, uint32_t* __restrict__ ticket,
                    KeyT* __restrict__ uniq, int32_t* __restrict__ inverse, int32_t* __restrict__ count,
                    int32_t* __restrict__ seg_start, int32_t* __restrict__ seg_of, const int32_t* __restrict__ n_valid) {
  __shared__ __align__(8) uint8_t s_head[SEG_TILE];
  __shared__ uint32_t s_seg[SEG_TILE + SEG_TILE / 32];   // padded: the blocked writes of a warp spread over the banks
  __shared__ uint32_t s_warp[OS_WARPS];
  __shared__ uint32_t s_excl;
  __shared__ int s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  n = eff_n(n, n_valid);
  if (tid == 0) s_tile = (int)atomicAdd(ticket, 1u);
  __syncthreads();
  const int tile = s_tile;
  const int64_t tile0 = (int64_t)tile * SEG_TILE;
  if (n == 0) {
    if (tile == 0 && tid == 0) { count[0] = 0; seg_start[0] = 0; }
    return;
  }
  if (tile0 >= n) return;
#pragma unroll
  for (int k = 0; k < SEG_ITEMS; ++k) {
    const int li = k * OS_THREADS + tid;
    const int64_t i = tile0 + li;
    uint8_t h = 0;
    if (i < n) h = (i == 0 || sorted[i] != sorted[i - 1]) ? 1 : 0;
    s_head[li] = h;
  }
  __syncthreads();
  // blocked view: thread t owns positions [8t, 8t + 8)
  const uint2 hb = reinterpret_cast<const uint2*>(s_head)[tid];
  const uint32_t c = __popc(hb.x) + __popc(hb.y);
  uint32_t v = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) s_warp[warp] = v;
  __syncthreads();
  uint32_t before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < OS_WARPS; ++w) {
    const uint32_t x = s_warp[w];
    if (w < warp) before += x;
    total += x;
  }
  if (warp == 0) {
    if (lane == 0) st_vol_u32(lookback + tile, total | (tile == 0 ? LB_INC : LB_AGG));
    uint32_t excl = 0;
    for (int t = tile - 1; t >= 0; t -= 32) {
      const int tq = t - lane;
      uint32_t x = (tq >= 0) ? ld_vol_u32(lookback + tq) : LB_INC;
      for (int spin = 0; __any_sync(0xffffffffu, (x >> 30) == 0u) && spin < kSpinLimit; ++spin)
        if ((x >> 30) == 0u) x = ld_vol_u32(lookback + tq);
      const uint32_t inc_mask = __ballot_sync(0xffffffffu, (x & LB_INC) != 0u);
      const int first_inc = inc_mask ? (__ffs(inc_mask) - 1) : 32;
      uint32_t part = (lane <= first_inc) ? (x & LB_MASK) : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      excl += part;
      if (inc_mask) break;
    }
    if (lane == 0) {
      if (tile != 0) st_vol_u32(lookback + tile, (excl + total) | LB_INC);
      s_excl = excl;
      if (tile0 + SEG_TILE >= n) {                   // the tile that holds the last key closes the outputs
        count[0] = (int32_t)(excl + total);
        seg_start[excl + total] = (int32_t)n;
      }
    }
  }
  __syncthreads();
  {
    uint32_t seg = s_excl + before + v - c;          // heads strictly before this thread's first position
    const uint64_t bits = (uint64_t)hb.x | ((uint64_t)hb.y << 32);
#pragma unroll
    for (int j = 0; j < SEG_ITEMS; ++j) {
      seg += (uint32_t)((bits >> (8 * j)) & 1u);
      const int li = tid * SEG_ITEMS + j;
      s_seg[li + (li >> 5)] = seg - 1u;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < SEG_ITEMS; ++k) {
    const int li = k * OS_THREADS + tid;
    const int64_t i = tile0 + li;
    if (i < n) {
      const int32_t sgi = (int32_t)s_seg[li + (li >> 5)];
      seg_of[i] = sgi;
      inverse[perm[i]] = sgi;
      if (s_head[li]) {
        uniq[sgi] = decode_key<KeyT>(sorted[i], bound);
        seg_start[sgi] = (int32_t)i;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct SortPlan {
  int64_t n;
  int n_tiles, n_blocks, tiles_per_block, passes, digit_bits;
  size_t off_keys_a, off_keys_b, off_vals_tmp, off_hist, off_hist2, off_tiles, total;
  // one-sweep control region (zeroed by one memset per call): digit totals, tickets, look-back words
  int os_tiles, os_items, seg_tiles;
  size_t off_ctrl, ctrl_bytes, off_ghist, off_ticket, off_lb_seg, off_lb;
};

static SortPlan make_plan(int64_t n, int key_bytes, int key_bits) {
  SortPlan p;
  p.n = n;
  p.n_tiles = (int)cdiv(n > 0 ? n : 1, RS_TILE);
  const int max_blocks = kNumSMs * 8;
  p.tiles_per_block = (int)cdiv(p.n_tiles, max_blocks);
  p.n_blocks = (int)cdiv(p.n_tiles, p.tiles_per_block);
  p.passes = (int)cdiv(key_bits, MAX_RADIX_BITS);
  if (p.passes < 1) p.passes = 1;
  p.digit_bits = (int)cdiv(key_bits, p.passes);
  size_t o = 0;
  p.off_keys_a = o; o = align_up(o + (size_t)n * key_bytes, 256);
  p.off_keys_b = o; o = align_up(o + (size_t)n * key_bytes, 256);
  p.off_vals_tmp = o; o = align_up(o + (size_t)n * 4, 256);
  p.off_hist = o; o = align_up(o + (size_t)(p.n_blocks + 1) * MAX_RADIX * 4, 256);
  p.off_hist2 = o; o = align_up(o + (size_t)(p.n_blocks + 1) * MAX_RADIX * 4, 256);
  p.off_tiles = o; o = align_up(o + (size_t)(p.n_tiles + 1) * 4, 256);
  {
    p.os_items = os_items_for(n);
    p.os_tiles = (int)cdiv(n > 0 ? n : 1, (int64_t)OS_THREADS * p.os_items);
    p.seg_tiles = (int)cdiv(n > 0 ? n : 1, SEG_TILE);
    p.off_ctrl = o;
    p.off_ghist = o; o += (size_t)OS_MAX_PASSES * MAX_RADIX * 4;
    p.off_ticket = o; o += 64;
    p.off_lb_seg = o; o = align_up(o + (size_t)p.seg_tiles * 4, 256);
    // the look-back area is sized for the widest sort of this key type (the workspace is requested without knowing
    // the bound); one call zeroes only the passes * tiles * radix words it uses
    p.off_lb = o;
    p.ctrl_bytes = o + (size_t)p.passes * p.os_tiles * ((size_t)1 << p.digit_bits) * 4 - p.off_ctrl;
    o = align_up(o + (size_t)cdiv(8 * key_bytes, MAX_RADIX_BITS) * p.os_tiles * MAX_RADIX * 4, 256);
  }
  p.total = o;
  return p;
}

size_t unique_workspace_bytes(int64_t n, int key_bytes) {
  return make_plan(n, key_bytes, 8 * key_bytes).total;
}

int key_bits_for_bound(uint64_t bound) {
  // keys are clamped to [0, bound] -> need bits to represent `bound`
  int b = 1;
  while (b < 64 && (bound >> b) != 0) ++b;
  return b;
}

template <typename KeyT>
int unique_sorted(const KeyT* ids, int64_t n, uint64_t bound, KeyT* uniq, int32_t* inverse,
                  int32_t* count, int32_t* perm, int32_t* seg_start, int32_t* seg_of, void* ws,
                  size_t ws_bytes, cudaStream_t stream, const int32_t* n_valid) {
  using U = typename UKeyOf<KeyT>::type;
  const int key_bits = bound ? key_bits_for_bound(bound) : 8 * (int)sizeof(KeyT);
  const SortPlan p = make_plan(n, sizeof(KeyT), key_bits);
  if (ws_bytes < p.total)
    return fail(ERR_WORKSPACE, "mrec_unique: workspace %zu bytes < required %zu", ws_bytes, p.total);
  if (reinterpret_cast<uintptr_t>(ws) % 16 != 0)
    return fail(ERR_ALIGN, "mrec_unique: workspace must be 16-byte aligned");
  if (n >= (int64_t)1 << 31) return fail(ERR_SHAPE, "mrec_unique: N must be < 2^31");
  if (n == 0) {
    cudaMemsetAsync(count, 0, sizeof(int32_t), stream);
    cudaMemsetAsync(seg_start, 0, sizeof(int32_t), stream);
    return OK;
  }
  char* w = reinterpret_cast<char*>(ws);
  U* kbuf[2] = {reinterpret_cast<U*>(w + p.off_keys_a), reinterpret_cast<U*>(w + p.off_keys_b)};
  int32_t* vtmp = reinterpret_cast<int32_t*>(w + p.off_vals_tmp);
  uint32_t* hist = reinterpret_cast<uint32_t*>(w + p.off_hist);
  uint32_t* tiles = reinterpret_cast<uint32_t*>(w + p.off_tiles);

  const int radix = 1 << p.digit_bits;
  static const bool use_lsd = [] { const char* e = getenv("MREC_UNIQUE_LSD"); return e && atoi(e) != 0; }();
  if (!use_lsd && n < ((int64_t)1 << 30) && p.passes <= OS_MAX_PASSES) {
    uint32_t* ghist = reinterpret_cast<uint32_t*>(w + p.off_ghist);
    uint32_t* ticket = reinterpret_cast<uint32_t*>(w + p.off_ticket);
    uint32_t* lb_seg = reinterpret_cast<uint32_t*>(w + p.off_lb_seg);
    uint32_t* lb = reinterpret_cast<uint32_t*>(w + p.off_lb);
    cudaMemsetAsync(w + p.off_ctrl, 0, p.ctrl_bytes, stream);
    // dynamic shared memory: the sorted tile (keys + values) or the look-back partials, whichever is larger (opt-in
    // above 48 KB, set once per instantiation)
    constexpr size_t os_smem_max = (size_t)OS_THREADS * OS_MAX_ITEMS * (sizeof(U) + 4);
    static const bool attr_ok = [] {
      bool ok = true;
#define MREC_OS_ATTR(IT)                                                                                                       \
  ok = ok && cudaFuncSetAttribute(onesweep_pass_kernel<KeyT, true, IT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)os_smem_max) == cudaSuccess && \
       cudaFuncSetAttribute(onesweep_pass_kernel<KeyT, false, IT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)os_smem_max) == cudaSuccess;
      MREC_OS_ATTR(8) MREC_OS_ATTR(12) MREC_OS_ATTR(16)
#undef MREC_OS_ATTR
      return ok;
    }();
    if (!attr_ok) return fail(ERR_CUDA, "mrec_unique: cannot reserve %zu bytes of dynamic shared memory", os_smem_max);
    const size_t os_smem = std::max((size_t)OS_THREADS * p.os_items * (sizeof(U) + 4), (size_t)OS_WARPS * MAX_RADIX * 4);
    const int hist_grid = (int)std::min<int64_t>(cdiv(n, 4 * OS_HIST_THREADS), kNumSMs);
    MREC_LAUNCH(onesweep_hist_kernel<KeyT>, hist_grid, OS_HIST_THREADS, 0, stream, ids, n, bound, p.passes, p.digit_bits,
                ghist, n_valid);
    const void* kin = ids;
    const int32_t* vin = nullptr;
    for (int pass = 0; pass < p.passes; ++pass) {
      U* kout = kbuf[pass & 1];
      int32_t* vout = (((p.passes - 1 - pass) & 1) == 0) ? perm : vtmp;    // the last pass lands in `perm`
      uint32_t* lbp = lb + (size_t)pass * p.os_tiles * radix;
#define MREC_OS_PASS(RAW_, IT)                                                                                          \
  MREC_LAUNCH((onesweep_pass_kernel<KeyT, RAW_, IT>), p.os_tiles, OS_THREADS, os_smem, stream, kin, vin, kout, vout, n, \
              pass * p.digit_bits, radix, bound, ghist + pass * MAX_RADIX, lbp, ticket + pass, n_valid)
      if (pass == 0) {
        if (p.os_items == 8) MREC_OS_PASS(true, 8); else if (p.os_items == 12) MREC_OS_PASS(true, 12); else MREC_OS_PASS(true, 16);
      } else {
        if (p.os_items == 8) MREC_OS_PASS(false, 8); else if (p.os_items == 12) MREC_OS_PASS(false, 12); else MREC_OS_PASS(false, 16);
      }
#undef MREC_OS_PASS
      kin = kout;
      vin = vout;
    }
    MREC_LAUNCH(onesweep_seg_kernel<KeyT>, p.seg_tiles, OS_THREADS, 0, stream, reinterpret_cast<const U*>(kin), perm, n,
                bound, lb_seg, ticket + OS_MAX_PASSES, uniq, inverse, count, seg_start, seg_of, n_valid);
    return check_launch("unique_sorted");
  }
  const void* kin = ids;
  const int32_t* vin = nullptr;
  const int scan_grid = (int)cdiv(radix, 32);
  uint32_t* hbuf[2] = {hist, reinterpret_cast<uint32_t*>(w + p.off_hist2)};
  const size_t hist_bytes = (size_t)(p.n_blocks + 1) * radix * sizeof(uint32_t);
  for (int pass = 0; pass < p.passes; ++pass) {
    const int shift = pass * p.digit_bits;
    U* kout = kbuf[pass & 1];
    // the last pass must land in `perm`
    int32_t* vout = (((p.passes - 1 - pass) & 1) == 0) ? perm : vtmp;
    uint32_t* h = hbuf[pass & 1];
    const bool more = kFuseNextHist && (pass + 1 < p.passes);
    uint32_t* hn = more ? hbuf[(pass + 1) & 1] : nullptr;   // filled by this pass's scatter
    if (more) cudaMemsetAsync(hn, 0, hist_bytes, stream);
    if (pass == 0) {
      MREC_LAUNCH((radix_hist_kernel<KeyT, true>), p.n_blocks, RS_THREADS, 0, stream, kin, n, shift, radix,
                  bound, p.tiles_per_block, h, n_valid);
      MREC_LAUNCH(radix_scan_kernel, scan_grid, 1024, 0, stream, h, p.n_blocks, radix);
      MREC_LAUNCH((radix_scatter_kernel<KeyT, true>), p.n_blocks, RS_THREADS, 0, stream, kin, vin, kout,
                  vout, n, shift, radix, bound, p.tiles_per_block, h, p.n_blocks, hn, shift + p.digit_bits, n_valid);
    } else {
      if (!kFuseNextHist)
        MREC_LAUNCH((radix_hist_kernel<KeyT, false>), p.n_blocks, RS_THREADS, 0, stream, kin, n, shift, radix,
                    bound, p.tiles_per_block, h, n_valid);
      MREC_LAUNCH(radix_scan_kernel, scan_grid, 1024, 0, stream, h, p.n_blocks, radix);
      MREC_LAUNCH((radix_scatter_kernel<KeyT, false>), p.n_blocks, RS_THREADS, 0, stream, kin, vin, kout,
                  vout, n, shift, radix, bound, p.tiles_per_block, h, p.n_blocks, hn, shift + p.digit_bits, n_valid);
    }
    kin = kout;
    vin = vout;
  }
  const U* sorted = reinterpret_cast<const U*>(kin);
  MREC_LAUNCH(seg_count_kernel<U>, p.n_tiles, RS_THREADS, 0, stream, sorted, n, tiles, n_valid);
  MREC_LAUNCH(seg_scan_kernel, 1, 1024, 0, stream, tiles, p.n_tiles, count, seg_start, n, n_valid);
  MREC_LAUNCH(seg_emit_kernel<KeyT>, p.n_tiles, RS_THREADS, 0, stream, sorted, perm, n, bound, tiles,
              uniq, inverse, seg_start, seg_of, n_valid);
  return check_launch("unique_sorted");
}

template int unique_sorted<int32_t>(const int32_t*, int64_t, uint64_t, int32_t*, int32_t*, int32_t*,
                                    int32_t*, int32_t*, int32_t*, void*, size_t, cudaStream_t, const int32_t*);
template int unique_sorted<int64_t>(const int64_t*, int64_t, uint64_t, int64_t*, int32_t*, int32_t*,
                                    int32_t*, int32_t*, int32_t*, void*, size_t, cudaStream_t, const int32_t*);

// ---- first-occurrence order (upstream CPU Unique) ----
// Segments are re-ranked by their first position perm[seg_start[u]] (stable sort => minimum).
__global__ void first_pos_kernel(const int32_t* __restrict__ perm, const int32_t* __restrict__ seg_start,
                                 const int32_t* __restrict__ count, int32_t* __restrict__ first_pos,
                                 int64_t n) {
  const int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (u >= n) return;
  first_pos[u] = (u < count[0]) ? perm[seg_start[u]] : (int32_t)n;  // padding sorts last
}

template <typename KeyT>
__global__ void reorder_first_kernel(const int32_t* __restrict__ order /* rank -> ascending seg */,
                                     const int32_t* __restrict__ count, const KeyT* __restrict__ uniq_asc,
                                     KeyT* __restrict__ uniq_first, int32_t* __restrict__ remap, int64_t n) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n || r >= count[0]) return;
  const int32_t u = order[r];
  uniq_first[r] = uniq_asc[u];
  remap[u] = (int32_t)r;
}

__global__ void remap_inverse_kernel(int32_t* __restrict__ inverse, const int32_t* __restrict__ remap,
                                     int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) inverse[i] = remap[inverse[i]];
}

size_t unique_first_workspace_bytes(int64_t n, int key_bytes) {
  // ascending pass workspace + second (int32) sort workspace + 6 int32 arrays + uniq_asc
  return unique_workspace_bytes(n, key_bytes) + unique_workspace_bytes(n, 4) +
         align_up((size_t)(n + 1) * 4, 256) * 7 + align_up((size_t)n * key_bytes, 256);
}

template <typename KeyT>
int unique_first(const KeyT* ids, int64_t n, uint64_t bound, KeyT* uniq, int32_t* inverse,
                 int32_t* count, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (ws_bytes < unique_first_workspace_bytes(n, sizeof(KeyT)))
    return fail(ERR_WORKSPACE, "mrec_unique_first: workspace %zu bytes < required %zu", ws_bytes,
                unique_first_workspace_bytes(n, sizeof(KeyT)));
  char* w = reinterpret_cast<char*>(ws);
  size_t o = 0;
  void* ws1 = w + o; const size_t ws1_b = unique_workspace_bytes(n, sizeof(KeyT)); o += ws1_b;
  void* ws2 = w + o; const size_t ws2_b = unique_workspace_bytes(n, 4); o += ws2_b;
  const size_t arr = align_up((size_t)(n + 1) * 4, 256);
  int32_t* perm = reinterpret_cast<int32_t*>(w + o); o += arr;
  int32_t* seg_start = reinterpret_cast<int32_t*>(w + o); o += arr;
  int32_t* seg_of = reinterpret_cast<int32_t*>(w + o); o += arr;
  int32_t* first_pos = reinterpret_cast<int32_t*>(w + o); o += arr;
  int32_t* order = reinterpret_cast<int32_t*>(w + o); o += arr;
  int32_t* remap = reinterpret_cast<int32_t*>(w + o); o += arr;
  int32_t* scratch = reinterpret_cast<int32_t*>(w + o); o += arr;
  KeyT* uniq_asc = reinterpret_cast<KeyT*>(w + o);
  int rc = unique_sorted<KeyT>(ids, n, bound, uniq_asc, inverse, count, perm, seg_start, seg_of, ws1,
                               ws1_b, stream, nullptr);
  if (rc != OK || n == 0) return rc;
  const int grid = (int)cdiv(n, 256);
  MREC_LAUNCH(first_pos_kernel, grid, 256, 0, stream, perm, seg_start, count, first_pos, n);
  // sort (first_pos, segment) by first_pos; reuse the pair sort through unique_sorted's machinery:
  // keys are distinct for real segments, so `perm` of this second sort is the rank -> segment map.
  // Outputs of the second unique that we do not need land in scratch buffers.
  // After first_pos is built the first sort's perm/seg_start/seg_of are dead and are reused as the
  // second sort's throw-away outputs.
  rc = unique_sorted<int32_t>(first_pos, n, (uint64_t)n, /*uniq*/ seg_of, /*inverse*/ remap,
                              /*count*/ scratch, /*perm*/ order, /*seg_start*/ seg_start,
                              /*seg_of*/ perm, ws2, ws2_b, stream, nullptr);
  if (rc != OK) return rc;
  MREC_LAUNCH(reorder_first_kernel<KeyT>, grid, 256, 0, stream, order, count, uniq_asc, uniq, remap, n);
  MREC_LAUNCH(remap_inverse_kernel, grid, 256, 0, stream, inverse, remap, n);
  return check_launch("unique_first");
}

template int unique_first<int32_t>(const int32_t*, int64_t, uint64_t, int32_t*, int32_t*, int32_t*,
                                   void*, size_t, cudaStream_t);
template int unique_first<int64_t>(const int64_t*, int64_t, uint64_t, int64_t*, int32_t*, int32_t*,
                                   void*, size_t, cudaStream_t);

}  // namespace mrec

using namespace mrec;

// Workspace size helpers (plain C, host only).
MREC_API size_t mrec_unique_workspace_bytes(int64_t n, int key_bytes) {
  return unique_workspace_bytes(n, key_bytes);
}
MREC_API size_t mrec_unique_first_workspace_bytes(int64_t n, int key_bytes) {
  return unique_first_workspace_bytes(n, key_bytes);
}

// inputs : ids[N] i32|i64, (bounded variant: table_like[V, ...] — only its dim 0 is read)
// outputs: uniq[N] (ids dtype), inverse[N] i32, count[1] i32, perm[N] i32, seg_start[N+1] i32,
//          seg_of[N] i32, workspace[bytes] uint8|int8
static int unique_entry(const Aot& a, bool bounded) {
  // bounded: optional third input n_valid[1] i32 (entries at i >= n_valid are padding of a static buffer)
  const bool has_nv = bounded && a.nparam == 10;
  const int n_in = bounded ? (has_nv ? 3 : 2) : 1;
  if (a.nparam != n_in + 7)
    return fail(ERR_NPARAM, "mrec_unique%s: expected %d params, got %d", bounded ? "_bounded" : "",
                n_in + 7, a.nparam);
  const int o = n_in;
  const int64_t n = a.numel(0);
  for (int i = 0; i < a.nparam; ++i)
    if (!a.params[i] && a.numel(i) > 0 && !(bounded && i == 1))
      return fail(ERR_NULL, "mrec_unique: param %d is null", i);
  MREC_REQUIRE(a.is_i32(0) || a.is_i64(0), ERR_DTYPE, "mrec_unique: ids must be int32|int64");
  MREC_REQUIRE(strcmp(a.dtypes[0], a.dtypes[o]) == 0, ERR_DTYPE, "mrec_unique: uniq dtype must match ids");
  for (int i = 1; i <= 5; ++i)
    MREC_REQUIRE(a.is_i32(o + i), ERR_DTYPE, "mrec_unique: output %d must be int32", i);
  MREC_REQUIRE(a.numel(o) >= n && a.numel(o + 1) >= n && a.numel(o + 2) >= 1 && a.numel(o + 3) >= n &&
                   a.numel(o + 4) >= n + 1 && a.numel(o + 5) >= n,
               ERR_SHAPE, "mrec_unique: outputs must be padded to N (seg_start to N+1)");
  uint64_t bound = 0;
  const int32_t* n_valid = nullptr;
  if (bounded) {
    MREC_REQUIRE(a.ndims[1] >= 1 && a.dim(1, 0) > 0, ERR_SHAPE, "mrec_unique_bounded: bad table shape");
    bound = (uint64_t)a.dim(1, 0);
    if (a.is_i32(0)) MREC_REQUIRE(bound < 0x7fffffffull, ERR_SHAPE, "mrec_unique_bounded: V too large for int32 ids");
    if (has_nv) {
      MREC_REQUIRE(a.is_i32(2) && a.numel(2) >= 1, ERR_DTYPE, "mrec_unique_bounded: n_valid must be int32[1]");
      n_valid = a.ptr<int32_t>(2);
    }
  }
  const size_t ws_bytes = (size_t)a.numel(o + 6);
  if (a.is_i32(0))
    return unique_sorted<int32_t>(a.ptr<int32_t>(0), n, bound, a.ptr<int32_t>(o), a.ptr<int32_t>(o + 1),
                                  a.ptr<int32_t>(o + 2), a.ptr<int32_t>(o + 3), a.ptr<int32_t>(o + 4),
                                  a.ptr<int32_t>(o + 5), a.params[o + 6], ws_bytes, a.stream, n_valid);
  return unique_sorted<int64_t>(a.ptr<int64_t>(0), n, bound, a.ptr<int64_t>(o), a.ptr<int32_t>(o + 1),
                                a.ptr<int32_t>(o + 2), a.ptr<int32_t>(o + 3), a.ptr<int32_t>(o + 4),
                                a.ptr<int32_t>(o + 5), a.params[o + 6], ws_bytes, a.stream, n_valid);
}

MREC_API int mrec_unique(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                         void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  return unique_entry(a, false);
}

MREC_API int mrec_unique_bounded(int nparam, void** params, int* ndims, int64_t** shapes,
                                 const char** dtypes, void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  return unique_entry(a, true);
}

// First-occurrence order (the order of upstream's CPU Unique kernel).
// inputs : ids[N] ; outputs: uniq[N], inverse[N] i32, count[1] i32, workspace[bytes]
MREC_API int mrec_unique_first(int nparam, void** params, int* ndims, int64_t** shapes,
                               const char** dtypes, void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  MREC_CHECK_NPARAM(a, 5);
  const int64_t n = a.numel(0);
  MREC_REQUIRE(a.is_i32(0) || a.is_i64(0), ERR_DTYPE, "mrec_unique_first: ids must be int32|int64");
  MREC_REQUIRE(strcmp(a.dtypes[0], a.dtypes[1]) == 0, ERR_DTYPE, "mrec_unique_first: uniq dtype must match ids");
  MREC_REQUIRE(a.is_i32(2) && a.is_i32(3), ERR_DTYPE, "mrec_unique_first: inverse/count must be int32");
  MREC_REQUIRE(a.numel(1) >= n && a.numel(2) >= n && a.numel(3) >= 1, ERR_SHAPE,
               "mrec_unique_first: outputs must be padded to N");
  const size_t ws_bytes = (size_t)a.numel(4);
  if (a.is_i32(0))
    return unique_first<int32_t>(a.ptr<int32_t>(0), n, 0, a.ptr<int32_t>(1), a.ptr<int32_t>(2),
                                 a.ptr<int32_t>(3), a.params[4], ws_bytes, a.stream);
  return unique_first<int64_t>(a.ptr<int64_t>(0), n, 0, a.ptr<int64_t>(1), a.ptr<int32_t>(2),
                               a.ptr<int32_t>(3), a.params[4], ws_bytes, a.stream);
}

// ---- shard bucketing helper -----------------------------------------------------------------------
// bounds[r] = number of entries of the ascending array uniq[0 : count) that are < edges[r].
// With keys remapped owner-major (key' = owner * rows_per_rank + local_row) the unique keys of one rank form
// G contiguous runs; their boundaries are the all-to-all split sizes (mindrec_b200/sharded.py).
template <typename KeyT>
__global__ void shard_bounds_kernel(const KeyT* __restrict__ uniq, const int32_t* __restrict__ count,
                                    const KeyT* __restrict__ edges, int n_edges, int32_t* __restrict__ bounds) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_edges) return;
  const KeyT e = edges[r];
  int lo = 0, hi = count[0];
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (uniq[mid] < e) lo = mid + 1; else hi = mid;
  }
  bounds[r] = lo;
}

// in : uniq[N] i32|i64 (ascending, first count entries valid), count[1] i32, edges[E] (uniq dtype)
// out: bounds[E] i32
MREC_API int mrec_shard_bounds(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                               void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  MREC_CHECK_NPARAM(a, 4);
  MREC_REQUIRE((a.is_i32(0) || a.is_i64(0)) && strcmp(a.dtypes[0], a.dtypes[2]) == 0, ERR_DTYPE,
               "mrec_shard_bounds: uniq and edges must share an integer dtype");
  MREC_REQUIRE(a.is_i32(1) && a.is_i32(3), ERR_DTYPE, "mrec_shard_bounds: count/bounds must be int32");
  const int e = (int)a.numel(2);
  MREC_REQUIRE(a.numel(3) >= e, ERR_SHAPE, "mrec_shard_bounds: bounds must have one entry per edge");
  if (e == 0) return OK;
  if (a.is_i32(0)) {
    MREC_LAUNCH(shard_bounds_kernel<int32_t>, (int)cdiv(e, 128), 128, 0, a.stream, a.ptr<int32_t>(0),
                a.ptr<int32_t>(1), a.ptr<int32_t>(2), e, a.ptr<int32_t>(3));
  } else {
    MREC_LAUNCH(shard_bounds_kernel<int64_t>, (int)cdiv(e, 128), 128, 0, a.stream, a.ptr<int64_t>(0),
                a.ptr<int32_t>(1), a.ptr<int64_t>(2), e, a.ptr<int32_t>(3));
  }
  return check_launch("shard_bounds");
}

// key -> owner-major key' = (key mod G) * R + key div G ; keys outside [0, V) -> G * R (dropped by the bounded dedup)
template <typename KeyT>
__global__ void shard_remap_kernel(const KeyT* __restrict__ ids, KeyT* __restrict__ out, int64_t n,
                                   int64_t vocab, int world, int64_t rows_per_rank) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = (int64_t)ids[i];
    out[i] = ((uint64_t)k < (uint64_t)vocab) ? (KeyT)((k % world) * rows_per_rank + k / world)
                                             : (KeyT)((int64_t)world * rows_per_rank);
  }
}

// in : ids[...] i32|i64, table_like[V, ...] (dim 0 = global vocab), owners_like[G, R] (dims = world size, rows per rank)
// out: keys[...] (ids dtype)
MREC_API int mrec_shard_remap(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                              void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  if (a.nparam != 4) return fail(ERR_NPARAM, "mrec_shard_remap: expected 4 params, got %d", a.nparam);
  MREC_REQUIRE((a.is_i32(0) || a.is_i64(0)) && strcmp(a.dtypes[0], a.dtypes[3]) == 0, ERR_DTYPE,
               "mrec_shard_remap: ids/keys must share an integer dtype");
  MREC_REQUIRE(a.ndims[1] >= 1 && a.ndims[2] >= 2, ERR_SHAPE, "mrec_shard_remap: table_like[V,..], owners_like[G,R,..]");
  const int64_t n = a.numel(0), vocab = a.dim(1, 0), rows = a.dim(2, 1);
  const int world = (int)a.dim(2, 0);
  MREC_REQUIRE(a.numel(3) == n, ERR_SHAPE, "mrec_shard_remap: keys must match ids");
  MREC_REQUIRE(world >= 1 && rows * world >= vocab, ERR_SHAPE, "mrec_shard_remap: G * R must cover V");
  if (a.is_i32(0)) MREC_REQUIRE((int64_t)world * rows < 0x7fffffffll, ERR_SHAPE, "mrec_shard_remap: G*R exceeds int32");
  if (n == 0) return OK;
  if (!a.params[0] || !a.params[3]) return fail(ERR_NULL, "mrec_shard_remap: null ids/keys");
  if (a.is_i32(0)) {
    MREC_LAUNCH(shard_remap_kernel<int32_t>, grid_for(cdiv(n, 256), 8), 256, 0, a.stream, a.ptr<int32_t>(0),
                a.ptr<int32_t>(3), n, vocab, world, rows);
  } else {
    MREC_LAUNCH(shard_remap_kernel<int64_t>, grid_for(cdiv(n, 256), 8), 256, 0, a.stream, a.ptr<int64_t>(0),
                a.ptr<int64_t>(3), n, vocab, world, rows);
  }
  return check_launch("shard_remap");
}
