// Fused owner-side gather + NVLink peer store for row-sharded tables, and the CUDA-IPC plumbing it needs.
//
// Forward exchange of mindrec_b200/sharded.py: rank `me` owns the rows {k : k mod G = me}.  After the key
// all-to-all it holds rows[n_r] — local row ids requested by every rank, concatenated by source rank — and must
// return table[rows[i]] to the rank that asked.  Instead of gathering into a staging buffer and handing that to
// an NCCL all-to-all (one extra HBM write + read of every row, two collectives for the deep and wide tables),
// this kernel reads each row once and stores it directly into the requester's landing buffer through a
// peer-mapped pointer: the transfer rides the SM store path over NVLink 5 / NVSwitch and overlaps the gather
// tile by tile.  Row i of source s lands at  peer_ptr[s] + (peer_off[s] + i - src_off[s]) * D * 4  — exactly where
// the requester's dedup placed its s-th bucket, so the landing buffer is the same [U, D] array an all-to-all
// would have produced.  Completion is published by the stream-ordered barrier collective that follows.
#include "common.cuh"
#include <type_traits>

namespace mrec {

constexpr int kPeerMaxRanks = 16;

struct PeerTable {
  float* ptr[kPeerMaxRanks];
  int32_t dst_off[kPeerMaxRanks];
  int32_t src_off[kPeerMaxRanks + 1];
};

// `dirty` (optional): one bit per local table row, set for the rows the CURRENT step's update rewrites.  mode 1 serves
// only the rows whose bit is clear (early, underneath the current step's DenseLayers: their values cannot change any
// more before the next step reads them), mode 2 only the rows whose bit is set (after the update); mode 0 all rows.
template <typename Vec>
__global__ void __launch_bounds__(256)
gather_to_peers_kernel(const float* __restrict__ table, const int32_t* __restrict__ rows, int64_t n_rows, int cpr,
                       int64_t vocab, int world, const int64_t* __restrict__ peer_ptrs,
                       const int32_t* __restrict__ dst_off, const int32_t* __restrict__ src_off,
                       const uint32_t* __restrict__ dirty, int mode) {
  __shared__ PeerTable s_t;
  if (threadIdx.x < world) {
    s_t.ptr[threadIdx.x] = reinterpret_cast<float*>(peer_ptrs[threadIdx.x]);
    s_t.dst_off[threadIdx.x] = dst_off[threadIdx.x];
  }
  if (threadIdx.x <= world) s_t.src_off[threadIdx.x] = src_off[threadIdx.x];
  __syncthreads();
  // rows beyond src_off[world] are padding of a statically sized inbox (device-driven exchange)
  const int64_t total = min(n_rows, (int64_t)s_t.src_off[world]) * cpr;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e0 < total; e0 += 4 * stride) {
    Vec v[4];
    Vec* dst[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t e = e0 + k * stride;
      dst[k] = nullptr;
      v[k] = Vec();
      if (e < total) {
        const int64_t i = e / cpr;
        const int c = (int)(e - i * cpr);
        const int64_t row = rows[i];
        const bool in_range = (uint64_t)row < (uint64_t)vocab;
        bool take = true;
        if (mode != 0 && in_range) take = (((dirty[row >> 5] >> (row & 31)) & 1u) != 0u) == (mode == 2);
        if (take) {
          int s = 0;
          while (s + 1 < world && i >= s_t.src_off[s + 1]) ++s;   // G <= 16: a short scan of the bucket edges
          dst[k] = reinterpret_cast<Vec*>(s_t.ptr[s]) + ((int64_t)s_t.dst_off[s] + (i - s_t.src_off[s])) * cpr + c;
          if (in_range) v[k] = reinterpret_cast<const Vec*>(table)[row * cpr + c];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (dst[k]) *dst[k] = v[k];   // st.global on a peer-mapped address: goes out over NVLink
  }
}

// bitmap[r >> 5] |= 1 << (r & 31) for the first count[0] entries of rows
__global__ void bitmap_set_kernel(const int32_t* __restrict__ rows, const int32_t* __restrict__ count, int64_t n,
                                  uint32_t* __restrict__ bitmap, int64_t n_bits) {
  const int64_t m = min(n, (int64_t)count[0]);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = rows[i];
    if ((uint64_t)r < (uint64_t)n_bits) atomicOr(&bitmap[r >> 5], 1u << (r & 31));
  }
}

}  // namespace mrec

using namespace mrec;

// ---- CUDA IPC plumbing (host only, not aot; called once at set-up) -----------------------------------------
// Opens a cudaIpcMemHandle_t exported by another process of the same node and returns the mapped base pointer.
MREC_API void* mrec_ipc_open_handle(const char* handle64) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    fail(ERR_CUDA, "mrec_ipc_open_handle: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
MREC_API int mrec_ipc_close_handle(void* p) {
  return cudaIpcCloseMemHandle(p) == cudaSuccess ? OK : fail(ERR_CUDA, "mrec_ipc_close_handle failed");
}
MREC_API int mrec_ipc_get_handle(void* base_ptr, char* handle64) {
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, base_ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(ERR_CUDA, "mrec_ipc_get_handle: %s", cudaGetErrorString(e));
  }
  memcpy(handle64, &h, sizeof(h));
  return OK;
}
// Plain cudaMalloc / cudaFree for buffers that peers map (the caching allocator of the host framework may
// sub-allocate or use virtual-memory segments that cannot be exported with cudaIpcGetMemHandle).
MREC_API void* mrec_peer_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    fail(ERR_CUDA, "mrec_peer_alloc: cudaMalloc(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}
MREC_API int mrec_peer_free(void* p) { return cudaFree(p) == cudaSuccess ? OK : ERR_CUDA; }

// in : table[R,D] f32 (local shard), rows[n_r] i32 (local row ids, concatenated by source rank),
//      peer_ptrs[G] i64 (base address of every rank's landing buffer [cap, D], as mapped in THIS process),
//      dst_off[G] i32 (row offset inside rank s's landing buffer = s's bucket start for this owner),
//      src_off[G+1] i32 (rows of source s are rows[src_off[s] : src_off[s+1]])
//      optional: dirty[ceil(R/32)] i32 (one bit per local row), mode_like[mode, ..] (shape carrier: 1 = serve rows whose bit
//      is clear, 2 = rows whose bit is set, 0 = all)
// out: dummy[1] i32
MREC_API int mrec_gather_to_peers(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                                  void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  if (a.nparam != 6 && a.nparam != 8) return fail(ERR_NPARAM, "mrec_gather_to_peers: expected 6 or 8 params, got %d", a.nparam);
  MREC_REQUIRE(a.is_f32(0) && a.is_i32(1) && a.is_i64(2) && a.is_i32(3) && a.is_i32(4), ERR_DTYPE,
               "mrec_gather_to_peers: table f32, rows i32, peer_ptrs i64, dst_off/src_off i32");
  const int64_t vocab = a.dim(0, 0);
  const int dim = a.ndims[0] >= 2 ? (int)a.dim(0, 1) : 1;
  const int world = (int)a.numel(2);
  MREC_REQUIRE(world >= 1 && world <= kPeerMaxRanks, ERR_SHAPE, "mrec_gather_to_peers: 1 <= G <= %d", kPeerMaxRanks);
  MREC_REQUIRE(a.numel(3) >= world && a.numel(4) >= world + 1, ERR_SHAPE, "mrec_gather_to_peers: dst_off[G], src_off[G+1]");
  const uint32_t* dirty = nullptr;
  int mode = 0;
  if (a.nparam == 8) {                                // dirty[ceil(R / 32)] i32, mode_like[mode, ..] (1: clean rows, 2: dirty rows)
    MREC_REQUIRE(a.is_i32(5), ERR_DTYPE, "mrec_gather_to_peers: the dirty bitmap is int32");
    mode = (int)a.dim(6, 0);
    MREC_REQUIRE(mode >= 0 && mode <= 2 && a.numel(5) * 32 >= vocab, ERR_SHAPE,
                 "mrec_gather_to_peers: mode in {0,1,2}, one bitmap bit per table row");
    dirty = reinterpret_cast<const uint32_t*>(a.params[5]);
    if (mode != 0 && !dirty) return fail(ERR_NULL, "mrec_gather_to_peers: null bitmap");
  }
  const int64_t n = a.numel(1);
  if (n == 0) return OK;
  if (dim % 4 == 0) {
    MREC_REQUIRE(a.aligned(0, 16), ERR_ALIGN, "mrec_gather_to_peers: table must be 16-byte aligned");
    const int cpr = dim / 4;
    MREC_LAUNCH(gather_to_peers_kernel<float4>, grid_for(cdiv(n * cpr, 1024), 8), 256, 0, a.stream, a.ptr<float>(0),
                a.ptr<int32_t>(1), n, cpr, vocab, world, a.ptr<int64_t>(2), a.ptr<int32_t>(3), a.ptr<int32_t>(4), dirty, mode);
  } else {
    MREC_LAUNCH(gather_to_peers_kernel<float>, grid_for(cdiv(n * dim, 1024), 8), 256, 0, a.stream, a.ptr<float>(0),
                a.ptr<int32_t>(1), n, dim, vocab, world, a.ptr<int64_t>(2), a.ptr<int32_t>(3), a.ptr<int32_t>(4), dirty, mode);
  }
  return check_launch("gather_to_peers");
}

// in : rows[N] i32, count[1] i32 (only the first count rows are read)      out: bitmap[W] i32 — bit r is set for every row r
MREC_API int mrec_bitmap_set(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes, void* stream,
                             void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  if (a.nparam != 3) return fail(ERR_NPARAM, "mrec_bitmap_set: expected 3 params, got %d", a.nparam);
  MREC_REQUIRE(a.is_i32(0) && a.is_i32(1) && a.is_i32(2) && a.numel(1) >= 1, ERR_DTYPE, "mrec_bitmap_set: rows, count, bitmap int32");
  const int64_t n = a.numel(0);
  if (n == 0 || a.numel(2) == 0) return OK;
  MREC_LAUNCH(bitmap_set_kernel, grid_for(cdiv(n, 256), 4), 256, 0, a.stream, a.ptr<int32_t>(0), a.ptr<int32_t>(1), n,
              reinterpret_cast<uint32_t*>(a.params[2]), a.numel(2) * 32);
  return check_launch("bitmap_set");
}

// =================================================================================================
// Device-driven exchange protocol (no host-side sizes): every rank publishes its bucket bounds to all peers, and
// keys / rows / gradients are stored straight into peer inboxes at offsets each rank derives on the device from
// the bounds matrix  B[s][o] = first unique-key index of rank s's bucket for owner o  (B[s][G] = U_s):
//   cnt[s][o]   = B[s][o+1] - B[s][o]                     keys rank s asks of owner o
//   inbox_off   = sum_{s' < s} cnt[s'][o]                  where s's segment starts in o's inbox (keys and grads)
//   n_r[o]      = sum_s cnt[s][o]                          valid entries of o's inbox
// Phase completion is a pair of tiny kernels: `signal` stores an epoch into a flag slot on every peer (after a
// system-scope fence), `wait` spins (bounded) until all G local slots reached the epoch.  Kernel boundaries
// give the ordering of the data stores against the flags.  Everything is static-shaped, so the whole training
// step can be captured in one CUDA graph.
// =================================================================================================
namespace mrec {

// ctrl[0] = my rank, ctrl[1] = world size
__global__ void shard_offsets_kernel(const int32_t* __restrict__ ball, const int32_t* __restrict__ ctrl,
                                     int32_t* __restrict__ dst_off, int32_t* __restrict__ src_off,
                                     int32_t* __restrict__ inbox_off, int32_t* __restrict__ n_r) {
  const int me = ctrl[0], g = ctrl[1];
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int run = 0;
  for (int s = 0; s < g; ++s) {                      // what every rank asks of me
    src_off[s] = run;
    dst_off[s] = ball[s * (g + 1) + me];             // where rank s wants my rows in its landing buffer
    run += ball[s * (g + 1) + me + 1] - ball[s * (g + 1) + me];
  }
  src_off[g] = run;
  n_r[0] = run;
  for (int o = 0; o < g; ++o) {                      // where my segment starts in owner o's inbox
    int off = 0;
    for (int s = 0; s < me; ++s) off += ball[s * (g + 1) + o + 1] - ball[s * (g + 1) + o];
    inbox_off[o] = off;
  }
}

// rows[u, :] (`width` elements of 4 or 8 bytes per row) of my buckets -> owner o's inbox row
// inbox_off[o] + (u - B[me][o]).  transform_mod > 0: the element is an owner-major integer key and is sent as
// key % transform_mod (the owner's local row id, or the original hash key).
template <typename E>
__global__ void __launch_bounds__(256)
push_rows_to_peers_kernel(const E* __restrict__ rows, int width, const int32_t* __restrict__ my_bounds,
                          const int32_t* __restrict__ inbox_off, const int64_t* __restrict__ peer_ptrs, int world,
                          int64_t transform_mod, int64_t cap_rows, int32_t* __restrict__ err) {
  __shared__ int32_t s_b[kPeerMaxRanks + 1];
  __shared__ int32_t s_off[kPeerMaxRanks];
  __shared__ E* s_ptr[kPeerMaxRanks];
  if (threadIdx.x <= world) s_b[threadIdx.x] = my_bounds[threadIdx.x];
  if (threadIdx.x < world) {
    s_off[threadIdx.x] = inbox_off[threadIdx.x];
    s_ptr[threadIdx.x] = reinterpret_cast<E*>(peer_ptrs[threadIdx.x]);
  }
  __syncthreads();
  const int64_t total = (int64_t)s_b[world] * width;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = e / width;
    const int c = (int)(e - u * width);
    int o = 0;
    while (o + 1 < world && u >= s_b[o + 1]) ++o;
    const int64_t drow = (int64_t)s_off[o] + (u - s_b[o]);
    if (drow >= cap_rows) {                      // inbox capacity exceeded: flag it, drop the row
      if (err) atomicOr(err, 2);
      continue;
    }
    E v = rows[e];
    if constexpr (std::is_integral<E>::value) {
      if (transform_mod > 0) v = (E)((int64_t)v % transform_mod);
    }
    s_ptr[o][drow * width + c] = v;
  }
}

// owner-major key of a hash-table key: key' = owner << bits | key with owner = (mix64(key) >> 40) mod G (high hash
// bits: the table itself indexes its slots with the low ones).  Reserved (-1, -2), negative and >= 2^bits keys map
// to G << bits, which the bounded dedup drops.
__device__ __forceinline__ uint64_t peer_mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

template <typename KeyT>
__global__ void shard_remap_hash_kernel(const KeyT* __restrict__ keys, int64_t* __restrict__ out, int64_t n, int world,
                                        int bits) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = (int64_t)keys[i];
    const bool ok = k >= 0 && (k >> bits) == 0;
    const int64_t owner = (int64_t)((peer_mix64((uint64_t)k) >> 40) % (uint64_t)world);
    out[i] = ok ? ((owner << bits) | k) : ((int64_t)world << bits);
  }
}

// buf[i] = value for i >= *n_valid (the stale tail of a statically sized inbox)
template <typename E>
__global__ void fill_tail_kernel(E* __restrict__ buf, int64_t n, const int32_t* __restrict__ n_valid, const E* __restrict__ value) {
  const int64_t first = max(0, n_valid[0]);
  const E v = value[0];
  for (int64_t i = first + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    buf[i] = v;
}

// payload[K] -> every peer's payload slot, then (after a system fence) epoch -> every peer's flag slot.
__global__ void peer_signal_kernel(const int32_t* __restrict__ payload, int k, const int64_t* __restrict__ payload_ptrs,
                                   const int64_t* __restrict__ flag_ptrs, int world, int32_t* __restrict__ epoch) {
  const int e = epoch[0] + 1;
  for (int i = threadIdx.x; i < k * world; i += blockDim.x) {
    const int s = i / k, j = i - s * k;
    reinterpret_cast<int32_t*>(payload_ptrs[s])[j] = payload[j];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world) {
    volatile int32_t* f = reinterpret_cast<volatile int32_t*>(flag_ptrs[threadIdx.x]);
    *f = e;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) epoch[0] = e;
}

// spin until all `world` local flag slots reached epoch[0]; bounded: on time-out raise err and return
__global__ void peer_wait_kernel(const int32_t* __restrict__ flags, int world, const int32_t* __restrict__ epoch,
                                 int32_t* __restrict__ err, const int32_t* __restrict__ limit_log2,
                                 long long max_cycles) {
  if (threadIdx.x >= world) return;
  if (limit_log2) max_cycles = 1ll << limit_log2[0];
  const int e = epoch[0];
  const volatile int32_t* f = flags + threadIdx.x;
  const long long t0 = clock64();
  while (*f < e) {
    if (clock64() - t0 > max_cycles) {
      atomicOr(err, 1);
      break;
    }
    __nanosleep(200);
  }
  __threadfence_system();
}


// Sum of the G ranks' buffers over peer memory, result delivered to every rank (the DenseLayer gradient all-reduce of
// the data-parallel step, inside the step's CUDA graph): rank r owns the r-th slice; for every 16-byte chunk of it the
// G sources are loaded (G - 1 of them over NVLink), added in RANK ORDER — the same order on every rank, so the
// replicas stay bit-identical and the result does not depend on timing — and the sum is stored into all G
// destination buffers.  Callers bracket it with signal / wait pairs: sources complete before, sums delivered after.
template <int UNROLL>
__global__ void __launch_bounds__(256)
peer_allreduce_kernel(const int64_t* __restrict__ src_ptrs, const int64_t* __restrict__ dst_ptrs,
                      const int32_t* __restrict__ ctrl, int64_t n) {
  __shared__ const float* s_src[kPeerMaxRanks];
  __shared__ float* s_dst[kPeerMaxRanks];
  const int me = ctrl[0], g = ctrl[1];
  if (threadIdx.x < g) {
    s_src[threadIdx.x] = reinterpret_cast<const float*>(src_ptrs[threadIdx.x]);
    s_dst[threadIdx.x] = reinterpret_cast<float*>(dst_ptrs[threadIdx.x]);
  }
  __syncthreads();
  const int64_t n4 = n / 4;
  const int64_t per = (n4 + g - 1) / g;
  const int64_t lo = min(n4, (int64_t)me * per), hi = min(n4, lo + per);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < hi; i0 += UNROLL * stride) {
    float4 acc[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < g; ++s) {                    // rank order; the UNROLL loads of a source are in flight together
      float4 v[UNROLL];
#pragma unroll
      for (int k = 0; k < UNROLL; ++k) {
        const int64_t i = i0 + k * stride;
        v[k] = i < hi ? ld_volatile_f4(reinterpret_cast<const float4*>(s_src[s]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < UNROLL; ++k) {
        acc[k].x += v[k].x; acc[k].y += v[k].y; acc[k].z += v[k].z; acc[k].w += v[k].w;
      }
    }
    for (int s = 0; s < g; ++s) {
#pragma unroll
      for (int k = 0; k < UNROLL; ++k) {
        const int64_t i = i0 + k * stride;
        if (i < hi) reinterpret_cast<float4*>(s_dst[s])[i] = acc[k];
      }
    }
  }
  // the n % 4 tail: the last rank, scalar
  if (me == g - 1 && blockIdx.x == 0 && threadIdx.x < (int)(n - n4 * 4)) {
    const int64_t i = n4 * 4 + threadIdx.x;
    float acc = 0.f;
    for (int s = 0; s < g; ++s) acc += *reinterpret_cast<const volatile float*>(s_src[s] + i);
    for (int s = 0; s < g; ++s) s_dst[s][i] = acc;
  }
}


// buf[idx[0], :] = 0 when 0 <= idx[0] < rows.  The bounded dedup gives every out-of-range id the inverse index U (the
// unique count): zeroing row U of the landing buffers makes those lookups read the zero row the unsharded gather
// returns for them, whatever an earlier step left there.
__global__ void zero_row_kernel(float* __restrict__ buf, int64_t rows, int64_t width, const int32_t* __restrict__ idx) {
  const int64_t r = idx[0];
  if (r < 0 || r >= rows) return;
  for (int64_t c = threadIdx.x; c < width; c += blockDim.x) buf[r * width + c] = 0.f;
}

}  // namespace mrec

// in : bounds_all[G*(G+1)] i32, ctrl[2] i32 {rank, world}     out: dst_off[G], src_off[G+1], inbox_off[G], n_r[1] (i32)
MREC_API int mrec_shard_offsets(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                                void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  MREC_CHECK_NPARAM(a, 6);
  for (int i = 0; i < 6; ++i) MREC_REQUIRE(a.is_i32(i), ERR_DTYPE, "mrec_shard_offsets: all params are int32");
  MREC_REQUIRE(a.numel(1) >= 2, ERR_SHAPE, "mrec_shard_offsets: ctrl[2] expected");
  MREC_LAUNCH(shard_offsets_kernel, 1, 32, 0, a.stream, a.ptr<int32_t>(0), a.ptr<int32_t>(1), a.ptr<int32_t>(2),
              a.ptr<int32_t>(3), a.ptr<int32_t>(4), a.ptr<int32_t>(5));
  return check_launch("shard_offsets");
}

// in : rows[U_cap, W] f32|i32|i64 (first B[me][G] rows valid), my_bounds[G+1] i32, inbox_off[G] i32, peer_ptrs[G] i64,
//      cap_like[cap_rows, 0..] (shape carrier: capacity of every inbox in rows), mod_like[M, 0..] (M > 0 and rows
//      int32: send rows % M; numel-0 tensor with M = 0 rows: no transform)         out: err[1] i32 (bit 1 = overflow)
MREC_API int mrec_push_rows_to_peers(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                                     void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  if (a.nparam != 7) return fail(ERR_NPARAM, "mrec_push_rows_to_peers: expected 7 params, got %d", a.nparam);
  const bool wide = a.is_i64(0);
  MREC_REQUIRE(a.is_f32(0) || a.is_i32(0) || wide, ERR_DTYPE, "mrec_push_rows_to_peers: rows must be f32|i32|i64");
  MREC_REQUIRE(a.is_i32(1) && a.is_i32(2) && a.is_i64(3) && a.is_i32(6), ERR_DTYPE,
               "mrec_push_rows_to_peers: bounds/inbox_off/err int32, peer_ptrs int64");
  const int world = (int)a.numel(3);
  MREC_REQUIRE(world >= 1 && world <= kPeerMaxRanks && a.numel(1) >= world + 1 && a.numel(2) >= world, ERR_SHAPE,
               "mrec_push_rows_to_peers: G out of range or bounds/inbox_off too short");
  const int64_t n = a.ndims[0] >= 2 ? a.dim(0, 0) : a.numel(0);
  const int width = a.ndims[0] >= 2 ? (int)(a.numel(0) / (n > 0 ? n : 1)) : 1;
  const int64_t cap_rows = a.dim(4, 0);
  const int64_t mod = a.dim(5, 0);
  MREC_REQUIRE(mod == 0 || ((a.is_i32(0) || wide) && width == 1), ERR_DTYPE,
               "mrec_push_rows_to_peers: the modulo transform is for integer keys");
  if (n == 0) return OK;
  if (wide) {
    MREC_LAUNCH(push_rows_to_peers_kernel<int64_t>, grid_for(cdiv(n * width, 1024), 8), 256, 0, a.stream, a.ptr<int64_t>(0),
                width, a.ptr<int32_t>(1), a.ptr<int32_t>(2), a.ptr<int64_t>(3), world, mod, cap_rows, a.ptr<int32_t>(6));
  } else if (a.is_i32(0)) {
    MREC_LAUNCH(push_rows_to_peers_kernel<int32_t>, grid_for(cdiv(n * width, 1024), 8), 256, 0, a.stream, a.ptr<int32_t>(0),
                width, a.ptr<int32_t>(1), a.ptr<int32_t>(2), a.ptr<int64_t>(3), world, mod, cap_rows, a.ptr<int32_t>(6));
  } else if (width % 4 == 0 && reinterpret_cast<uintptr_t>(a.params[0]) % 16 == 0) {
    // float rows of a multiple of 16 bytes: one 16-byte NVLink store per thread (the inboxes are 256-byte aligned
    // cudaMalloc allocations, so row starts stay 16-byte aligned)
    MREC_LAUNCH(push_rows_to_peers_kernel<uint4>, grid_for(cdiv(n * (width / 4), 1024), 8), 256, 0, a.stream, a.ptr<uint4>(0),
                width / 4, a.ptr<int32_t>(1), a.ptr<int32_t>(2), a.ptr<int64_t>(3), world, (int64_t)0, cap_rows, a.ptr<int32_t>(6));
  } else {
    MREC_LAUNCH(push_rows_to_peers_kernel<uint32_t>, grid_for(cdiv(n * width, 1024), 8), 256, 0, a.stream, a.ptr<uint32_t>(0),
                width, a.ptr<int32_t>(1), a.ptr<int32_t>(2), a.ptr<int64_t>(3), world, (int64_t)0, cap_rows, a.ptr<int32_t>(6));
  }
  return check_launch("push_rows_to_peers");
}

// in : idx[1] i32        out: buf[R, W] f32 — row idx[0] is zeroed when it lies inside the buffer
MREC_API int mrec_zero_row(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes, void* stream,
                           void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  if (a.nparam != 2) return fail(ERR_NPARAM, "mrec_zero_row: expected 2 params, got %d", a.nparam);
  MREC_REQUIRE(a.is_i32(0) && a.is_f32(1) && a.numel(0) >= 1 && a.ndims[1] >= 1, ERR_DTYPE, "mrec_zero_row: idx[1] i32, buf[R,W] f32");
  const int64_t rows = a.dim(1, 0);
  if (rows == 0) return OK;
  const int64_t width = a.numel(1) / rows;
  MREC_LAUNCH(zero_row_kernel, 1, 128, 0, a.stream, a.ptr<float>(1), rows, width, a.ptr<int32_t>(0));
  return check_launch("zero_row");
}

// in : src_ptrs[G] i64 (every rank's source buffer [n] f32 as mapped in THIS process), dst_ptrs[G] i64 (every rank's
//      destination buffer [n] f32), ctrl[2] i32 {rank, world}          out: dst[n] f32 (this rank's destination: carries n)
// Every rank calls it between a signal / wait pair (sources complete) and another (sums delivered).
MREC_API int mrec_peer_allreduce(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                                 void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  if (a.nparam != 4) return fail(ERR_NPARAM, "mrec_peer_allreduce: expected 4 params, got %d", a.nparam);
  MREC_REQUIRE(a.is_i64(0) && a.is_i64(1) && a.is_i32(2) && a.is_f32(3), ERR_DTYPE,
               "mrec_peer_allreduce: src_ptrs/dst_ptrs i64, ctrl i32, dst f32");
  const int world = (int)a.numel(0);
  MREC_REQUIRE(world >= 1 && world <= kPeerMaxRanks && a.numel(1) == world && a.numel(2) >= 2, ERR_SHAPE,
               "mrec_peer_allreduce: 1 <= G <= %d, dst_ptrs[G], ctrl[2]", kPeerMaxRanks);
  MREC_REQUIRE(a.aligned(3, 16), ERR_ALIGN, "mrec_peer_allreduce: buffers must be 16-byte aligned");
  const int64_t n = a.numel(3);
  if (n == 0) return OK;
  const int64_t per = cdiv(cdiv(n, 4), world);
  MREC_LAUNCH(peer_allreduce_kernel<4>, grid_for(cdiv(per, 1024), 4), 256, 0, a.stream, a.ptr<int64_t>(0), a.ptr<int64_t>(1),
              a.ptr<int32_t>(2), n);
  return check_launch("peer_allreduce");
}

// in : payload[K] i32 (K may be 0), payload_ptrs[G] i64, flag_ptrs[G] i64, epoch[1] i32 (incremented)   out: dummy[1]
MREC_API int mrec_peer_signal(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                              void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  if (a.nparam != 5) return fail(ERR_NPARAM, "mrec_peer_signal: expected 5 params, got %d", a.nparam);
  MREC_REQUIRE(a.is_i32(0) && a.is_i64(1) && a.is_i64(2) && a.is_i32(3), ERR_DTYPE, "mrec_peer_signal: dtype mismatch");
  const int world = (int)a.numel(2);
  MREC_REQUIRE(world >= 1 && world <= kPeerMaxRanks, ERR_SHAPE, "mrec_peer_signal: G out of range");
  MREC_LAUNCH(peer_signal_kernel, 1, 256, 0, a.stream, a.ptr<int32_t>(0), (int)a.numel(0), a.ptr<int64_t>(1),
              a.ptr<int64_t>(2), world, a.ptr<int32_t>(3));
  return check_launch("peer_signal");
}

// in : flags[G] i32 (local slots written by the peers), epoch[1] i32, optional limit_log2[1] i32 (spin limit in
//      cycles = 2^limit; default ~4 s)                                       out: err[1] i32 (bit 0 = time-out)
MREC_API int mrec_peer_wait(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                            void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  if (a.nparam != 3 && a.nparam != 4) return fail(ERR_NPARAM, "mrec_peer_wait: expected 3 or 4 params, got %d", a.nparam);
  const int o = a.nparam - 1;                       // err is the last param
  MREC_REQUIRE(a.is_i32(0) && a.is_i32(1) && a.is_i32(o) && a.is_i32(2), ERR_DTYPE, "mrec_peer_wait: int32 expected");
  const int world = (int)a.numel(0);
  MREC_REQUIRE(world >= 1 && world <= kPeerMaxRanks, ERR_SHAPE, "mrec_peer_wait: G out of range");
  // ~17 s at 2 GHz (first steps on a cold box can be seconds apart between ranks): a dead peer raises err instead
  // of hanging the GPU
  MREC_LAUNCH(peer_wait_kernel, 1, 32, 0, a.stream, a.ptr<int32_t>(0), world, a.ptr<int32_t>(1), a.ptr<int32_t>(o),
              a.nparam == 4 ? a.ptr<int32_t>(2) : (int32_t*)nullptr, 1ll << 35);
  return check_launch("peer_wait");
}

// in : keys[N] i32|i64, owners_like[G, ..] (dim 0 = ranks), bits_like[B, ..] (dim 0 = key bits, B + log2 G < 63)
// out: keys'[N] i64 = owner << B | key
MREC_API int mrec_shard_remap_hash(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                                   void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  if (a.nparam != 4) return fail(ERR_NPARAM, "mrec_shard_remap_hash: expected 4 params, got %d", a.nparam);
  MREC_REQUIRE((a.is_i32(0) || a.is_i64(0)) && a.is_i64(3), ERR_DTYPE, "mrec_shard_remap_hash: keys i32|i64, out i64");
  MREC_REQUIRE(a.ndims[1] >= 1 && a.ndims[2] >= 1, ERR_SHAPE, "mrec_shard_remap_hash: owners_like[G,..], bits_like[B,..]");
  const int64_t n = a.numel(0);
  const int world = (int)a.dim(1, 0), bits = (int)a.dim(2, 0);
  MREC_REQUIRE(a.numel(3) == n, ERR_SHAPE, "mrec_shard_remap_hash: out must match keys");
  MREC_REQUIRE(world >= 1 && world <= kPeerMaxRanks && bits >= 1 && bits <= 56, ERR_SHAPE,
               "mrec_shard_remap_hash: G in [1, 64], key bits in [1, 56]");
  if (n == 0) return OK;
  if (!a.params[0] || !a.params[3]) return fail(ERR_NULL, "mrec_shard_remap_hash: null keys/out");
  if (a.is_i32(0))
    MREC_LAUNCH(shard_remap_hash_kernel<int32_t>, grid_for(cdiv(n, 256), 8), 256, 0, a.stream, a.ptr<int32_t>(0),
                a.ptr<int64_t>(3), n, world, bits);
  else
    MREC_LAUNCH(shard_remap_hash_kernel<int64_t>, grid_for(cdiv(n, 256), 8), 256, 0, a.stream, a.ptr<int64_t>(0),
                a.ptr<int64_t>(3), n, world, bits);
  return check_launch("shard_remap_hash");
}

// in : n_valid[1] i32, value[1] (dtype of buf)     out: buf[cap] i32|i64 — buf[i] = value for i >= n_valid
MREC_API int mrec_fill_tail(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                            void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  if (a.nparam != 3) return fail(ERR_NPARAM, "mrec_fill_tail: expected 3 params, got %d", a.nparam);
  MREC_REQUIRE(a.is_i32(0) && (a.is_i32(2) || a.is_i64(2)) && strcmp(a.dtypes[1], a.dtypes[2]) == 0, ERR_DTYPE,
               "mrec_fill_tail: n_valid int32, value and buf the same integer dtype");
  MREC_REQUIRE(a.numel(0) >= 1 && a.numel(1) >= 1, ERR_SHAPE, "mrec_fill_tail: n_valid[1], value[1]");
  const int64_t n = a.numel(2);
  if (n == 0) return OK;
  if (a.is_i32(2))
    MREC_LAUNCH(fill_tail_kernel<int32_t>, grid_for(cdiv(n, 256), 4), 256, 0, a.stream, a.ptr<int32_t>(2), n, a.ptr<int32_t>(0),
                a.ptr<int32_t>(1));
  else
    MREC_LAUNCH(fill_tail_kernel<int64_t>, grid_for(cdiv(n, 256), 4), 256, 0, a.stream, a.ptr<int64_t>(2), n, a.ptr<int32_t>(0),
                a.ptr<int64_t>(1));
  return check_launch("fill_tail");
}
