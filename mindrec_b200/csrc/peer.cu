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

namespace mrec {

constexpr int kPeerMaxRanks = 16;

struct PeerTable {
  float* ptr[kPeerMaxRanks];
  int32_t dst_off[kPeerMaxRanks];
  int32_t src_off[kPeerMaxRanks + 1];
};

template <typename Vec>
__global__ void __launch_bounds__(256)
gather_to_peers_kernel(const float* __restrict__ table, const int32_t* __restrict__ rows, int64_t n_rows, int cpr,
                       int64_t vocab, int world, const int64_t* __restrict__ peer_ptrs,
                       const int32_t* __restrict__ dst_off, const int32_t* __restrict__ src_off) {
  __shared__ PeerTable s_t;
  if (threadIdx.x < world) {
    s_t.ptr[threadIdx.x] = reinterpret_cast<float*>(peer_ptrs[threadIdx.x]);
    s_t.dst_off[threadIdx.x] = dst_off[threadIdx.x];
  }
  if (threadIdx.x <= world) s_t.src_off[threadIdx.x] = src_off[threadIdx.x];
  __syncthreads();
  const int64_t total = n_rows * cpr;
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
        int s = 0;
        while (s + 1 < world && i >= s_t.src_off[s + 1]) ++s;   // G <= 16: a short scan of the bucket edges
        const int64_t row = rows[i];
        dst[k] = reinterpret_cast<Vec*>(s_t.ptr[s]) + ((int64_t)s_t.dst_off[s] + (i - s_t.src_off[s])) * cpr + c;
        if ((uint64_t)row < (uint64_t)vocab) v[k] = reinterpret_cast<const Vec*>(table)[row * cpr + c];
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (dst[k]) *dst[k] = v[k];   // st.global on a peer-mapped address: goes out over NVLink
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
// out: dummy[1] i32
MREC_API int mrec_gather_to_peers(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                                  void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  MREC_CHECK_NPARAM(a, 6);
  MREC_REQUIRE(a.is_f32(0) && a.is_i32(1) && a.is_i64(2) && a.is_i32(3) && a.is_i32(4), ERR_DTYPE,
               "mrec_gather_to_peers: table f32, rows i32, peer_ptrs i64, dst_off/src_off i32");
  const int64_t vocab = a.dim(0, 0);
  const int dim = a.ndims[0] >= 2 ? (int)a.dim(0, 1) : 1;
  const int world = (int)a.numel(2);
  MREC_REQUIRE(world >= 1 && world <= kPeerMaxRanks, ERR_SHAPE, "mrec_gather_to_peers: 1 <= G <= %d", kPeerMaxRanks);
  MREC_REQUIRE(a.numel(3) >= world && a.numel(4) >= world + 1, ERR_SHAPE, "mrec_gather_to_peers: dst_off[G], src_off[G+1]");
  const int64_t n = a.numel(1);
  if (n == 0) return OK;
  if (dim % 4 == 0) {
    MREC_REQUIRE(a.aligned(0, 16), ERR_ALIGN, "mrec_gather_to_peers: table must be 16-byte aligned");
    const int cpr = dim / 4;
    MREC_LAUNCH(gather_to_peers_kernel<float4>, grid_for(cdiv(n * cpr, 1024), 8), 256, 0, a.stream, a.ptr<float>(0),
                a.ptr<int32_t>(1), n, cpr, vocab, world, a.ptr<int64_t>(2), a.ptr<int32_t>(3), a.ptr<int32_t>(4));
  } else {
    MREC_LAUNCH(gather_to_peers_kernel<float>, grid_for(cdiv(n * dim, 1024), 8), 256, 0, a.stream, a.ptr<float>(0),
                a.ptr<int32_t>(1), n, dim, vocab, world, a.ptr<int64_t>(2), a.ptr<int32_t>(3), a.ptr<int32_t>(4));
  }
  return check_launch("gather_to_peers");
}
