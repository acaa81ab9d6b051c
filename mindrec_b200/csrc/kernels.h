// Internal C++ interfaces shared between the .cu files of libmindrec_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace mrec {

// unique.cu --------------------------------------------------------------------------------------
size_t unique_workspace_bytes(int64_t n, int key_bytes);
size_t unique_first_workspace_bytes(int64_t n, int key_bytes);
int key_bits_for_bound(uint64_t bound);
// bound > 0: keys outside [0, bound) are collapsed onto `bound` and only the low bits are sorted.
template <typename KeyT>
int unique_sorted(const KeyT* ids, int64_t n, uint64_t bound, KeyT* uniq, int32_t* inverse,
                  int32_t* count, int32_t* perm, int32_t* seg_start, int32_t* seg_of, void* ws,
                  size_t ws_bytes, cudaStream_t stream, const int32_t* n_valid = nullptr);
template <typename KeyT>
int unique_first(const KeyT* ids, int64_t n, uint64_t bound, KeyT* uniq, int32_t* inverse,
                 int32_t* count, void* ws, size_t ws_bytes, cudaStream_t stream);

// sparse_opt.cu ----------------------------------------------------------------------------------
size_t sparse_opt_workspace_bytes(int64_t n, int dim);
size_t segment_sum_workspace_bytes(int64_t n, int dim);

}  // namespace mrec
