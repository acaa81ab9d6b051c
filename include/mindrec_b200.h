/* libmindrec_b200.so — C-ABI of the B200-native MindRec embedding-and-interaction hot path.
 *
 * Every compute entry point has the signature MindSpore's ops.Custom(func_type="aot") binds
 * (mindspore/ops/operations/custom_ops.py, "aot" section [upstream]):
 *
 *     int f(int nparam, void **params, int *ndims, int64_t **shapes,
 *           const char **dtypes, void *stream, void *extra);
 *
 *   params  inputs, then outputs (workspaces are declared as trailing uint8 outputs) — DEVICE pointers
 *   ndims / shapes / dtypes   per-param rank, extents and MindSpore dtype name ("float32","int32","int64",...)
 *   stream  cudaStream_t the kernels are enqueued on; the call never synchronises, allocates or frees
 *   extra   ignored (all scalars travel in small device tensors, so no attr / AotExtra C++ ABI is needed)
 *   return  0 on success, else one of MREC_ERR_* (MindSpore raises RuntimeError); text in mrec_last_error()
 *
 * Persistent state (tables, optimizer moments, hash slots) is passed as INPUTS and mutated in place;
 * such ops return a dummy int32[1] output.  Data-dependent sizes are returned as device scalars
 * (count[1]) with outputs padded to their static maximum.
 *
 * The reference interface each symbol replaces is cited as /root/reference/<file>:<line>.
 */
#ifndef MINDREC_B200_H_
#define MINDREC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MREC_OK 0
#define MREC_ERR_NPARAM 1    /* wrong number of params */
#define MREC_ERR_DTYPE 2     /* dtype string mismatch */
#define MREC_ERR_SHAPE 3     /* rank / extent mismatch */
#define MREC_ERR_ALIGN 4     /* pointer not 16-byte aligned where vector access needs it */
#define MREC_ERR_DIM 5       /* unsupported embedding dim */
#define MREC_ERR_CUDA 6      /* cudaPeekAtLastError() after launch */
#define MREC_ERR_WORKSPACE 7 /* workspace too small (see *_workspace_bytes) */
#define MREC_ERR_NULL 8      /* null pointer for a non-empty param */

#define MREC_AOT_ARGS                                                                      \
  int nparam, void **params, int *ndims, int64_t **shapes, const char **dtypes, void *stream, \
      void *extra

/* ---- helpers (host only, not aot) ------------------------------------------------------------ */
const char *mrec_version(void);
const char *mrec_last_error(void);            /* thread-local text of the last non-zero return */
unsigned long long mrec_launch_count(void);   /* kernels launched by this library in this process */

/* ---- K1 gather -------------------------------------------------------------------------------
 * Replaces nn.EmbeddingLookup / P.Gather(table, ids, 0):
 *   models/wide_deep/src/wide_and_deep.py:277-290,300-302; models/deepfm/src/deepfm.py:217,221;
 *   models/deep_and_cross/src/deep_and_cross.py:199; mindspore_rec/ops/embedding.py:194.
 * Out-of-range ids -> zero row (GPU Gather) and the optional oob flag is set.
 *   in : table[V,D] f32, ids[...] i32|i64            out: out[N*D] f32, (oob[1] i32)            */
int mrec_gather(MREC_AOT_ARGS);
/* gather + Mul(mask) + Reshape fused (wide_and_deep.py:303,308-309; deepfm.py:215,222,230;
 * deep_and_cross.py:295-298):
 *   in : table[V,D], ids[B,F], mask[B,F] f32         out: out[B,F*D] f32, (oob[1])              */
int mrec_gather_masked(MREC_AOT_ARGS);
/* dim-1 gather + Mul(mask) + ReduceSum(axis 1) + bias (wide_and_deep.py:300,305-306;
 * deepfm.py:217-219):
 *   in : table[V]|[V,1], ids[B,F], mask[B,F], bias[1] out: out[B]|[B,1] f32, (oob[1])           */
int mrec_gather_reduce(MREC_AOT_ARGS);
/* multi-hot pooled lookup: gather + Mul(mask) + ReduceMean over the S slots (masked slots count in the mean),
 * models/wide_and_deep_multitable/src/wide_and_deep.py:301-346.  The backward is mrec_sparse_* with the pooled
 * gradient broadcast over the slots (g[B,D] for N = B*S positions) and mask / S as the weight.
 *   in : table[V,D] (D % 4 == 0), ids[B,S], mask[B,S] f32                     out: out[B,D] f32, (oob[1])   */
int mrec_gather_pool(MREC_AOT_ARGS);

/* ---- K2 unique -------------------------------------------------------------------------------
 * Replaces P.Unique (mindspore_rec/ops/embedding.py:192) and the optimizer-side RowTensor dedup
 * (nn.Optimizer, reached from wide_and_deep.py:420-445).  Ascending order = upstream GPU kernel.
 *   in : ids[N] i32|i64
 *   out: uniq[N] (ids dtype, padded), inverse[N] i32, count[1] i32, perm[N] i32 (stable sort
 *        permutation), seg_start[N+1] i32, seg_of[N] i32, workspace[mrec_unique_workspace_bytes] u8 */
int mrec_unique(MREC_AOT_ARGS);
/* Same, second input table_like[V,...]: only ceil(log2(V+1)) key bits are sorted; ids outside [0,V)
 * collapse onto the value V (they form the last segment, which the optimizers skip).  Optional third input
 * n_valid[1] i32 (device-side count): only ids[0 .. n_valid) exist — the dedup works on that prefix (its cost
 * follows n_valid, not N), seg_start[count] = n_valid, outputs past the prefix are not written.           */
int mrec_unique_bounded(MREC_AOT_ARGS);
/* First-occurrence order = upstream CPU Unique kernel (BASELINE config 1 runs device_target=CPU).
 *   in : ids[N]   out: uniq[N], inverse[N] i32, count[1] i32, workspace[mrec_unique_first_workspace_bytes] */
int mrec_unique_first(MREC_AOT_ARGS);
/* Shard bucketing (row-sharded tables, replaces the AllGather/ReduceScatter pair auto-parallel inserts for
 * nn.EmbeddingLookup(slice_mode=TABLE_ROW_SLICE), wide_and_deep.py:234-249): with keys remapped owner-major the
 * per-owner runs of the unique keys are found by binary search on the device.
 *   in : uniq[N] (ascending; first count entries valid), count[1] i32, edges[E] (uniq dtype)   out: bounds[E] i32 */
int mrec_shard_bounds(MREC_AOT_ARGS);
/* Owner-major key remap, key' = (key mod G) * R + key div G (out-of-range -> G*R):
 *   in : ids[...] i32|i64, table_like[V,...], owners_like[G,R] (only the shapes are read)   out: keys[...] */
int mrec_shard_remap(MREC_AOT_ARGS);
/* Fused owner-side gather + NVLink peer store (forward row exchange of row-sharded tables): every requested row
 * is read once and stored directly into the requesting rank's landing buffer through a peer-mapped pointer.
 *   in : table[R,D] f32 (local shard), rows[n_r] i32 (local row ids, concatenated by source rank),
 *        peer_ptrs[G] i64 (landing-buffer base of every rank as mapped in this process), dst_off[G] i32,
 *        src_off[G+1] i32                                                                      out: dummy[1] */
int mrec_gather_to_peers(MREC_AOT_ARGS);
/* Split serve: with two more inputs — dirty[ceil(R/32)] i32 (bit r set = row r is rewritten by the current step's update,
 * filled by mrec_bitmap_set from the owner-side dedup) and mode_like[mode,..] — mrec_gather_to_peers serves only the clean
 * rows (mode 1: one step early, underneath the DenseLayers) or only the dirty rows (mode 2: after the update).
 *   mrec_bitmap_set   in : rows[N] i32, count[1] i32        out: bitmap[W] i32 (bits OR-ed in; the caller clears) */
int mrec_bitmap_set(MREC_AOT_ARGS);
/* Device-driven exchange (no host-side sizes; the whole sharded step can be one CUDA graph).  With
 * B[s][o] = start of rank s's bucket for owner o among its sorted unique keys (B[s][G] = U_s):
 *   mrec_shard_offsets       in : bounds_all[G*(G+1)] i32, ctrl[2] i32 {rank, G}
 *                            out: dst_off[G], src_off[G+1], inbox_off[G], n_r[1]  (all i32)
 *   mrec_push_rows_to_peers  in : rows[cap,W] f32|i32|i64, my_bounds[G+1], inbox_off[G], peer_ptrs[G] i64,
 *                                 cap_like[cap_rows,..], mod_like[M,..] (M > 0: integer keys are sent as key % M)
 *                            out: err[1] i32 (bit 1: an inbox overflowed)
 *   mrec_peer_signal         in : payload[K] i32, payload_ptrs[G] i64, flag_ptrs[G] i64, epoch[1] i32 (+1)  out: dummy[1]
 *   mrec_peer_wait           in : flags[G] i32, epoch[1] i32 [, limit_log2[1] i32: spin limit 2^limit cycles, default 2^35 (~17 s)]
 *                            out: err[1] i32 (bit 0: time-out) */
int mrec_shard_offsets(MREC_AOT_ARGS);
/* Hash-table (MapParameter) sharding, SURVEY 8e: owner = hash(key) mod G, every rank owns an independent table.
 *   mrec_shard_remap_hash    in : keys[N] i32|i64, owners_like[G,..], bits_like[B,..] (keys live in [0, 2^B))
 *                            out: keys'[N] i64 = owner << B | key (reserved / out-of-range keys -> G << B, dropped
 *                                 by mrec_unique_bounded with bound G << B); the owner gets key' % 2^B back
 *   mrec_fill_tail           in : n_valid[1] i32, value[1]            out: buf[cap] i32|i64 (buf[i >= n_valid] = value:
 *                                 blanks the stale tail of a static key inbox with the reserved key -1) */
int mrec_shard_remap_hash(MREC_AOT_ARGS);
int mrec_fill_tail(MREC_AOT_ARGS);
int mrec_push_rows_to_peers(MREC_AOT_ARGS);
int mrec_peer_signal(MREC_AOT_ARGS);
int mrec_peer_wait(MREC_AOT_ARGS);
/* Sum of the G ranks' buffers over peer memory, delivered to every rank: the mean all-reduce of the data-parallel
 * DenseLayer gradients (DistributedGradReducer, wide_and_deep.py:455-470; the 1/G is folded into the optimizer's
 * gradient scale) as a node of the step's CUDA graph.  Rank r sums the r-th slice in rank order (deterministic,
 * replicas bit-identical) and stores it into all G destinations; callers bracket it with signal / wait pairs.
 *   in : src_ptrs[G] i64, dst_ptrs[G] i64 (every rank's [n] f32 buffers as mapped here), ctrl[2] i32 {rank, G}
 *   out: dst[n] f32 (this rank's destination buffer) */
int mrec_peer_allreduce(MREC_AOT_ARGS);
/* buf[idx[0], :] = 0 when the index lies inside buf (the landing-buffer row that out-of-range ids expand from: they read
 * the zero row, as nn.EmbeddingLookup's gather returns for them).   in : idx[1] i32   out: buf[R,W] f32 */
int mrec_zero_row(MREC_AOT_ARGS);
/* CUDA-IPC plumbing for the peer buffers (host only, set-up time, not aot) */
void *mrec_peer_alloc(size_t bytes);
int mrec_peer_free(void *p);
int mrec_ipc_get_handle(void *base_ptr, char *handle64);
void *mrec_ipc_open_handle(const char *handle64);
int mrec_ipc_close_handle(void *p);
size_t mrec_unique_workspace_bytes(int64_t n, int key_bytes);
size_t mrec_unique_first_workspace_bytes(int64_t n, int key_bytes);

/* ---- K3-K5 deterministic segment-sum fused with sparse optimizers ------------------------------
 * Replaces Gather-bprop RowTensor -> Unique -> UnsortedSegmentSum -> LazyAdam / FTRL:
 *   wide_and_deep.py:420-430,479-492 (LazyAdam lr 3.5e-4 eps 1e-8; FTRL lr 5e-2 l1=l2=1e-8 accum 1.0);
 *   models/wide_and_deep_multitable/src/wide_and_deep.py:525-535.
 * g[N/div, D] holds one gradient row per `div` consecutive lookup positions (div = F for the wide
 * logit gradient, 1 for the deep input gradient); mask[N] is the Mul(mask) bprop factor (numel 0 = none).
 * Hyper blocks are f32[16] device tensors:
 *   Adam : lr, beta1, beta2, eps, beta1_power, beta2_power, lr_t, 1/loss_scale, l2 (dense-mode table
 *          regulariser, wide_and_deep.py:359-360), 7 reserved
 *   FTRL : lr, l1, l2, lr_power, 1/loss_scale, 11 reserved
 *   in : w[V,D] m[V,D] v[V,D] hyper[16] g mask uniq[N] perm[N] seg_start[N+1] seg_of[N], (n_valid[1] i32:
 *        only the first n_valid sorted positions are real — statically sized inboxes of the sharded path)
 *   out: dummy[1] i32, workspace[mrec_sparse_opt_workspace_bytes(N, D)] u8                       */
int mrec_sparse_lazy_adam(MREC_AOT_ARGS);
/*   in : w[V,D] accum[V,D] linear[V,D] hyper[16] g mask uniq perm seg_start seg_of   out: dummy, workspace */
int mrec_sparse_ftrl(MREC_AOT_ARGS);
/* Stand-alone UnsortedSegmentSum in sorted-segment order (no atomics, bit-reproducible):
 *   in : g mask perm seg_start seg_of     out: gsum[N,D] f32 (rows >= count untouched),
 *        workspace[mrec_segment_sum_workspace_bytes(N, D)]                                        */
int mrec_segment_sum(MREC_AOT_ARGS);
/* Dense gradient of a non-sparse Gather (bprop = UnsortedSegmentSum into [V,D]; the multitable model,
 * models/wide_and_deep_multitable/src/wide_and_deep.py:291-346, looks one table up from several inputs):
 *   in : g[N/div,D] f32|f16, mask[N|0], uniq[N], perm[N], seg_start[N+1], seg_of[N]
 *   out: table[V,D] f32 with table[uniq[u]] += segment sum (accumulates over calls; no atomics), workspace */
int mrec_segment_sum_scatter_add(MREC_AOT_ARGS);
/* Fused UnsortedSegmentSum + gradient push of row-sharded tables (the backward all-to-all of SURVEY 8e): the sum of
 * segment u is stored straight into the inbox of the rank that owns u's key, through its peer-mapped pointer — no local
 * gsum buffer, no second pass.  With keys owner-major, segment u belongs to owner o = #{r >= 1 : u >= my_bounds[r]} and
 * lands at row inbox_off[o] + u - my_bounds[o] (segments >= my_bounds[G], i.e. out-of-range ids, are dropped).
 *   in : g[N/div,D] f32|f16 (D % 4 == 0), mask[N|0], perm[N], seg_start[N+1], seg_of[N], my_bounds[G+1] i32,
 *        inbox_off[G] i32, peer_ptrs[G] i64, cap_like[cap_rows,..]
 *   out: err[1] i32 (bit 1: an inbox overflowed; the row is dropped), workspace[mrec_segment_sum_workspace_bytes(N, D)] */
int mrec_segment_sum_to_peers(MREC_AOT_ARGS);
/* nn.Adam (not Lazy) with a RowTensor gradient = dense-equivalent update of the WHOLE table (every
 * row's moments decay; wide_and_deep.py:435-437 when sparse=True on one device):
 *   in : w m v hyper[16] g mask uniq perm seg_start seg_of row_flags[V] u8 (zero on entry and exit)
 *   out: dummy[1], workspace                                                                       */
int mrec_adam_rowsparse(MREC_AOT_ARGS);
size_t mrec_sparse_opt_workspace_bytes(int64_t n, int dim);
size_t mrec_segment_sum_workspace_bytes(int64_t n, int dim);
/* nn.Adam preamble: beta powers advance, lr_t = lr*sqrt(1-b2^t)/(1-b1^t).  in: hyper[16]  out: dummy[1] */
int mrec_adam_begin_step(MREC_AOT_ARGS);
/* nn.Adam dense kernel (MLP weights, Wide_b: wide_and_deep.py:405-413,435-437).
 *   in : w m v hyper[16] g (same numel)   out: dummy[1]                                            */
int mrec_adam_dense(MREC_AOT_ARGS);
/* nn.FTRL dense kernel (ApplyFtrl).  in : w accum linear hyper[16] g   out: dummy[1]               */
int mrec_ftrl_dense(MREC_AOT_ARGS);

/* ---- a8 loss ------------------------------------------------------------------------------------------
 * SigmoidCrossEntropyWithLogits + ReduceMean and the bprop seed (wide_and_deep.py:315,354-355,479-486;
 * deepfm.py:254-255; deep_and_cross.py:323-325), one launch, deterministic reduction:
 *   in : a[B] f32, b[B] f32 | numel 0 (logit = a + b), label[B] f32, sens[1] f32
 *   out: logit[B], loss[1] (mean), delta[B] = sens*(sigmoid(logit)-label)/B, delta16[B] f16 | numel 0, delta_sum[1] */
int mrec_sigmoid_xent(MREC_AOT_ARGS);

/* ---- a7 DenseLayer backward glue ------------------------------------------------------------------------
 * ReluGrad + BiasAddGrad of one DenseLayer (wide_and_deep.py:72-133 bprop; deepfm.py:150-170;
 * deep_and_cross.py:161-200) in one pass, deterministic fp32 column sums:
 *   in : g[B,N] f16|f32, y[B,N] same dtype | numel 0 (no mask: plain BiasAddGrad)
 *   out: gz[B,N] = g * (y > 0) (may be g's own buffer; not written without a mask), gb[N] f32, workspace uint8
 * The workspace (mrec_relu_bwd_bias_workspace_bytes(N)) holds ticket counters: zero-fill it once; every launch
 * leaves them at zero. */
int mrec_relu_bwd_bias(MREC_AOT_ARGS);
size_t mrec_relu_bwd_bias_workspace_bytes(int64_t n_cols);

/* The one-unit output DenseLayer (the logit head: wide_and_deep.py:293-297 dense_layer_5; deepfm.py:215;
 * deep_and_cross.py:309) without skinny library GEMMs:
 *   mrec_dense_head_fwd  in : h[B,K] f16|f32, w[K] same dtype, bias[1] same dtype        out: out[B] f32
 *   mrec_dense_head_bwd  in : delta[B] f16|f32, h[B,K], w[K] (same dtype), relu_like[0|1] (numel 1: mask with h > 0)
 *                        out: gh[B,K] = (delta x w) * mask, gw[K] f32 = delta^T h, gb_head[1] f32 = sum(delta),
 *                             gb_prev[K] f32 | numel 0 = column sums of gh (previous layer's BiasAddGrad), workspace
 * Workspace: mrec_dense_head_workspace_bytes(K), zero-filled once (self-resetting ticket counters). */
int mrec_dense_head_fwd(MREC_AOT_ARGS);
int mrec_dense_head_bwd(MREC_AOT_ARGS);
size_t mrec_dense_head_workspace_bytes(int64_t k_dim);

/* ---- K7 FM second-order interaction -------------------------------------------------------------
 * Replaces Square/ReduceSum/Sub x6 of models/deepfm/src/deepfm.py:222-228 and their autodiff.
 *   fwd  in : vx[B,F,D] f32 (already multiplied by the mask)      out: fm[B]|[B,1] f32
 *   bwd  in : vx[B,F,D], gout[B]|[B,1], (addend[B,F,D] f32|f16)   out: dvx[B,F,D] = g * (sum_f vx - vx) + addend */
int mrec_fm_fwd(MREC_AOT_ARGS);
int mrec_fm_bwd(MREC_AOT_ARGS);

/* ---- K8 DCN cross stack --------------------------------------------------------------------------
 * Replaces CrossLayer.construct (models/deep_and_cross/src/deep_and_cross.py:139-149) applied L times
 * (:301-306) and its autodiff.  w[l], b[l] are the layer's cross_weight / cross_bias ([D',1] each in the
 * reference, stacked here as [L, D']); 1 <= L <= 8.
 *   fwd  in : x0[B,D'] w[L,D'] b[L,D']            out: y[B,D'] (= x_L), p[B,L] (saved dots x0.w_l),
 *                                                      workspace[mrec_cross_workspace_bytes(L, D')]
 *   bwd  in : x0 dy[B,D'] w b p[B,L]              out: dx[B,D'] dw[L,D'] db[L,D'], workspace          */
int mrec_cross_fwd(MREC_AOT_ARGS);
int mrec_cross_bwd(MREC_AOT_ARGS);
size_t mrec_cross_workspace_bytes(int64_t layers, int dp);

/* ---- K6 hash table behind MapParameter / HashEmbeddingLookup -------------------------------------
 * Replaces mindspore.experimental.MapParameter's GPUHashTable and the MapTensorGet/Put/Erase primitives
 * (mindspore_rec/ops/embedding.py:136-149,193; README.md:176-195).  The table maps key -> slot; values and
 * optimizer state are [C+1, D] arenas indexed by slot (row C = default row), so rows are fetched with
 * mrec_gather and updated with mrec_sparse_lazy_adam / mrec_sparse_ftrl on slot indices.
 * Framework-owned state, passed on every call and mutated in place:
 *   tkeys[C] i64 (-1 empty, -2 erased; C a power of two >= 8, 64-byte aligned), meta[C] i64
 *   ((sightings << 32) | last_step), state[8] i32 {resident, step, tombstones, overflow, occupied, erase-log
 *   entries, erase-log overflow, -},
 *   cfg[2] i32 {permit_filter_value, evict_filter_value}.
 * Probe entry points:   in : keys[N] i32|i64, tkeys, meta, state, cfg
 *   mrec_hash_find            MapTensorGet(insert_default_value=False)   out: slots[N] i32 (C = default row)
 *   mrec_hash_find_or_insert  MapTensorGet(True): a key becomes resident on its permit-th sighting
 *                             (one sighting per key per call)             out: slots[N], new_slots[N], new_count[1]
 *   mrec_hash_insert          MapTensorPut / import_data                  out: slots[N], new_slots[N], new_count[1]
 *   mrec_hash_erase           MapTensorErase                              out: slots[N] [, erase_log[L] i64]
 * erase_log (optional, also on mrec_hash_evict): every removed key is appended at state[5]++; state[6] is raised
 * when the log is full.  Together with mrec_hash_export(since) it gives the incremental export of
 * RELEASE.md:18 / README.md:213-214 (keys changed since a step + keys erased since then).              */
int mrec_hash_find(MREC_AOT_ARGS);
int mrec_hash_find_or_insert(MREC_AOT_ARGS);
int mrec_hash_insert(MREC_AOT_ARGS);
int mrec_hash_erase(MREC_AOT_ARGS);
/* Initialise the rows of newly resident keys in one arena:
 *   in : arena[C+1,D] f32, new_slots[N], new_count[1], tkeys[C], rng[2] i64 {seed, mode}, sigma[1] f32
 *   out: dummy[1].  mode 0 = copy the arena's default row; mode 1 = N(0, sigma^2), Philox keyed by the KEY. */
int mrec_hash_init_rows(MREC_AOT_ARGS);
/*   in : arena[C+1,D], slots[N], values[N,D]     out: dummy[1]     (put / import: arena[slots[i]] = values[i]) */
int mrec_hash_scatter_rows(MREC_AOT_ARGS);
/* Eviction sweep: erase every key not looked up for more than evict_filter_value calls.
 *   in : tkeys, meta, state, cfg                  out: dummy[1] [, erase_log[L] i64]                          */
int mrec_hash_evict(MREC_AOT_ARGS);
/* get_keys / export_data:  in : tkeys, meta, state, cfg [, since[1] i32: only keys looked up / put after that step]
 *                          out: keys_out[C] i64, slots_out[C] i32, count[1] i32                               */
int mrec_hash_export(MREC_AOT_ARGS);
/* Growth (upstream's GPUHashTable is a cuco dynamic_map that grows; mindspore_rec/ops/embedding.py:136-144 never sizes
 * it): rebuild the key -> slot map into a larger table, then move every arena's rows by the slot map.
 *   mrec_hash_rehash     in : tkeys_old[C], meta_old[C], state[8]
 *                        out: tkeys_new[C2] (pre-filled with -1), meta_new[C2] (zeros), slot_map[C] i32 (-1 = free slot)
 *   mrec_hash_move_rows  in : arena_old[C+1,D] f32, slot_map[C] i32         out: arena_new[C2+1,D] f32              */
int mrec_hash_rehash(MREC_AOT_ARGS);
int mrec_hash_move_rows(MREC_AOT_ARGS);

/* ---- input pipeline, host side (not aot, no GPU): TFRecord framing + tf.train.Example decoding -------------------
 * Replaces MindSpore's C++ dataset engine behind ds.TFRecordDataset for the reference's Criteo files
 * (models/wide_deep/src/datasets.py:272-326: columns feat_ids int32, feat_vals float32, label float32, one record =
 * line_per_sample = 1000 samples; writer datasets/criteo_1tb/process_data.py:203-283).  The Python loader
 * (mindrec_b200/data.py) maps the files, decodes into pinned host buffers and copies to the device on a copy stream.
 *   mrec_tfrecord_index  scan a file image: offsets / lengths of the record payloads; returns the record count or
 *                        -(byte position + 1) of the first framing / CRC error
 *   mrec_tfrecord_parse  decode feature `name` of one serialized Example; kind 0: Int64List -> int32, 1: FloatList -> float
 *   mrec_crc32c(_masked) CRC-32C of a buffer (the masked form is what TFRecord stores); mrec_varint_pack: writer side */
int64_t mrec_tfrecord_index(const uint8_t *buf, int64_t n, int64_t *offsets, int64_t *lengths, int64_t max_records,
                            int check_crc);
int mrec_tfrecord_parse(const uint8_t *rec, int64_t len, const char *name, int kind, void *out, int64_t cap,
                        int64_t *count);
uint32_t mrec_crc32c(const void *data, size_t n);
uint32_t mrec_crc32c_masked(const void *data, size_t n);
int64_t mrec_varint_pack(const int32_t *v, int64_t n, uint8_t *out, int64_t cap);

/* ---- host runtime (not aot): what the Python host code needs to drive the aot kernels without PyTorch ---------------
 * (mindrec_b200/runtime.py binds these with ctypes; under MindSpore the framework owns memory and streams instead).
 * Memory and copies: kind 1 = host -> device, 2 = device -> host, 3 = device -> device, asynchronous on `stream`.
 * Graph: mrec_rt_graph_begin(stream) ... enqueue aot calls on that stream ... mrec_rt_graph_end(stream) -> executable. */
int mrec_rt_device_count(void);
int mrec_rt_set_device(int index);
int mrec_rt_mem_info(size_t *free_bytes, size_t *total_bytes);
void *mrec_rt_malloc(size_t bytes);
int mrec_rt_free(void *p);
void *mrec_rt_malloc_host(size_t bytes);
int mrec_rt_free_host(void *p);
int mrec_rt_memcpy(void *dst, const void *src, size_t bytes, int kind, void *stream);
int mrec_rt_memset(void *p, int byte, size_t bytes, void *stream);
int mrec_rt_fill32(void *p, uint32_t pattern, int64_t n, void *stream);
void *mrec_rt_stream_create(void);
int mrec_rt_stream_destroy(void *stream);
int mrec_rt_stream_sync(void *stream);
int mrec_rt_device_sync(void);
void *mrec_rt_event_create(int timing);
int mrec_rt_event_record(void *event, void *stream);
int mrec_rt_event_sync(void *event);
int mrec_rt_stream_wait_event(void *stream, void *event);
float mrec_rt_event_elapsed_ms(void *start, void *stop);
int mrec_rt_event_destroy(void *event);
int mrec_rt_graph_begin(void *stream);
void *mrec_rt_graph_end(void *stream);
int mrec_rt_graph_launch(void *graph_exec, void *stream);
int mrec_rt_graph_destroy(void *graph_exec);
/* DenseLayer GEMM of the torch-free host path (wide_and_deep.py:72-133: MatMul + BiasAdd + ReLU; fp16 storage with fp32
 * accumulation under use_mixed_precision).  cuBLASLt, bound at run time with dlopen; asynchronous on `stream`, capturable.
 * Row-major C[M,N] = act(alpha * op(A) op(B) + beta * C + bias[N]); A stored [M,K] (trans_a = 0) or [K,M]; B stored [K,N]
 * (trans_b = 0) or [N,K]; ab_half / c_half: 1 = fp16, 0 = fp32 storage; epilogue 0 none | 1 bias | 2 relu(bias). */
int mrec_rt_gemm(const void *a, const void *b, void *c, const void *bias, int64_t m, int64_t n, int64_t k, int trans_a,
                 int trans_b, int ab_half, int c_half, float alpha, float beta, int epilogue, void *stream);
/* Cast(weight, float16) of the mixed-precision DenseLayers for the whole flat parameter buffer (aot signature).
 *   in : src[n] f32      out: dst[n] f16 */
int mrec_cast_f32_f16(MREC_AOT_ARGS);

#ifdef __cplusplus
}
#endif
#endif /* MINDREC_B200_H_ */
