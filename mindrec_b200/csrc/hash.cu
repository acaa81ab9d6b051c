// K6 — GPU open-addressing hash table backing MapParameter / HashEmbeddingLookup.
//
// Replaces upstream's GPUHashTable (cuCollections dynamic_map + value arena) reached from
//   mindspore_rec/ops/embedding.py:136-149,193  (MapParameter + MapTensorGet(insert_default_value=True))
//   README.md:176-195                            (user-level put / get / erase)
//
// The table only maps key -> row index ("slot").  Values, and the optimizer state that lives beside them,
// are ordinary [C+1, D] arenas addressed by that index, so lookup = find-or-insert + the K1 gather, and the
// backward pass = K2 unique over slot indices (log2 C bits) + the fused K3/K4/K5 row updates: one probe per
// key per step however many sibling arenas (w, m, v / accum, linear) the optimizer keeps.  Row C of every
// arena is the "default row" returned for keys that are not (yet) resident.
//
// Layout (all framework-owned, passed on every call; the library keeps no hidden state):
//   keys [C]  int64   EMPTY = -1, ERASED = -2 (the two key values the reference reserves, embedding.py:55-57)
//   meta [C]  int64   (sightings << 32) | last_step   — admission counter and eviction timestamp; 0 when free
//   state[8]  int32   [0] resident keys, [1] step, [2] tombstones, [3] overflow flag, [4] occupied slots
// C is a power of two >= 8.  Probing is linear over groups of 8 slots: a tile of 8 lanes loads one aligned
// 64-byte line of keys per probe and votes (match / empty / reusable) with a ballot; inserts claim the
// first reusable slot of the probe sequence with a 64-bit CAS, so every inserter of a key walks the same
// sequence and duplicates inside one call resolve to one slot.  Admission ("permit") counts one sighting per
// key per call: the (sightings, last_step) word is advanced by a CAS loop whose single winner per (key, call)
// also detects the transition into residency and reports the slot in new_slots, so that
// mrec_hash_init_rows can initialise the rows of every arena exactly once.
#include "common.cuh"

namespace mrec {

constexpr long long kEmpty = -1;
constexpr long long kErased = -2;
constexpr int kHashThreads = 256;
// A probe sequence is at most this many 8-slot groups long, for finds, inserts and the rebuild alike: a key therefore
// always sits within kMaxProbeGroups groups of its home group, and a miss on a FULL table costs 64 KB of key reads
// instead of the whole table (a table driven past its capacity raises the overflow flag; it must not also stall
// every lookup for seconds).  At the load factors the table is run at (growth at 0.6) chains are a few groups long.
constexpr int64_t kMaxProbeGroups = 1024;
enum { ST_SIZE = 0, ST_STEP = 1, ST_TOMB = 2, ST_OVERFLOW = 3, ST_OCC = 4, ST_LOGN = 5, ST_LOGOVF = 6, ST_LEN = 8 };

// Erase log (incremental export): every key removed by erase / evict is appended to a
// caller-owned int64 buffer; state[ST_LOGN] counts the appends, state[ST_LOGOVF] is raised when one did not fit
// (the next incremental export must then be a full one).
__device__ __forceinline__ void log_erased(long long key, long long* log, int64_t log_cap, int32_t* state) {
  if (!log) return;
  const int i = atomicAdd(&state[ST_LOGN], 1);
  if (i < log_cap) log[i] = key;
  else state[ST_LOGOVF] = 1;
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {  // murmur3 finaliser
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

__global__ void hash_begin_kernel(int32_t* state, int32_t* new_count, int advance) {
  if (threadIdx.x == 0) {
    if (advance) state[ST_STEP] += 1;
    if (new_count) new_count[0] = 0;
  }
}

// Advance (sightings, last_step) at most once per call; returns the count after this call's increment.
// *became_resident is true for exactly one caller per key: the one whose CAS moved the count across `permit`.
__device__ __forceinline__ uint32_t touch_meta(unsigned long long* meta, int step, uint32_t permit,
                                               bool force_admit, bool* became_resident) {
  unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(meta);
  while (true) {
    const uint32_t cnt = (uint32_t)(old >> 32);
    const int last = (int)(uint32_t)old;
    uint32_t cnt_new = (last == step) ? cnt : (cnt == 0xffffffffu ? cnt : cnt + 1u);
    if (force_admit && cnt_new < permit) cnt_new = permit;
    const unsigned long long want = ((unsigned long long)cnt_new << 32) | (uint32_t)step;
    if (want == old) {
      *became_resident = false;
      return cnt_new;
    }
    const unsigned long long prev = atomicCAS(meta, old, want);
    if (prev == old) {
      *became_resident = (cnt < permit) && (cnt_new >= permit);
      return cnt_new;
    }
    old = prev;
  }
}

// MODE 0: find only (missing / not admitted -> default row C)     MODE 1: find or insert, admission by permit
// MODE 2: insert, admitted outright (put / import)                MODE 3: erase
template <typename KeyT, int MODE>
__global__ void __launch_bounds__(kHashThreads)
hash_probe_kernel(const KeyT* __restrict__ keys_in, int64_t n, long long* __restrict__ keys,
                  unsigned long long* __restrict__ meta, int32_t* __restrict__ state, int64_t capacity,
                  const int32_t* __restrict__ cfg, int32_t* __restrict__ slots, int32_t* __restrict__ new_slots,
                  int32_t* __restrict__ new_count, long long* __restrict__ erase_log, int64_t log_cap) {
  const int permit = cfg[0];  // MapParameter(permit_filter_value): device scalar like every other "attr"
  const int lane = threadIdx.x & 31;
  const int lt = lane & 7;      // lane in tile
  const int tbase = lane & ~7;  // first lane of my tile == bit offset of my tile inside a warp ballot
  const int64_t n_groups = capacity >> 3;
  const int64_t probe_cap = n_groups < kMaxProbeGroups ? n_groups : kMaxProbeGroups;
  const int step = state[ST_STEP];
  const int64_t base = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) & ~(int64_t)7;  // tile's first item
  const int64_t mine = base + lt;
  const long long my_key = (mine < n) ? (long long)keys_in[mine] : kEmpty;
  int32_t my_slot = (int32_t)capacity;

  for (int m = 0; m < 8; ++m) {
    const long long key = __shfl_sync(0xffffffffu, my_key, tbase + m);
    // reserved key values (and the padding of a ragged last tile) map to the default row
    bool done = (base + m >= n) || key == kEmpty || key == kErased;
    const int64_t g0 = (int64_t)(mix64((uint64_t)key) & (uint64_t)(n_groups - 1));
    int64_t g = g0, first_free = -1, probes = 0;
    long long first_free_val = kEmpty;
    int32_t result = (int32_t)capacity;
    bool claimed = false;
    while (__any_sync(0xffffffffu, !done)) {
      long long k = kEmpty;
      if (!done) k = *reinterpret_cast<volatile long long*>(&keys[g * 8 + lt]);
      const uint32_t bm = (__ballot_sync(0xffffffffu, !done && k == key) >> tbase) & 0xffu;
      const uint32_t be = (__ballot_sync(0xffffffffu, !done && k == kEmpty) >> tbase) & 0xffu;
      const uint32_t bf = (__ballot_sync(0xffffffffu, !done && (k == kEmpty || k == kErased)) >> tbase) & 0xffu;
      bool want_cas = false;
      if (!done) {
        if (bm) {
          result = (int32_t)(g * 8 + (__ffs(bm) - 1));
          done = true;
        } else {
          if (first_free < 0 && bf) {
            const int fl = __ffs(bf) - 1;
            first_free = g * 8 + fl;
            first_free_val = ((be >> fl) & 1u) ? kEmpty : kErased;
          }
          ++probes;
          if (be || probes >= probe_cap) {  // end of the probe chain: the key is not in the table
            if ((MODE == 1 || MODE == 2) && first_free >= 0) {
              want_cas = true;
            } else {
              if ((MODE == 1 || MODE == 2) && lt == 0) atomicExch(&state[ST_OVERFLOW], 1);
              done = true;
            }
          } else {
            g = (g + 1) & (n_groups - 1);
          }
        }
      }
      // claim — at a warp-uniform point so that every lane takes part in the shuffle
      long long cas_old = 0;
      if (want_cas && lt == 0)
        cas_old = (long long)atomicCAS(reinterpret_cast<unsigned long long*>(&keys[first_free]),
                                       (unsigned long long)first_free_val, (unsigned long long)key);
      cas_old = __shfl_sync(0xffffffffu, cas_old, tbase);
      if (want_cas) {
        if (cas_old == first_free_val) {
          result = (int32_t)first_free;
          claimed = true;
          done = true;
        } else if (cas_old == key) {  // a duplicate of this key won the race for the same slot
          result = (int32_t)first_free;
          done = true;
        } else {  // another key took the slot: walk the sequence again
          g = g0;
          first_free = -1;
          probes = 0;
        }
      }
    }
    // bookkeeping by the tile leader; the outcome goes to the lane that owns the key
    int32_t out_slot = (int32_t)capacity;
    if (lt == 0 && result < (int32_t)capacity) {
      if (MODE == 3) {
        const long long prev = (long long)atomicCAS(reinterpret_cast<unsigned long long*>(&keys[result]),
                                                    (unsigned long long)key, (unsigned long long)kErased);
        if (prev == key) {  // one eraser per key
          const unsigned long long mo = atomicExch(&meta[result], 0ull);
          atomicAdd(&state[ST_TOMB], 1);
          atomicSub(&state[ST_OCC], 1);
          if ((uint32_t)(mo >> 32) >= (uint32_t)permit) atomicSub(&state[ST_SIZE], 1);
          log_erased(key, erase_log, log_cap, state);
        }
        out_slot = result;
      } else if (MODE == 0) {
        const uint32_t cnt = (uint32_t)(*reinterpret_cast<volatile unsigned long long*>(&meta[result]) >> 32);
        out_slot = (cnt >= (uint32_t)permit) ? result : (int32_t)capacity;
      } else {
        if (claimed) {
          atomicAdd(&state[ST_OCC], 1);
          if (first_free_val == kErased) atomicSub(&state[ST_TOMB], 1);
        }
        bool became = false;
        const uint32_t cnt = touch_meta(&meta[result], step, (uint32_t)permit, MODE == 2, &became);
        if (became) {
          atomicAdd(&state[ST_SIZE], 1);
          if (new_slots) new_slots[atomicAdd(new_count, 1)] = result;
        }
        out_slot = (cnt >= (uint32_t)permit) ? result : (int32_t)capacity;
      }
    }
    out_slot = __shfl_sync(0xffffffffu, out_slot, tbase);
    if (lt == m) my_slot = out_slot;
  }
  if (mine < n && slots) slots[mine] = my_slot;
}

// ---- Philox4x32-10 (counter = (slot-independent key hash, column), key = seed) for 'normal' init ----
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
  c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
}
__device__ __forceinline__ float4 philox_normal4(uint64_t ctr_hi, uint32_t ctr_lo, uint64_t seed) {
  uint32_t c[4] = {ctr_lo, 0u, (uint32_t)ctr_hi, (uint32_t)(ctr_hi >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const float u0 = (c[0] + 0.5f) * 2.3283064365386963e-10f, u1 = (c[1] + 0.5f) * 2.3283064365386963e-10f;
  const float u2 = (c[2] + 0.5f) * 2.3283064365386963e-10f, u3 = (c[3] + 0.5f) * 2.3283064365386963e-10f;
  const float r0 = sqrtf(-2.f * __logf(u0)), r1 = sqrtf(-2.f * __logf(u2));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u1, &s0, &c0);
  __sincosf(6.283185307179586f * u3, &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

// arena[new_slots[i], :] = arena[C, :] (mode 0: the arena's default row) or N(0, sigma^2) keyed by the KEY
// (mode 1: the same key always draws the same row, whatever slot it lands in).
__global__ void __launch_bounds__(256)
hash_init_rows_kernel(float* __restrict__ arena, int64_t capacity, int dim, const int32_t* __restrict__ new_slots,
                      const int32_t* __restrict__ new_count, const long long* __restrict__ keys,
                      const long long* __restrict__ rng /* {seed, mode} */, const float* __restrict__ sigma_p) {
  const uint64_t seed = (uint64_t)rng[0];
  const int mode = (int)rng[1];
  const float sigma = sigma_p[0];
  const int n_new = new_count[0];
  const int64_t total = (int64_t)n_new * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / dim;
    const int d = (int)(e - i * dim);
    const int32_t s = new_slots[i];
    float v;
    if (mode == 0) {
      v = arena[capacity * dim + d];
    } else {
      const float4 z = philox_normal4((uint64_t)keys[s], (uint32_t)(d >> 2), seed);
      const int q = d & 3;
      v = sigma * (q == 0 ? z.x : q == 1 ? z.y : q == 2 ? z.z : z.w);
    }
    arena[(int64_t)s * dim + d] = v;
  }
}

// put / import: arena[slots[i], :] = values[i, :]  (slot C = not admitted: skipped)
__global__ void __launch_bounds__(256)
hash_scatter_rows_kernel(float* __restrict__ arena, int64_t capacity, int dim, const int32_t* __restrict__ slots,
                         const float* __restrict__ values, int64_t n) {
  const int64_t total = n * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / dim;
    const int32_t s = slots[i];
    if (s >= 0 && s < capacity) arena[(int64_t)s * dim + (e - i * dim)] = values[e];
  }
}

// eviction: every resident or candidate key not looked up for more than `evict_after` steps is erased
__global__ void __launch_bounds__(256)
hash_evict_kernel(long long* __restrict__ keys, unsigned long long* __restrict__ meta, int32_t* __restrict__ state,
                  int64_t capacity, const int32_t* __restrict__ cfg, long long* __restrict__ erase_log, int64_t log_cap) {
  const int permit = cfg[0], evict_after = cfg[1];
  const int step = state[ST_STEP];
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < capacity;
       s += (int64_t)gridDim.x * blockDim.x) {
    const long long k = keys[s];
    if (k == kEmpty || k == kErased) continue;
    const unsigned long long mo = meta[s];
    const int last = (int)(uint32_t)mo;
    if ((int64_t)step - (int64_t)last > (int64_t)evict_after) {
      keys[s] = kErased;
      meta[s] = 0ull;
      atomicAdd(&state[ST_TOMB], 1);
      atomicSub(&state[ST_OCC], 1);
      if ((uint32_t)(mo >> 32) >= (uint32_t)permit) atomicSub(&state[ST_SIZE], 1);
      log_erased(k, erase_log, log_cap, state);
    }
  }
}

// export: resident (key, slot) pairs, compacted (order unspecified); values follow with mrec_gather(arena, slots)
__global__ void __launch_bounds__(256)
hash_export_kernel(const long long* __restrict__ keys, const unsigned long long* __restrict__ meta,
                   int64_t capacity, const int32_t* __restrict__ cfg, const int32_t* __restrict__ since,
                   long long* __restrict__ keys_out, int32_t* __restrict__ slots_out, int32_t* __restrict__ count) {
  const int permit = cfg[0];
  const bool incremental = since != nullptr;       // only keys looked up / put after step since[0]
  const int since_step = incremental ? since[0] : 0;
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < capacity;
       s += (int64_t)gridDim.x * blockDim.x) {
    const long long k = keys[s];
    if (k == kEmpty || k == kErased) continue;
    const unsigned long long mo = meta[s];
    if ((uint32_t)(mo >> 32) < (uint32_t)permit) continue;
    if (incremental && (int)(uint32_t)mo <= since_step) continue;
    const int idx = atomicAdd(count, 1);
    keys_out[idx] = k;
    slots_out[idx] = (int32_t)s;
  }
}


// ---- growth: rebuild the key -> slot map into a larger table (upstream's GPUHashTable grows) ----------
// One thread per OLD slot: a live key (resident or still a candidate) is inserted into the new table with the probe
// sequence of hash_probe_kernel (groups of 8 slots, linear over groups; a group is left only when all 8 of its slots
// were seen taken, and slots are never freed during the rebuild, so every later lookup that walks the same sequence
// finds the key before it meets an EMPTY slot).  Keys of the old table are distinct: a plain CAS per slot suffices.
// The admission / eviction word travels with the key; slot_map[old] = new slot (or -1) drives mrec_hash_move_rows.
__global__ void __launch_bounds__(256)
hash_rehash_kernel(const long long* __restrict__ keys_old, const unsigned long long* __restrict__ meta_old,
                   int64_t cap_old, long long* __restrict__ keys_new, unsigned long long* __restrict__ meta_new,
                   int64_t cap_new, int32_t* __restrict__ slot_map, int32_t* __restrict__ state) {
  const int64_t n_groups = cap_new >> 3;
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < cap_old; s += (int64_t)gridDim.x * blockDim.x) {
    const long long k = keys_old[s];
    if (k == kEmpty || k == kErased) {
      slot_map[s] = -1;
      continue;
    }
    int64_t g = (int64_t)(mix64((uint64_t)k) & (uint64_t)(n_groups - 1));
    int32_t placed = -1;
    const int64_t probe_cap = n_groups < kMaxProbeGroups ? n_groups : kMaxProbeGroups;
    for (int64_t probes = 0; probes < probe_cap && placed < 0; ++probes) {
      for (int j = 0; j < 8; ++j) {
        const int64_t slot = g * 8 + j;
        if (*reinterpret_cast<volatile long long*>(&keys_new[slot]) != kEmpty) continue;
        const long long prev = (long long)atomicCAS(reinterpret_cast<unsigned long long*>(&keys_new[slot]),
                                                    (unsigned long long)kEmpty, (unsigned long long)k);
        if (prev == kEmpty) {
          placed = (int32_t)slot;
          break;
        }
      }
      g = (g + 1) & (n_groups - 1);
    }
    slot_map[s] = placed;
    if (placed >= 0) meta_new[placed] = meta_old[s];
    else atomicExch(&state[ST_OVERFLOW], 1);          // cannot happen when cap_new >= cap_old
  }
}

__global__ void hash_rehash_finish_kernel(int32_t* state) {
  if (threadIdx.x == 0 && blockIdx.x == 0) state[ST_TOMB] = 0;   // tombstones do not travel
}

// arena_new[slot_map[s], :] = arena_old[s, :] for every live slot, and the default row C_old -> C_new
__global__ void __launch_bounds__(256)
hash_move_rows_kernel(const float* __restrict__ arena_old, int64_t cap_old, const int32_t* __restrict__ slot_map,
                      float* __restrict__ arena_new, int64_t cap_new, int dim) {
  const int64_t total = (cap_old + 1) * dim;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = e / dim;
    const int64_t dst = (s == cap_old) ? cap_new : (int64_t)slot_map[s];
    if (dst >= 0) arena_new[dst * dim + (e - s * dim)] = arena_old[e];
  }
}

static int table_args(const Aot& a, int ki, int mi, int si, int64_t* capacity, const char* who) {
  MREC_REQUIRE(a.is_i64(ki) && a.is_i64(mi) && a.is_i32(si), ERR_DTYPE,
               "%s: table keys/meta must be int64, state int32", who);
  *capacity = a.numel(ki);
  MREC_REQUIRE(*capacity >= 8 && ((*capacity) & (*capacity - 1)) == 0, ERR_SHAPE,
               "%s: capacity must be a power of two >= 8", who);
  MREC_REQUIRE(a.numel(mi) == *capacity && a.numel(si) >= ST_LEN, ERR_SHAPE, "%s: meta[C] / state[8] expected", who);
  MREC_REQUIRE(*capacity < ((int64_t)1 << 31) - 1, ERR_SHAPE, "%s: capacity must be < 2^31", who);
  MREC_REQUIRE(a.aligned(ki, 64), ERR_ALIGN, "%s: table keys must be 64-byte aligned", who);
  return OK;
}

template <int MODE>
static int probe_entry(const Aot& a, const char* who, bool has_new) {
  // inputs : keys_in[N] i32|i64, table_keys[C] i64, meta[C] i64, state[8] i32, permit[1] i32
  // outputs: slots[N] i32, (new_slots[N] i32, new_count[1] i32)
  const int expect = has_new ? 8 : 6;
  const bool with_log = (MODE == 3) && a.nparam == expect + 1;     // erase: optional trailing erase_log[L] i64
  if (a.nparam != expect && !with_log)
    return fail(ERR_NPARAM, "%s: expected %d params, got %d", who, expect, a.nparam);
  for (int i = 0; i < a.nparam; ++i)
    if (!a.params[i] && a.numel(i) > 0) return fail(ERR_NULL, "%s: param %d is null", who, i);
  MREC_REQUIRE(a.is_i32(0) || a.is_i64(0), ERR_DTYPE, "%s: keys must be int32|int64", who);
  int64_t capacity = 0;
  int rc = table_args(a, 1, 2, 3, &capacity, who);
  if (rc) return rc;
  MREC_REQUIRE(a.is_i32(4) && a.numel(4) >= 2, ERR_DTYPE, "%s: params must be int32[2] = {permit, evict_after}", who);
  const int64_t n = a.numel(0);
  MREC_REQUIRE(a.is_i32(5) && a.numel(5) >= n, ERR_SHAPE, "%s: slots must be int32[N]", who);
  int32_t *new_slots = nullptr, *new_count = nullptr;
  if (has_new) {
    MREC_REQUIRE(a.is_i32(6) && a.is_i32(7) && a.numel(6) >= n && a.numel(7) >= 1, ERR_SHAPE,
                 "%s: new_slots[N] / new_count[1] int32 expected", who);
    new_slots = a.ptr<int32_t>(6);
    new_count = a.ptr<int32_t>(7);
  }
  long long* erase_log = nullptr;
  int64_t log_cap = 0;
  if (with_log) {
    MREC_REQUIRE(a.is_i64(6), ERR_DTYPE, "%s: erase_log must be int64", who);
    erase_log = a.ptr<long long>(6);
    log_cap = a.numel(6);
  }
  MREC_LAUNCH(hash_begin_kernel, 1, 32, 0, a.stream, a.ptr<int32_t>(3), new_count, (MODE == 1 || MODE == 2) ? 1 : 0);
  if (n == 0) return check_launch(who);
  const int grid = (int)cdiv(n, kHashThreads);
  if (a.is_i32(0)) {
    MREC_LAUNCH((hash_probe_kernel<int32_t, MODE>), grid, kHashThreads, 0, a.stream, a.ptr<int32_t>(0), n,
                a.ptr<long long>(1), a.ptr<unsigned long long>(2), a.ptr<int32_t>(3), capacity,
                a.ptr<int32_t>(4), a.ptr<int32_t>(5), new_slots, new_count, erase_log, log_cap);
  } else {
    MREC_LAUNCH((hash_probe_kernel<int64_t, MODE>), grid, kHashThreads, 0, a.stream, a.ptr<int64_t>(0), n,
                a.ptr<long long>(1), a.ptr<unsigned long long>(2), a.ptr<int32_t>(3), capacity,
                a.ptr<int32_t>(4), a.ptr<int32_t>(5), new_slots, new_count, erase_log, log_cap);
  }
  return check_launch(who);
}

}  // namespace mrec

using namespace mrec;

#define MREC_AOT_SIG \
  int nparam, void **params, int *ndims, int64_t **shapes, const char **dtypes, void *stream, void * /*extra*/
#define MREC_AOT_PACK Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream}

// Common params of the four probe entry points:
//   in : keys[N] i32|i64, tkeys[C] i64, meta[C] i64, state[8] i32, cfg[2] i32 = {permit, evict_after}
// MapTensorGet(insert_default_value=False):                      out: slots[N] i32
MREC_API int mrec_hash_find(MREC_AOT_SIG) {
  MREC_AOT_PACK;
  return probe_entry<0>(a, "mrec_hash_find", false);
}
// MapTensorGet(insert_default_value=True):                       out: slots[N], new_slots[N], new_count[1]
MREC_API int mrec_hash_find_or_insert(MREC_AOT_SIG) {
  MREC_AOT_PACK;
  return probe_entry<1>(a, "mrec_hash_find_or_insert", true);
}
// MapTensorPut / import_data (keys admitted outright):           out: slots[N], new_slots[N], new_count[1]
MREC_API int mrec_hash_insert(MREC_AOT_SIG) {
  MREC_AOT_PACK;
  return probe_entry<2>(a, "mrec_hash_insert", true);
}
// MapTensorErase:                                                out: slots[N] (freed slot, or C if absent)
MREC_API int mrec_hash_erase(MREC_AOT_SIG) {
  MREC_AOT_PACK;
  return probe_entry<3>(a, "mrec_hash_erase", false);
}

// in : arena[C+1,D] f32, new_slots[N] i32, new_count[1] i32, tkeys[C] i64, rng[2] i64 = {seed, mode}, sigma[1] f32
// out: dummy[1].  mode 0: copy the arena's default row C; mode 1: N(0, sigma^2) from Philox(seed; key, column).
MREC_API int mrec_hash_init_rows(MREC_AOT_SIG) {
  MREC_AOT_PACK;
  MREC_CHECK_NPARAM(a, 7);
  MREC_REQUIRE(a.is_f32(0) && a.ndims[0] == 2, ERR_SHAPE, "mrec_hash_init_rows: arena must be f32 [C+1, D]");
  MREC_REQUIRE(a.is_i32(1) && a.is_i32(2) && a.is_i64(3) && a.is_i64(4) && a.is_f32(5), ERR_DTYPE,
               "mrec_hash_init_rows: new_slots/new_count int32, tkeys/rng int64, sigma float32");
  const int64_t capacity = a.numel(3);
  MREC_REQUIRE(a.dim(0, 0) == capacity + 1, ERR_SHAPE, "mrec_hash_init_rows: arena must have C+1 rows");
  MREC_REQUIRE(a.numel(4) >= 2 && a.numel(5) >= 1, ERR_SHAPE, "mrec_hash_init_rows: rng[2], sigma[1] expected");
  const int dim = (int)a.dim(0, 1);
  const int64_t n = a.numel(1);
  if (n == 0) return OK;
  MREC_LAUNCH(hash_init_rows_kernel, grid_for(cdiv(n * dim, 256), 8), 256, 0, a.stream, a.ptr<float>(0), capacity,
              dim, a.ptr<int32_t>(1), a.ptr<int32_t>(2), a.ptr<long long>(3), a.ptr<long long>(4), a.ptr<float>(5));
  return check_launch("hash_init_rows");
}

// put / import: in arena[C+1,D] f32, slots[N] i32, values[N,D] f32     out: dummy[1]
MREC_API int mrec_hash_scatter_rows(MREC_AOT_SIG) {
  MREC_AOT_PACK;
  MREC_CHECK_NPARAM(a, 4);
  MREC_REQUIRE(a.is_f32(0) && a.ndims[0] == 2 && a.is_i32(1) && a.is_f32(2), ERR_DTYPE,
               "mrec_hash_scatter_rows: arena f32 [C+1,D], slots int32, values f32");
  const int64_t capacity = a.dim(0, 0) - 1;
  const int dim = (int)a.dim(0, 1);
  const int64_t n = a.numel(1);
  MREC_REQUIRE(a.numel(2) == n * dim, ERR_SHAPE, "mrec_hash_scatter_rows: values must be [N, D]");
  if (n == 0) return OK;
  MREC_LAUNCH(hash_scatter_rows_kernel, grid_for(cdiv(n * dim, 256), 8), 256, 0, a.stream, a.ptr<float>(0),
              capacity, dim, a.ptr<int32_t>(1), a.ptr<float>(2), n);
  return check_launch("hash_scatter_rows");
}

// eviction sweep: in tkeys[C] meta[C] state[8] cfg[2]     out: dummy[1] [, erase_log[L] i64]
MREC_API int mrec_hash_evict(MREC_AOT_SIG) {
  MREC_AOT_PACK;
  if (a.nparam != 5 && a.nparam != 6) return fail(ERR_NPARAM, "mrec_hash_evict: expected 5 or 6 params, got %d", a.nparam);
  long long* erase_log = nullptr;
  int64_t log_cap = 0;
  if (a.nparam == 6) {
    MREC_REQUIRE(a.is_i64(5), ERR_DTYPE, "mrec_hash_evict: erase_log must be int64");
    erase_log = a.ptr<long long>(5);
    log_cap = a.numel(5);
  }
  int64_t capacity = 0;
  int rc = table_args(a, 0, 1, 2, &capacity, "mrec_hash_evict");
  if (rc) return rc;
  MREC_REQUIRE(a.is_i32(3) && a.numel(3) >= 2, ERR_DTYPE, "mrec_hash_evict: cfg must be int32[2]");
  MREC_LAUNCH(hash_evict_kernel, grid_for(cdiv(capacity, 256), 8), 256, 0, a.stream, a.ptr<long long>(0),
              a.ptr<unsigned long long>(1), a.ptr<int32_t>(2), capacity, a.ptr<int32_t>(3), erase_log, log_cap);
  return check_launch("hash_evict");
}

// export_data / get_keys: in tkeys[C] meta[C] state[8] cfg[2] [, since[1] i32]
//                          out: keys_out[C] i64, slots_out[C] i32, count[1] i32
// (values follow with mrec_gather(arena, slots_out); order of the pairs is unspecified).  With `since` only the
// resident keys looked up or put after step since[0] are exported (incremental export; the erased keys of the
// same interval are in the erase log of mrec_hash_erase / mrec_hash_evict).
MREC_API int mrec_hash_export(MREC_AOT_SIG) {
  MREC_AOT_PACK;
  if (a.nparam != 7 && a.nparam != 8) return fail(ERR_NPARAM, "mrec_hash_export: expected 7 or 8 params, got %d", a.nparam);
  const int o = a.nparam - 3;                       // first output
  int64_t capacity = 0;
  int rc = table_args(a, 0, 1, 2, &capacity, "mrec_hash_export");
  if (rc) return rc;
  MREC_REQUIRE(a.is_i32(3) && a.is_i64(o) && a.is_i32(o + 1) && a.is_i32(o + 2), ERR_DTYPE,
               "mrec_hash_export: cfg int32, keys_out int64, slots_out/count int32");
  const int32_t* since = nullptr;
  if (a.nparam == 8) {
    MREC_REQUIRE(a.is_i32(4) && a.numel(4) >= 1, ERR_DTYPE, "mrec_hash_export: since must be int32[1]");
    since = a.ptr<int32_t>(4);
  }
  MREC_REQUIRE(a.numel(o) >= capacity && a.numel(o + 1) >= capacity && a.numel(o + 2) >= 1, ERR_SHAPE,
               "mrec_hash_export: outputs must be padded to C");
  cudaMemsetAsync(a.params[o + 2], 0, sizeof(int32_t), a.stream);
  MREC_LAUNCH(hash_export_kernel, grid_for(cdiv(capacity, 256), 8), 256, 0, a.stream, a.ptr<long long>(0),
              a.ptr<unsigned long long>(1), capacity, a.ptr<int32_t>(3), since, a.ptr<long long>(o), a.ptr<int32_t>(o + 1),
              a.ptr<int32_t>(o + 2));
  return check_launch("hash_export");
}

// Growth (rehash into a larger table).  in : tkeys_old[C] meta_old[C] state[8]
//                                       out: tkeys_new[C2] (pre-filled with -1), meta_new[C2] (zeros), slot_map[C] i32
// C2 a power of two >= C.  Resident keys, candidates and their admission / eviction words move; tombstones vanish.
MREC_API int mrec_hash_rehash(MREC_AOT_SIG) {
  MREC_AOT_PACK;
  MREC_CHECK_NPARAM(a, 6);
  int64_t cap_old = 0, cap_new = 0;
  int rc = table_args(a, 0, 1, 2, &cap_old, "mrec_hash_rehash");
  if (rc) return rc;
  MREC_REQUIRE(a.is_i64(3) && a.is_i64(4) && a.is_i32(5), ERR_DTYPE, "mrec_hash_rehash: new keys/meta int64, slot_map int32");
  cap_new = a.numel(3);
  MREC_REQUIRE(cap_new >= cap_old && (cap_new & (cap_new - 1)) == 0 && a.numel(4) == cap_new && a.numel(5) >= cap_old,
               ERR_SHAPE, "mrec_hash_rehash: new capacity must be a power of two >= the old one; slot_map[C]");
  MREC_REQUIRE(cap_new < ((int64_t)1 << 31) - 1, ERR_SHAPE, "mrec_hash_rehash: capacity must be < 2^31");
  MREC_REQUIRE(a.aligned(3, 64), ERR_ALIGN, "mrec_hash_rehash: table keys must be 64-byte aligned");
  MREC_LAUNCH(hash_rehash_kernel, grid_for(cdiv(cap_old, 256), 8), 256, 0, a.stream, a.ptr<long long>(0),
              a.ptr<unsigned long long>(1), cap_old, a.ptr<long long>(3), a.ptr<unsigned long long>(4), cap_new,
              a.ptr<int32_t>(5), a.ptr<int32_t>(2));
  MREC_LAUNCH(hash_rehash_finish_kernel, 1, 32, 0, a.stream, a.ptr<int32_t>(2));
  return check_launch("hash_rehash");
}

// in : arena_old[C+1,D] f32, slot_map[C] i32      out: arena_new[C2+1,D] f32 (rows of live slots + the default row)
MREC_API int mrec_hash_move_rows(MREC_AOT_SIG) {
  MREC_AOT_PACK;
  MREC_CHECK_NPARAM(a, 3);
  MREC_REQUIRE(a.is_f32(0) && a.ndims[0] == 2 && a.is_i32(1) && a.is_f32(2) && a.ndims[2] == 2, ERR_DTYPE,
               "mrec_hash_move_rows: arenas f32 [C+1,D], slot_map int32");
  const int64_t cap_old = a.dim(0, 0) - 1, cap_new = a.dim(2, 0) - 1;
  const int dim = (int)a.dim(0, 1);
  MREC_REQUIRE(a.dim(2, 1) == dim && a.numel(1) >= cap_old && cap_new >= cap_old, ERR_SHAPE,
               "mrec_hash_move_rows: arena_new must be [C2+1, D] with C2 >= C; slot_map[C]");
  MREC_LAUNCH(hash_move_rows_kernel, grid_for(cdiv((cap_old + 1) * dim, 256), 8), 256, 0, a.stream, a.ptr<float>(0),
              cap_old, a.ptr<int32_t>(1), a.ptr<float>(2), cap_new, dim);
  return check_launch("hash_move_rows");
}
