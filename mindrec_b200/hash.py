"""MapParameter and HashEmbeddingLookup over the GPU open-addressing hash table (K6).

    MapParameter          mindspore.experimental.MapParameter as used at mindspore_rec/ops/embedding.py:136-144
                          and README.md:176-195 (get / put / erase / get_keys / get_values / export_data)
    HashEmbeddingLookup   mindspore_rec/ops/embedding.py:47-206 (same constructor arguments and errors)

The table maps key -> slot; rows live in [C+1, D] arenas (row C = default row).  `get` is one
find-or-insert kernel plus the K1 gather; optimizers update rows by slot index with the same fused kernels
as dense tables (nn.LazyAdam / nn.FTRL accept a MapParameter via `hash_param.as_parameter()`).
"""
import sys

import torch

from . import _lib, ops
from .nn import Parameter

EMPTY_KEY, ERASED_KEY = -1, -2
MAX_SIZE = sys.maxsize


def _pow2_at_least(n):
    c = 8
    while c < n:
        c <<= 1
    return c


class MapParameter:
    """MapParameter(key_dtype=int32, value_dtype=float32, value_shape=1, key_tensor=None, value_tensor=None,
    default_value='normal', permit_filter_value=1, evict_filter_value=MAX_SIZE, name=None, requires_grad=True).

    `capacity` (slots, rounded to a power of two) is the INITIAL size: `maybe_grow` / `grow` rebuild the table into a
    larger one (mrec_hash_rehash + mrec_hash_move_rows), as upstream's cuco dynamic_map does, so online training with
    drifting keys does not end in the overflow flag.  Sibling arenas (optimizer state) register through `add_arena`
    and are re-read through `arena(i)` after a growth."""

    def __init__(self, key_dtype=torch.int32, value_dtype=torch.float32, value_shape=1, key_tensor=None,
                 value_tensor=None, default_value="normal", permit_filter_value=1, evict_filter_value=MAX_SIZE,
                 name=None, requires_grad=True, capacity=1 << 20, device="cuda", seed=0):
        if key_dtype not in (torch.int32, torch.int64):
            raise TypeError("For 'MapParameter', key_dtype must be int32 or int64, got %r" % (key_dtype,))
        if value_dtype != torch.float32:
            raise TypeError("For 'MapParameter', value_dtype must be float32, got %r" % (value_dtype,))
        if not isinstance(permit_filter_value, int) or permit_filter_value <= 0:
            raise ValueError("For 'MapParameter', permit_filter_value must be a positive int")
        if not isinstance(evict_filter_value, int) or evict_filter_value <= 0:
            raise ValueError("For 'MapParameter', evict_filter_value must be a positive int")
        self.key_dtype = key_dtype
        self.value_shape = (value_shape,) if isinstance(value_shape, int) else tuple(value_shape)
        if len(self.value_shape) != 1:
            raise ValueError("only 1-D value_shape is supported")
        self.dim = self.value_shape[0]
        self.name = name
        self.requires_grad = requires_grad
        self.device = torch.device(device)
        self.capacity = _pow2_at_least(int(capacity))
        self.permit_filter_value = permit_filter_value
        self.evict_filter_value = evict_filter_value
        c, dev = self.capacity, self.device
        self.tkeys = torch.full((c,), EMPTY_KEY, dtype=torch.int64, device=dev)
        self.meta = torch.zeros(c, dtype=torch.int64, device=dev)
        self.state = torch.zeros(8, dtype=torch.int32, device=dev)
        self.cfg = torch.tensor([permit_filter_value, min(evict_filter_value, 2 ** 31 - 2)], dtype=torch.int32,
                                device=dev)
        self.default_value = default_value
        self.values = torch.zeros((c + 1, self.dim), dtype=torch.float32, device=dev)
        if isinstance(default_value, str):
            if default_value == "normal":
                mode, sigma = 1, 0.01       # initializer('normal') = N(0, 0.01^2); candidates read zeros
            elif default_value in ("zeros", "zero"):
                mode, sigma = 0, 0.0
            elif default_value in ("ones", "one"):
                mode, sigma = 0, 0.0
                self.values[c].fill_(1.0)
            else:
                raise ValueError("unsupported default_value %r" % (default_value,))
        else:
            mode, sigma = 0, 0.0
            self.values[c].copy_(torch.as_tensor(default_value, dtype=torch.float32).expand(self.dim))
        self._rng = torch.tensor([seed, mode], dtype=torch.int64, device=dev)
        self._rng_copy = torch.tensor([seed, 0], dtype=torch.int64, device=dev)
        self._sigma = torch.tensor([sigma], dtype=torch.float32, device=dev)
        self._arenas = []          # sibling [C+1, D'] arenas, initialised from their own default row
        self._scratch = {}
        # incremental export: keys removed since the last export, and the step of that export
        self._erase_log = torch.empty(max(1024, min(c, 1 << 20)), dtype=torch.int64, device=dev)
        self._exported_at = torch.zeros(1, dtype=torch.int32, device=dev)
        if key_tensor is not None:
            self.put(key_tensor, value_tensor)

    # ---- plumbing --------------------------------------------------------------------------------
    def _bufs(self, n):
        b = self._scratch.get(n)
        if b is None:
            dev = self.device
            b = (torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.int32, device=dev),
                 torch.zeros(1, dtype=torch.int32, device=dev))
            self._scratch[n] = b
        return b

    def _table(self):
        return [self.tkeys, self.meta, self.state, self.cfg]

    def add_arena(self, fill=0.0, dim=None):
        """A sibling arena (e.g. Adam moments): rows of new keys are set to `fill`.  Returns the arena tensor; holders
        that must survive a growth keep the index `len(self._arenas) - 1` and read `arena(i)` instead."""
        a = torch.full((self.capacity + 1, dim or self.dim), float(fill), dtype=torch.float32, device=self.device)
        self._arenas.append(a)
        return a

    def arena(self, i):
        return self._arenas[i]

    # ---- growth ------------------------------------------------------------------------------------
    MAX_LOAD = 0.6

    def grow(self, capacity=None):
        """Rebuild into `capacity` slots (default 2x): keys, admission / eviction words and the rows of every arena
        move; tombstones vanish.  Slot indices handed out before the call are invalid afterwards."""
        c_old = self.capacity
        c_new = _pow2_at_least(int(capacity or 2 * c_old))
        if c_new <= c_old:
            return self
        dev = self.device
        tkeys = torch.full((c_new,), EMPTY_KEY, dtype=torch.int64, device=dev)
        meta = torch.zeros(c_new, dtype=torch.int64, device=dev)
        slot_map = torch.empty(c_old, dtype=torch.int32, device=dev)
        _lib.aot_call("mrec_hash_rehash", [self.tkeys, self.meta, self.state, tkeys, meta, slot_map])

        def move(old):
            new = torch.empty((c_new + 1, old.shape[1]), dtype=torch.float32, device=dev)
            _lib.aot_call("mrec_hash_move_rows", [old, slot_map, new])
            return new
        self.values = move(self.values)
        self._arenas = [move(a) for a in self._arenas]
        self.tkeys, self.meta, self.capacity = tkeys, meta, c_new
        self._scratch = {}
        self.grown = getattr(self, "grown", 0) + 1
        return self

    def maybe_grow(self, incoming):
        """Grow (host-side check: one read of the occupancy counter) so that `incoming` more keys keep the load factor
        at or below MAX_LOAD.  Call it before a lookup, outside captured graphs."""
        occupied = int(self.state[4].item())
        need = (occupied + int(incoming)) / self.MAX_LOAD
        if need > self.capacity:
            self.grow(_pow2_at_least(int(need) + 1))
        return self

    def _init_new(self, new_slots, new_count):
        d = ops._dummy(self.device)
        _lib.aot_call("mrec_hash_init_rows", [self.values, new_slots, new_count, self.tkeys, self._rng,
                                              self._sigma, d])
        for a in self._arenas:
            _lib.aot_call("mrec_hash_init_rows", [a, new_slots, new_count, self.tkeys, self._rng_copy,
                                                  self._sigma, d])

    def as_parameter(self):
        """The C resident rows of the value arena as a dense Parameter: optimizers address it by slot index.  Row C
        (the default row every unadmitted / overflowed key reads) is NOT part of it, so a gradient that arrives for
        slot C is dropped by the bounded dedup instead of moving the default row."""
        return Parameter(self.values[:self.capacity], name=self.name or "map_parameter")

    def arena_rows(self, arena):
        """The C slot rows of a sibling arena (same slicing as as_parameter)."""
        return arena[:self.capacity]

    # ---- MapTensor ops ---------------------------------------------------------------------------
    def lookup_slots(self, key, insert_default_value=True):
        """key -> slot index (int32, C = default row); new keys are admitted / initialised on the way."""
        flat = key.reshape(-1)
        slots, new_slots, new_count = self._bufs(flat.numel())
        if insert_default_value:
            _lib.aot_call("mrec_hash_find_or_insert", [flat] + self._table() + [slots, new_slots, new_count])
            self._init_new(new_slots, new_count)
        else:
            _lib.aot_call("mrec_hash_find", [flat] + self._table() + [slots])
        return slots

    def get(self, key, insert_default_value=True):
        """MapTensorGet: rows for `key` (shape key.shape + (D,))."""
        slots = self.lookup_slots(key, insert_default_value)
        return ops.gather(self.values, slots).view(tuple(key.shape) + (self.dim,))

    def put(self, key, value):
        """MapTensorPut: insert or overwrite."""
        flat = key.reshape(-1)
        slots, new_slots, new_count = self._bufs(flat.numel())
        _lib.aot_call("mrec_hash_insert", [flat] + self._table() + [slots, new_slots, new_count])
        for a in self._arenas:  # fresh optimizer state for keys that were not resident
            _lib.aot_call("mrec_hash_init_rows", [a, new_slots, new_count, self.tkeys, self._rng_copy,
                                                  self._sigma, ops._dummy(self.device)])
        vals = value.reshape(flat.numel(), self.dim).to(torch.float32).contiguous()
        _lib.aot_call("mrec_hash_scatter_rows", [self.values, slots, vals, ops._dummy(self.device)])
        return self

    def erase(self, key):
        """MapTensorErase."""
        flat = key.reshape(-1)
        slots, _, _ = self._bufs(flat.numel())
        _lib.aot_call("mrec_hash_erase", [flat] + self._table() + [slots, self._erase_log])
        return self

    def evict(self):
        """Erase keys not looked up for more than evict_filter_value calls (README.md:182-183)."""
        _lib.aot_call("mrec_hash_evict", self._table() + [ops._dummy(self.device), self._erase_log])
        return self

    def _export(self, since=None):
        c = self.capacity
        keys_out = torch.empty(c, dtype=torch.int64, device=self.device)
        slots_out = torch.empty(c, dtype=torch.int32, device=self.device)
        count = torch.zeros(1, dtype=torch.int32, device=self.device)
        _lib.aot_call("mrec_hash_export", self._table() + ([since] if since is not None else []) +
                      [keys_out, slots_out, count])
        n = int(count.item())
        order = torch.argsort(keys_out[:n])  # deterministic presentation order
        return keys_out[:n][order], slots_out[:n][order]

    def get_keys(self):
        return self._export()[0].to(self.key_dtype)

    def get_values(self):
        return ops.gather(self.values, self._export()[1].contiguous())

    def get_data(self):
        k, s = self._export()
        return k.to(self.key_dtype), ops.gather(self.values, s.contiguous())

    STATUS_NORMAL, STATUS_MODIFIED, STATUS_ERASED = 0, 1, 2

    def export_data(self, incremental=False):
        """(keys, values, statuses).  Full export: every resident key, status 0.  incremental=True: only what
        changed since the previous export_data call — keys looked up / put since then (status 1, current values)
        and keys erased or evicted since then that are not resident again (status 2, zero values).  If the erase
        log overflowed in between, the export falls back to a full one (all statuses 0: replace the replica)."""
        st = self.state.tolist()
        log_n, log_ovf = st[5], st[6]
        if incremental and not log_ovf:
            k, s = self._export(since=self._exported_at)
            v = ops.gather(self.values, s.contiguous())
            status = torch.full((k.numel(),), self.STATUS_MODIFIED, dtype=torch.int32, device=self.device)
            if log_n:
                gone = torch.unique(self._erase_log[:log_n])
                # a key erased and inserted again since the last export is simply "modified"
                still = self.lookup_slots(gone.to(self.key_dtype), insert_default_value=False) == self.capacity
                gone = gone[still]
                k = torch.cat([k, gone])
                v = torch.cat([v, torch.zeros((gone.numel(), self.dim), dtype=torch.float32, device=self.device)])
                status = torch.cat([status, torch.full((gone.numel(),), self.STATUS_ERASED, dtype=torch.int32,
                                                       device=self.device)])
        else:
            k, v = self.get_data()
            status = torch.zeros(k.numel(), dtype=torch.int32, device=self.device)
        self._exported_at.copy_(self.state[1:2])
        self.state[5:7] = 0
        return k.to(self.key_dtype), v, status

    def import_data(self, data):
        """Apply a full or incremental export: erase the status-2 keys, put the rest."""
        keys, values = data[0], data[1]
        if len(data) > 2 and data[2] is not None and bool((data[2] == self.STATUS_ERASED).any()):
            gone = data[2] == self.STATUS_ERASED
            self.erase(keys[gone].contiguous())
            keys, values = keys[~gone].contiguous(), values[~gone].contiguous()
        if keys.numel() == 0:
            return self
        return self.put(keys, values)

    def __getitem__(self, key):
        return self.get(key)

    def __setitem__(self, key, value):
        self.put(key, value)

    def __len__(self):
        return int(self.state[0].item())

    @property
    def overflowed(self):
        return bool(self.state[3].item())


class HashEmbeddingLookup:
    """HashEmbeddingLookup(embedding_size, key_dtype=int32, param_init='normal', sparse=True, max_norm=None,
    permit_filter_value=1, evict_filter_value=sys.maxsize, vocab_cache_size=0)  — embedding.py:85-95."""

    def __init__(self, embedding_size, key_dtype=torch.int32, param_init="normal", sparse=True, max_norm=None,
                 permit_filter_value=1, evict_filter_value=MAX_SIZE, vocab_cache_size=0, capacity=1 << 20,
                 device="cuda", seed=0, auto_grow=True):
        if not isinstance(sparse, bool):
            raise TypeError("For 'HashEmbeddingLookup', the type of 'sparse' should be bool, but got %s"
                            % type(sparse).__name__)
        if not isinstance(vocab_cache_size, int) or isinstance(vocab_cache_size, bool) or vocab_cache_size < 0:
            raise ValueError("For 'HashEmbeddingLookup', 'vocab_cache_size' must be a non-negative int, got %r"
                             % (vocab_cache_size,))
        if vocab_cache_size > 0:
            # embedding.py:104-110: cache mode only exists under parameter-server training
            raise RuntimeError("The configuration of 'vocab_cache_size' is greater than 0 means enable embedding "
                               "cache mode, this mode only support in parameter server training mode, please "
                               "enable ps mode by 'context.set_ps_context(enable_ps=True)'")
        if not isinstance(embedding_size, int) or isinstance(embedding_size, bool) or embedding_size <= 0:
            raise ValueError("For 'HashEmbeddingLookup', 'embedding_size' must be a positive int, got %r"
                             % (embedding_size,))
        if max_norm is not None and (not isinstance(max_norm, float) or max_norm <= 0):
            raise ValueError("For 'HashEmbeddingLookup', 'max_norm' must be a positive float")
        self.forward_unique = sparse
        self.embedding_size = embedding_size
        self.max_norm = max_norm
        self.embedding_table = MapParameter(key_dtype=key_dtype, value_dtype=torch.float32,
                                            value_shape=(embedding_size,), default_value=param_init,
                                            name="embedding_table", permit_filter_value=permit_filter_value,
                                            evict_filter_value=evict_filter_value, capacity=capacity,
                                            device=device, seed=seed)
        self.embedding_table.unique = self.forward_unique
        self.last_slots = None
        # auto_grow: size the table before every lookup so that it cannot overflow (one host read per call; switch
        # it off — and size `capacity` up front — to capture lookups in a CUDA graph)
        self.auto_grow = auto_grow

    def __call__(self, indices):
        return self.construct(indices)

    def construct(self, indices):
        """embedding.py:184-206.  The reference deduplicates first (Unique -> MapTensorGet -> Gather by the
        inverse index); the table here resolves duplicates itself (all copies of a key probe to one slot), so
        the keys go straight to find-or-insert and the rows come from one gather by slot — same output."""
        table = self.embedding_table
        if self.auto_grow:
            table.maybe_grow(indices.numel())
        slots = table.lookup_slots(indices, insert_default_value=True)
        self.last_slots = slots.view(indices.shape)
        out = ops.gather(table.values, slots).view(tuple(indices.shape) + (self.embedding_size,))
        if self.max_norm is not None:
            norm = out.norm(dim=-1, keepdim=True).clamp_min(1e-12)
            out = out * torch.clamp(self.max_norm / norm, max=1.0)
        return out
