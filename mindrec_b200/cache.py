"""Two-level embedding cache: the full table (and its optimizer state) lives in pinned host memory,
the GPU holds the rows of the current working set.

    nn.EmbeddingLookup(..., vocab_cache_size=N) / HashEmbeddingLookup(vocab_cache_size=N)
        mindspore_rec/ops/embedding.py:99-131,163-182 — "the Host ... is responsible for mapping the id to the index on
        the device side, and device side uses Tensor for storage and computation"; README.md:146-150 (Device <-> Local
        Host <-> Remote Host <-> SSD; this module is the first two levels)
    models/wide_deep/src/wide_and_deep.py:176-230 (vocab_cache_size plumbing), train_and_eval_parameter_server_*.py

Here the id -> device-row map is the GPU hash table itself (mindrec_b200.hash.MapParameter with the id as key), so a
lookup is the same find-or-insert + gather as a dynamic embedding; the only new step is what happens to a MISS: the rows
of the newly resident ids — weights and every optimizer-state arena — are fetched from the host tier (one pinned gather
+ one H2D copy per arena) instead of being initialised.  When the next batch would push the load factor over MAX_LOAD,
the cache is flushed: every resident row and its state are written back to the host tier (one D2H copy per arena) and
the map is cleared.  Optimizers see a MapParameter and update rows by slot (nn.LazyAdam / nn.FTRL), so a cached table
trains exactly like a resident one; `sync_to_host()` makes the host copy current for checkpoints.
A miss costs a host round trip by construction (the host owns the cold rows), so this path is eager, not graph-captured.
"""
import torch

from . import _lib, ops
from .hash import EMPTY_KEY, MapParameter, _pow2_at_least
from .nn import _init_table


class CachedMapParameter(MapParameter):
    """MapParameter whose misses are served from a pinned host table instead of an initializer."""

    def __init__(self, vocab_size, dim, cache_rows, param_init="normal", device="cuda", key_dtype=torch.int32, seed=0,
                 generator=None, name="embedding_table"):
        cap = _pow2_at_least(int(cache_rows / self.MAX_LOAD) + 1)
        super().__init__(key_dtype=key_dtype, value_shape=dim, default_value="zeros", capacity=cap, device=device, seed=seed,
                         name=name)
        self.vocab_size, self.cache_rows = int(vocab_size), int(cache_rows)
        init = _init_table((vocab_size, dim), param_init, device, generator)        # same values as a resident table
        self.host_values = init.cpu().pin_memory()
        del init
        self.host_arenas = []
        self.flushes = 0
        self.misses = 0

    def add_arena(self, fill=0.0, dim=None):
        a = super().add_arena(fill, dim)
        self.host_arenas.append(torch.full((self.vocab_size, dim or self.dim), float(fill), dtype=torch.float32).pin_memory())
        return a

    # a newly resident id: rows come from the host tier
    def _init_new(self, new_slots, new_count):
        n = int(new_count.item())                       # host sync: serving a miss is host work anyway
        if n == 0:
            return
        self.misses += n
        slots = new_slots[:n].contiguous()
        keys = self.tkeys[slots.long()].cpu()
        if bool(((keys < 0) | (keys >= self.vocab_size)).any()):
            raise IndexError("CachedMapParameter: an id outside [0, vocab_size) was looked up")
        d = ops._dummy(self.device)
        for arena, host in [(self.values, self.host_values)] + list(zip(self._arenas, self.host_arenas)):
            rows = host.index_select(0, keys).pin_memory().to(self.device, non_blocking=True)
            _lib.aot_call("mrec_hash_scatter_rows", [arena, slots, rows, d])

    def sync_to_host(self):
        """Write every resident row (and its optimizer state) back to the host tier."""
        k, s = self._export()
        if k.numel() == 0:
            return 0
        kc = k.cpu()
        s = s.contiguous()
        for arena, host in [(self.values, self.host_values)] + list(zip(self._arenas, self.host_arenas)):
            host.index_copy_(0, kc, ops.gather(arena, s).cpu())
        return int(k.numel())

    def flush(self):
        """Write back, then empty the device cache."""
        self.sync_to_host()
        self.tkeys.fill_(EMPTY_KEY)
        self.meta.zero_()
        self.state[0] = 0
        self.state[2:5] = 0
        self.flushes += 1

    def maybe_flush(self, incoming):
        """Make room for a lookup of `incoming` ids (an upper bound of its distinct ids)."""
        incoming = min(int(incoming), self.vocab_size)
        if incoming > self.MAX_LOAD * self.capacity:
            raise ValueError("one lookup of %d ids does not fit the cache (%d slots at load %.1f): raise vocab_cache_size"
                             % (incoming, self.capacity, self.MAX_LOAD))
        if int(self.state[4].item()) + incoming > self.MAX_LOAD * self.capacity:
            self.flush()

    def maybe_grow(self, incoming):                      # a cache does not grow: it flushes
        return self.maybe_flush(incoming)

    def grow(self, capacity=None):
        raise RuntimeError("the device cache has a fixed size (vocab_cache_size); it flushes instead of growing")

    def full_table(self):
        """The current [V, D] table (host tier brought up to date)."""
        self.sync_to_host()
        return self.host_values


class CachedEmbeddingLookup:
    """nn.EmbeddingLookup(vocab_size, embedding_size, ..., vocab_cache_size=N) in cache mode."""

    def __init__(self, vocab_size, embedding_size, vocab_cache_size, param_init="normal", sparse=True, max_norm=None,
                 device="cuda", name="embedding_table", generator=None, seed=0):
        if not isinstance(vocab_cache_size, int) or vocab_cache_size <= 0:
            raise ValueError("For 'EmbeddingLookup', 'vocab_cache_size' must be a positive int in cache mode")
        if vocab_cache_size > vocab_size:
            vocab_cache_size = vocab_size
        self.vocab_size, self.embedding_size, self.vocab_cache_size = vocab_size, embedding_size, vocab_cache_size
        self.sparse, self.max_norm = sparse, max_norm
        self.embedding_table = CachedMapParameter(vocab_size, embedding_size, vocab_cache_size, param_init, device,
                                                  generator=generator, seed=seed, name=name)
        self.auto_grow = True                            # cells.WideDeepModel: "prepare the table before the lookup"
        self.last_slots = None

    def __call__(self, indices):
        return self.construct(indices)

    def construct(self, indices):
        t = self.embedding_table
        t.maybe_flush(indices.numel())
        slots = t.lookup_slots(indices, insert_default_value=True)
        self.last_slots = slots.view(indices.shape)
        out = ops.gather(t.values, slots).view(tuple(indices.shape) + (self.embedding_size,))
        if self.max_norm is not None:
            norm = out.norm(dim=-1, keepdim=True).clamp_min(1e-12)
            out = out * torch.clamp(self.max_norm / norm, max=1.0)
        return out
