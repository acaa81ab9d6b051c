"""Row-sharded embedding tables over the GPUs of one box: dedup -> all-to-all(keys) -> owner gather ->
all-to-all(rows) forward, all-to-all(gradients) -> owner dedup -> fused row updates backward.

This is the B200 form of nn.EmbeddingLookup(slice_mode=TABLE_ROW_SLICE) under auto-parallel
(models/wide_deep/src/wide_and_deep.py:232-249, train_and_eval_parameter_server_distribute.py:92-96), where
upstream inserts AllGather(ids) / masked local gather / ReduceScatter(rows).  The DenseLayers stay
data-parallel with a mean AllReduce of their gradients (gradients_mean=True,
train_and_eval_distribute.py:135-138).

Ownership: owner = key mod G, local_row = key div G (Zipf-balanced; a contiguous-range split would put the
hot head of every small field on rank 0).  Keys are remapped owner-major, key' = owner * R + local_row, before
the dedup, so one radix sort both deduplicates and buckets by owner: the sorted unique keys form G contiguous
runs whose boundaries (mrec_shard_bounds) are the all-to-all split sizes, and key' mod R is already the owner's
local row.  One process per GPU; collectives are torch.distributed (NCCL on the box, gloo in the CPU tests of
the host logic).  The split sizes are data dependent, so each step reads G+1 ints back to the host.
"""
import os

import torch
import torch.distributed as dist

from . import ops

# The device runtime the streams / events / pinned buffers come from.  There is one path (CUDA); the gloo CPU tests of
# the exchange logic substitute stand-ins for `ops` and `_cu` from the test side (tests/oracle_kernels.py,
# tests/fake_cuda.py), like tests/fake_peer_ops.py does for the peer protocol.
_cu = torch.cuda


def _pinned(t):
    return t.pin_memory()

# MREC_SHARDED_GRAPH=0 keeps the DenseLayer segment eager; MREC_SHARDED_AHEAD=0 plans every batch in line.
_ENV_GRAPH = os.environ.get("MREC_SHARDED_GRAPH", "1") != "0"
_ENV_AHEAD = os.environ.get("MREC_SHARDED_AHEAD", "1") != "0"
# MREC_SHARDED_PEER=0 returns rows with NCCL all-to-all instead of the fused gather + NVLink peer store.
_ENV_PEER = os.environ.get("MREC_SHARDED_PEER", "1") != "0"


class _RawCuda:
    """__cuda_array_interface__ view of a raw device allocation (lets torch wrap memory from mrec_peer_alloc)."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


class PeerLanding:
    """Per-rank landing buffers [cap, D] that every other rank of the node can store into (CUDA IPC).

    Allocated with cudaMalloc through the library (exportable with cudaIpcGetMemHandle), handles exchanged once
    over the process group, peers mapped with cudaIpcOpenMemHandle (which also enables peer access)."""

    def __init__(self, cap_rows, dims, device, group):
        import ctypes
        from . import _lib
        lib = _lib.lib()
        lib.mrec_peer_alloc.restype = ctypes.c_void_p
        lib.mrec_peer_alloc.argtypes = [ctypes.c_size_t]
        lib.mrec_ipc_open_handle.restype = ctypes.c_void_p
        lib.mrec_ipc_open_handle.argtypes = [ctypes.c_char_p]
        lib.mrec_ipc_get_handle.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        self.buffers, self.ptr_tensors, handles = [], [], []
        for d in dims:
            ptr = lib.mrec_peer_alloc(cap_rows * d * 4)
            if not ptr:
                raise RuntimeError("mrec_peer_alloc failed: " + _lib.last_error())
            h = ctypes.create_string_buffer(64)
            if lib.mrec_ipc_get_handle(ctypes.c_void_p(ptr), h) != 0:
                raise RuntimeError("mrec_ipc_get_handle failed: " + _lib.last_error())
            handles.append((ptr, h.raw))
            self.buffers.append(torch.as_tensor(_RawCuda(ptr, (cap_rows, d)), device=device))
        gathered = [None] * world
        dist.all_gather_object(gathered, [h for _, h in handles], group=group)
        for j, (ptr, _) in enumerate(handles):
            ptrs = []
            for r in range(world):
                if r == rank:
                    ptrs.append(ptr)
                else:
                    p = lib.mrec_ipc_open_handle(gathered[r][j])
                    if not p:
                        raise RuntimeError("mrec_ipc_open_handle failed: " + _lib.last_error())
                    ptrs.append(p)
            self.ptr_tensors.append(torch.tensor(ptrs, dtype=torch.int64, device=device))


class ShardPlan:
    """Pure index arithmetic of the exchange (testable without a GPU)."""

    def __init__(self, vocab_size, world_size):
        self.vocab_size = int(vocab_size)
        self.world = int(world_size)
        self.rows_per_rank = (self.vocab_size + self.world - 1) // self.world
        if self.rows_per_rank * self.world >= 2 ** 31:
            raise ValueError("owner-major key space exceeds int32: use int64 ids")

    def remap(self, ids):
        """key -> owner * R + local_row; out-of-range keys -> G * R (dropped by the bounded dedup)."""
        g, r = self.world, self.rows_per_rank
        ok = (ids >= 0) & (ids < self.vocab_size)
        km = (ids % g) * r + torch.div(ids, g, rounding_mode="floor")
        return torch.where(ok, km, torch.full_like(ids, g * r))

    def edges(self, like):
        return (torch.arange(self.world + 1, device=like.device, dtype=like.dtype) * self.rows_per_rank)

    def owner_of(self, key):
        return key % self.world

    def local_row(self, key):
        return key // self.world


def _a2a(out, inp, out_splits, in_splits, group):
    dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)
    return out


class BatchPlan:
    """The data-dependent part of one batch's exchange: dedup result, per-owner split sizes.

    Built by ShardedWideDeepTables.plan_batch — possibly one step ahead, on a side stream with its own
    communicator — and finalised (one host read of G*(G+1) ints) when the batch is consumed."""

    __slots__ = ("ids", "uq", "bounds_host", "event", "send", "recv", "n_u", "n_r", "allb", "dst_off", "src_off")

    def finalize(self, rank, world):
        if self.send is not None:
            return self
        if self.event is not None:
            self.event.synchronize()
        b = self.bounds_host.view(world, world + 1).tolist()
        self.send = [b[rank][r + 1] - b[rank][r] for r in range(world)]
        self.recv = [b[r][rank + 1] - b[r][rank] for r in range(world)]
        self.n_u, self.n_r = b[rank][world], sum(self.recv)
        return self


class ShardedWideDeepTables:
    """The wide (dim 1) and deep (dim D) tables of Wide&Deep, row-sharded, with their FTRL / LazyAdam state."""

    def __init__(self, vocab_size, emb_dim, device, group=None, seed=1, sens=1024.0, init_std=0.01):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        # planning runs on a side stream: it needs its own communicator (NCCL serialises per communicator)
        self.plan_group = dist.new_group() if (dist.is_initialized() and self.world > 1) else None
        self.plan = ShardPlan(vocab_size, self.world)
        self.dim = emb_dim
        self.device = torch.device(device)
        self.plan_stream = _cu.Stream(device=self.device)
        self.owner_stream = _cu.Stream(device=self.device)
        r = self.plan.rows_per_rank
        gen = torch.Generator(device=self.device)
        gen.manual_seed(seed * 1000 + self.rank)
        self.wide = torch.empty((r, 1), dtype=torch.float32, device=self.device).normal_(0, init_std, generator=gen)
        self.deep = torch.empty((r, emb_dim), dtype=torch.float32, device=self.device).normal_(0, init_std, generator=gen)
        self.acc = torch.ones_like(self.wide)
        self.lin = torch.zeros_like(self.wide)
        self.m = torch.zeros_like(self.deep)
        self.v = torch.zeros_like(self.deep)
        # gradients_mean=True: the owner sums contributions of all ranks, each already divided by G
        self.adam_hyper = ops.adam_hyper(3.5e-4, eps=1e-8, loss_scale=sens * self.world, device=self.device)
        self.ftrl_hyper = ops.ftrl_hyper(5e-2, l1=1e-8, l2=1e-8, loss_scale=sens * self.world, device=self.device)
        self._bound_like = torch.empty((self.world * r, 0), device=self.device)
        self._vocab_like = torch.empty((vocab_size, 0), device=self.device)
        self._owners_like = torch.empty((self.world, r, 0), device=self.device)   # shape carrier: G, R
        self._edges = {}
        self._ctx = None
        self._uq = {}
        self._gs = None
        self._slot = 0
        self._bufs = {}
        self.peer = None
        self._peer_tried = False
        self._bar = torch.zeros(1, dtype=torch.float32, device=self.device)
        n_b = self.world * (self.world + 1)
        self._pinned = [_pinned(torch.empty(n_b, dtype=torch.int32)) for _ in range(2)]

    def _unique(self, key, table_like, tag):
        """mrec_unique into cached, geometrically grown output buffers (no per-step allocation)."""
        n = key.numel()
        base = self._uq.get(tag)
        if base is None or base.n < n:
            base = ops.UniqueResult(max(n, int(1.5 * base.n) if base else n), key.dtype, key.device)
            self._uq[tag] = base
        return ops.unique(key, table_like=table_like, result=base if base.n == n else base.sliced(n),
                          ws_tag="unique_" + tag)

    def _buf(self, tag, rows, cols, dtype=torch.float32):
        """Grow-only [rows, cols] scratch buffer (variable per-step sizes without reallocation)."""
        b = self._bufs.get(tag)
        if b is None or b.shape[0] < rows or b.shape[1] != cols:
            b = torch.empty((max(rows, int(1.5 * b.shape[0]) if b is not None else rows, 1), cols), dtype=dtype,
                            device=self.device)
            self._bufs[tag] = b
        return b

    # ---- planning (may run one batch ahead) ---------------------------------------------------------
    def plan_batch(self, ids, ahead=False):
        """Dedup + bucket the batch's keys and exchange the bucket bounds.  With ahead=True the work is queued
        on the side stream (double-buffered outputs) and the caller keeps enqueuing the current step."""
        k, g = ops, self.world
        plan = BatchPlan()
        plan.ids, plan.send = ids, None
        slot = self._slot
        self._slot ^= 1
        side = self.plan_stream if ahead else None
        if side is not None:
            side.wait_stream(_cu.current_stream())
        ctx = _cu.stream(side) if side is not None else _Null()
        with ctx:
            key = k.shard_remap(ids, self._vocab_like, self._owners_like)
            plan.uq = self._unique(key, self._bound_like, "plan%d" % slot)
            edges = self._edges.get(key.dtype)
            if edges is None:
                edges = self._edges[key.dtype] = self.plan.edges(key)
            bounds = k.shard_bounds(plan.uq.uniq, plan.uq.count, edges)
            if g > 1:
                allb = torch.empty(g * (g + 1), dtype=torch.int32, device=ids.device)
                dist.all_gather_into_tensor(allb, bounds, group=self.plan_group)
            else:
                allb = bounds
            plan.allb = allb
            if g > 1:
                # offsets of the fused gather + peer store, computed here so they are off the critical path:
                # dst_off[s] = where rank s's bucket for this owner starts in s's landing buffer,
                # src_off    = prefix sums of what every rank asks of this owner
                ab = allb.view(g, g + 1)
                plan.dst_off = ab[:, self.rank].contiguous()
                plan.src_off = torch.zeros(g + 1, dtype=torch.int32, device=ids.device)
                plan.src_off[1:] = torch.cumsum(ab[:, self.rank + 1] - ab[:, self.rank], 0)
            else:
                plan.dst_off = plan.src_off = None
            plan.bounds_host = self._pinned[slot]
            plan.bounds_host.copy_(allb, non_blocking=True)
            plan.event = _cu.Event()
            plan.event.record()
        return plan

    # ---- forward ----------------------------------------------------------------------------------
    def lookup(self, plan, wts, wide_bias, deep_out, wide_out):
        """Fill deep_out [B, F*D] (fp32 | fp16) and wide_out [B,1] for the planned batch."""
        k, g = ops, self.world
        plan.finalize(self.rank, g)                                   # the step's one host read-back
        _cu.current_stream().wait_event(plan.event)
        uq, n_u, n_r, send, recv = plan.uq, plan.n_u, plan.n_r, plan.send, plan.recv
        b, f = plan.ids.shape
        local_rows_send = (uq.uniq[:n_u] % self.plan.rows_per_rank).contiguous()
        rows_recv = self._buf("rows_recv", n_r, 1, plan.ids.dtype).view(-1)[:n_r]
        if g > 1:
            _a2a(rows_recv, local_rows_send, recv, send, self.group)
        else:
            rows_recv.copy_(local_rows_send)
        # owner side: return the requested rows of both tables
        if g > 1 and _ENV_PEER and not self._peer_tried:
            self._peer_tried = True
            try:                                   # landing buffers sized for the worst case U = N
                self.peer = PeerLanding(uq.n, (self.dim, 1), self.device, self.group)
            except Exception as exc:               # no IPC on this box: fall back to NCCL all-to-all
                import warnings
                warnings.warn("peer landing buffers unavailable (%s): using NCCL all-to-all" % (exc,))
                self.peer = None
        if g > 1 and self.peer is not None:
            # fused gather + NVLink peer store straight into each requester's landing buffer, then a
            # stream-ordered barrier publishes everybody's stores (the next key all-to-all orders re-use)
            k.gather_to_peers(self.deep, rows_recv, self.peer.ptr_tensors[0], plan.dst_off, plan.src_off)
            k.gather_to_peers(self.wide, rows_recv, self.peer.ptr_tensors[1], plan.dst_off, plan.src_off)
            dist.all_reduce(self._bar, group=self.group)
            got_deep, got_wide = self.peer.buffers
        else:
            deep_rows = k.gather(self.deep, rows_recv, out=self._buf("deep_rows", n_r, self.dim)[:n_r])
            wide_rows = k.gather(self.wide, rows_recv, out=self._buf("wide_rows", n_r, 1)[:n_r])
            got_deep = self._buf("got_deep", n_u, self.dim)
            got_wide = self._buf("got_wide", n_u, 1)
            if g > 1:
                _a2a(got_deep[:n_u], deep_rows, send, recv, self.group)
                _a2a(got_wide[:n_u], wide_rows, send, recv, self.group)
            else:
                got_deep[:n_u].copy_(deep_rows)
                got_wide[:n_u].copy_(wide_rows)
        # requester side: expand unique rows to lookups with the inverse index, fused with mask / reduce
        inverse = uq.inverse.view(b, f)
        k.gather_masked(got_deep[:max(n_u, 1)], inverse, wts, out=deep_out)
        k.gather_reduce(got_wide[:max(n_u, 1)], inverse, wts, wide_bias, out=wide_out)
        # owner-side dedup of the received rows: needs only rows_recv, so it runs on the side stream
        # underneath the DenseLayer segment (the same row can arrive from several ranks)
        uq2, ev = None, None
        if n_r > 0:
            self.owner_stream.wait_stream(_cu.current_stream())
            with _cu.stream(self.owner_stream):
                uq2 = self._unique(rows_recv, self.deep, "owner")
                ev = _cu.Event()
                ev.record()
        self._ctx = (plan, rows_recv, wts, uq2, ev)
        return wide_out, deep_out

    # ---- backward + update ------------------------------------------------------------------------
    def update(self, delta, gx):
        """delta: [B,1] logit gradient (x sens), gx: [B, F*D] deep-input gradient (x sens, fp32 or fp16)."""
        k, g = ops, self.world
        plan, rows_recv, wts, uq2, ev = self._ctx
        uq, n_u, n_r, send, recv = plan.uq, plan.n_u, plan.n_r, plan.send, plan.recv
        n = uq.n
        mask = wts.reshape(-1)
        gs_deep = k.segment_sum(gx.view(n, self.dim), mask, uq, dim=self.dim, out=self._buf("gs_deep", n, self.dim))
        gs_wide = k.segment_sum(delta, mask, uq, dim=1, out=self._buf("gs_wide", n, 1))
        rg_deep = self._buf("rg_deep", n_r, self.dim)
        rg_wide = self._buf("rg_wide", n_r, 1)
        if g > 1:
            _a2a(rg_deep[:n_r], gs_deep[:n_u], recv, send, self.group)
            _a2a(rg_wide[:n_r], gs_wide[:n_u], recv, send, self.group)
        else:
            rg_deep[:n_r].copy_(gs_deep[:n_u])
            rg_wide[:n_r].copy_(gs_wide[:n_u])
        k.adam_begin_step(self.adam_hyper)
        if n_r == 0:
            return
        # owner side: fused row updates over the (already deduplicated) received rows; the latency-bound FTRL
        # update of the wide shard runs on the side stream beside the LazyAdam update of the deep shard
        main = _cu.current_stream()
        main.wait_event(ev)
        self.owner_stream.wait_stream(main)
        with _cu.stream(self.owner_stream):
            k.sparse_ftrl(self.wide, self.acc, self.lin, self.ftrl_hyper, rg_wide[:n_r], None, uq2)
        k.sparse_lazy_adam(self.deep, self.m, self.v, self.adam_hyper, rg_deep[:n_r], None, uq2)
        main.wait_stream(self.owner_stream)

    # ---- test / checkpoint helper -----------------------------------------------------------------
    def gather_full(self):
        """Reassemble the full [V, 1] / [V, D] tables on every rank (tests, export)."""
        g, r, v = self.world, self.plan.rows_per_rank, self.plan.vocab_size
        if g == 1:
            return self.wide[:v].clone(), self.deep[:v].clone()
        wl = [torch.empty_like(self.wide) for _ in range(g)]
        dl = [torch.empty_like(self.deep) for _ in range(g)]
        dist.all_gather(wl, self.wide, group=self.group)
        dist.all_gather(dl, self.deep, group=self.group)
        wide = torch.stack(wl, 1).reshape(r * g, 1)[:v]     # row k lives at [k // G][k % G]
        deep = torch.stack(dl, 1).reshape(r * g, self.dim)[:v]
        return wide, deep


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class ShardedWideDeepStep:
    """Wide&Deep training step with row-sharded tables and a data-parallel DenseLayer stack.  Same call
    surface as cells.TrainStepWrap (`__call__`, `capture`, `replay`).

    The exchange is eager (its sizes are data dependent); the fixed-shape segment — DenseLayers forward, loss,
    backward, mean all-reduce of the flat gradient, dense Adam — is captured in a CUDA graph.  `replay` takes
    the NEXT batch too and plans its dedup one step ahead on a side stream, so the one host read-back per step
    is already resident when it is needed and the launch pipeline never drains."""

    def __init__(self, batch_size, vocab_size, emb_dim, hidden, device, seed=1, sens=1024.0, fields=39,
                 use_mixed_precision=True, group=None, graph_dense=True, tables_factory=None, dense_storage=None):
        """dense_storage(n) -> (flat, flat_grad): where the DenseLayers' flat parameter / gradient buffers live (the
        peer-memory step puts the gradients into an IPC buffer its in-graph all-reduce reads)."""
        from .nn import DenseStack
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device(device)
        self.sens = float(sens)
        self.fields, self.emb_dim = fields, emb_dim
        self.mixed = use_mixed_precision
        self.tables = (tables_factory() if tables_factory is not None else
                       ShardedWideDeepTables(vocab_size, emb_dim, device, group=group, seed=seed, sens=sens))
        gen = torch.Generator(device=device)
        gen.manual_seed(seed)                      # identical DenseLayer replicas on every rank
        dims = [fields * emb_dim] + list(hidden) + [1]
        self.dense = DenseStack(dims, use_mixed_precision, device, generator=gen, weight_init="normal",
                                bias_init="normal", extra=1,
                                storage=dense_storage(DenseStack.numel(dims, 1)) if dense_storage is not None else None)
        self.wide_b = self.dense.extra
        self.wide_b.normal_(0.0, 0.01, generator=gen)
        # the mean of DistributedGradReducer is folded into the gradient scale: 1 / (sens * G)
        self.dense_hyper = ops.adam_hyper(3.5e-4, eps=1e-8, loss_scale=sens * self.world, device=device)
        self.dense_m = torch.zeros_like(self.dense.flat)
        self.dense_v = torch.zeros_like(self.dense.flat)
        self._sens_t = torch.tensor([self.sens], dtype=torch.float32, device=device)
        self._use_graph = graph_dense and _ENV_GRAPH
        self._graph = None
        self._calls = 0
        self._io = None
        self._slots = None
        self._pending = None
        self._pending_src = None
        self.profile = None          # cells.StepProfile: per-phase CUDA-event timing of eager calls

    # ---- fixed-shape segment -----------------------------------------------------------------------
    def _dense_segment(self, label=None, on_weight_grads=None):
        k = ops
        deep_in, wide_out = self._io["deep_in"], self._io["wide_out"]
        label = self._io["label"] if label is None else label
        deep_out = self.dense.forward(deep_in)
        _, loss, delta, delta16, dsum = k.sigmoid_xent(wide_out, deep_out, label, self._sens_t, out=self._io["loss_out"])
        self.dense.extra_grad.copy_(dsum)            # Wide_b's gradient rides in the DenseLayers' flat buffer
        gx = self.dense.backward(delta16 if delta16.numel() else delta, on_weight_grads=on_weight_grads)
        return loss[0], delta, gx

    def _dense_update(self):
        k = ops
        if self.world > 1:                          # DistributedGradReducer(mean): wide_and_deep.py:455-470
            dist.all_reduce(self.dense.flat_grad, group=self.group)   # eager: capturing it in the graph hangs
        k.adam_begin_step(self.dense_hyper)
        k.adam_dense(self.dense.flat, self.dense_m, self.dense_v, self.dense_hyper, self.dense.flat_grad)

    def _run_dense(self):
        if not self._use_graph:
            return self._dense_segment()
        if self._graph is None:
            if self._calls < 3:                      # eager warm-up (cuBLAS handles, NCCL channels)
                return self._dense_segment()
            try:
                _cu.synchronize()
                g = _cu.CUDAGraph()
                with _cu.graph(g):
                    self._graph_out = self._dense_segment()
                self._graph = g
            except Exception:                        # capture refused: stay eager
                self._use_graph = False
                _cu.synchronize()
                return self._dense_segment()
        self._graph.replay()
        return self._graph_out

    def _ensure_io(self, ids):
        b = ids.shape[0]
        if self._io is None or self._io["label"].shape[0] != b:
            dev = ids.device
            self._io = dict(
                deep_in=torch.empty((b, self.fields * self.emb_dim), device=dev,
                                    dtype=torch.float16 if self.mixed else torch.float32),
                wide_out=torch.empty((b, 1), dtype=torch.float32, device=dev),
                label=torch.empty((b, 1), dtype=torch.float32, device=dev),
                loss_out=(torch.empty((b, 1), dtype=torch.float32, device=dev),
                          torch.empty(1, dtype=torch.float32, device=dev),
                          torch.empty((b, 1), dtype=torch.float32, device=dev),
                          torch.empty((b, 1) if self.mixed else (0,), dtype=torch.float16, device=dev),
                          torch.empty(1, dtype=torch.float32, device=dev)))
            self._graph = None

    def _range(self, name):
        return self.profile.range(name) if self.profile is not None else _Null()

    def _step(self, plan, wts, label):
        self._ensure_io(plan.ids)
        with self._range("exchange_forward"):
            self.tables.lookup(plan, wts, self.wide_b, self._io["deep_in"], self._io["wide_out"])
        with self._range("dense_fwd_bwd"):
            self._io["label"].copy_(label)
            loss, delta, gx = self._run_dense()
        with self._range("dense_allreduce_adam"):
            self._dense_update()
        with self._range("exchange_backward_update"):
            self.tables.update(delta, gx)
        self._calls += 1
        return loss, loss

    def __call__(self, ids, wts, label):
        with self._range("plan_dedup"):
            plan = self.tables.plan_batch(ids)
        return self._step(plan, wts, label)

    # ---- bench-facing surface (double-buffered inputs, dedup planned one batch ahead) ----------------
    def capture(self, ids, wts, label, warmup=3):
        self._slots = [tuple(t.clone() for t in (ids, wts, label)) for _ in range(2)]
        self._cur = 0
        for _ in range(warmup + 1):
            self(*self._slots[0])
        self._pending = None
        return self._slots[0]

    def replay(self, ids=None, wts=None, label=None, next_batch=None):
        cur = self._slots[self._cur]
        if self._pending is None:                    # first call, or no look-ahead was given
            if ids is not None:
                for d, s in zip(cur, (ids, wts, label)):
                    d.copy_(s, non_blocking=True)
            plan = self.tables.plan_batch(cur[0])
        else:
            if ids is not None and self._pending_src is not None and self._pending_src is not ids:
                raise ValueError("replay(): the batch passed in is not the one given as next_batch to the previous call "
                                 "(its dedup was planned one step ahead); pass the same tensors or next_batch=None")
            plan = self._pending
        self._pending = None
        self._pending_src = None
        if next_batch is not None and _ENV_AHEAD:
            nxt = self._slots[self._cur ^ 1]
            side = self.tables.plan_stream
            if side is not None:
                side.wait_stream(_cu.current_stream())
                with _cu.stream(side):
                    for d, s in zip(nxt, next_batch):
                        d.copy_(s, non_blocking=True)
            else:                                    # peer tables: the key phase is part of the step's graphs
                for d, s in zip(nxt, next_batch):
                    d.copy_(s)
            self._pending = self.tables.plan_batch(nxt[0], ahead=True)
            self._pending_src = next_batch[0]
        out = self._step(plan, cur[1], cur[2])
        if self._pending is not None:
            self._cur ^= 1
        return out


def build_sharded_wide_deep(batch_size, vocab_size, emb_dim, hidden, device, seed=1):
    return ShardedWideDeepStep(batch_size, vocab_size, emb_dim, hidden, device, seed=seed)
