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
import torch
import torch.distributed as dist

from . import ops as _cuda_ops


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


class ShardedWideDeepTables:
    """The wide (dim 1) and deep (dim D) tables of Wide&Deep, row-sharded, with their FTRL / LazyAdam state.

    `kernels` is the module providing gather / unique / segment_sum / sparse_* — mindrec_b200.ops (CUDA) in
    production; the gloo CPU tests of the exchange logic inject the oracle."""

    def __init__(self, vocab_size, emb_dim, device, group=None, seed=1, sens=1024.0, kernels=_cuda_ops,
                 init_std=0.01):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.plan = ShardPlan(vocab_size, self.world)
        self.k = kernels
        self.dim = emb_dim
        self.device = torch.device(device)
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
        self.adam_hyper = kernels.adam_hyper(3.5e-4, eps=1e-8, loss_scale=sens * self.world, device=self.device)
        self.ftrl_hyper = kernels.ftrl_hyper(5e-2, l1=1e-8, l2=1e-8, loss_scale=sens * self.world, device=self.device)
        self._edges = None
        self._ctx = None
        self._uq = {}
        self._gs = None

    def _unique(self, key, table_like, tag):
        """mrec_unique into cached output buffers (no per-step allocation)."""
        n = key.numel()
        if not hasattr(self.k, "UniqueResult"):
            return self.k.unique(key, table_like=table_like)
        base = self._uq.get(tag)
        if base is None or base.n < n:          # grow geometrically: the owner-side size varies per step
            base = self.k.UniqueResult(max(n, int(1.5 * base.n) if base else n), key.dtype, key.device)
            self._uq[tag] = base
        return self.k.unique(key, table_like=table_like, result=base if base.n == n else base.sliced(n))

    # ---- forward ----------------------------------------------------------------------------------
    def lookup(self, ids, wts, wide_bias, deep_dtype=torch.float32):
        """ids/wts: [B, F] of this rank's batch shard.  Returns (wide_out [B,1], deep_in [B, F*D])."""
        k, g = self.k, self.world
        b, f = ids.shape
        key = self.plan.remap(ids)
        bound_like = torch.empty((g * self.plan.rows_per_rank, 0), device=ids.device)
        uq = self._unique(key, bound_like, "fwd")
        if self._edges is None:
            self._edges = self.plan.edges(key)
        bounds = k.shard_bounds(uq.uniq, uq.count, self._edges).tolist()      # the step's host read-back
        n_u = bounds[g]
        send = [bounds[r + 1] - bounds[r] for r in range(g)]
        if g > 1:
            t_send = torch.tensor(send, dtype=torch.int64, device=ids.device)
            t_recv = torch.empty_like(t_send)
            dist.all_to_all_single(t_recv, t_send, group=self.group)
            recv = t_recv.tolist()
        else:
            recv = list(send)
        n_r = sum(recv)
        local_rows_send = (uq.uniq[:n_u] % self.plan.rows_per_rank).contiguous()
        rows_recv = torch.empty(n_r, dtype=ids.dtype, device=ids.device)
        if g > 1:
            _a2a(rows_recv, local_rows_send, recv, send, self.group)
        else:
            rows_recv.copy_(local_rows_send)
        # owner side: gather the requested rows of both tables
        deep_rows = k.gather(self.deep, rows_recv)                            # [n_r, D]
        wide_rows = k.gather(self.wide, rows_recv)                            # [n_r, 1]
        got_deep = torch.empty((max(n_u, 1), self.dim), dtype=torch.float32, device=ids.device)
        got_wide = torch.empty((max(n_u, 1), 1), dtype=torch.float32, device=ids.device)
        if g > 1:
            _a2a(got_deep[:n_u], deep_rows, send, recv, self.group)
            _a2a(got_wide[:n_u], wide_rows, send, recv, self.group)
        else:
            got_deep[:n_u].copy_(deep_rows)
            got_wide[:n_u].copy_(wide_rows)
        # requester side: expand unique rows to lookups with the inverse index, fused with mask / reduce
        inverse = uq.inverse.view(b, f)
        deep_in = k.gather_masked(got_deep, inverse, wts, out_dtype=deep_dtype)
        wide_out = k.gather_reduce(got_wide, inverse, wts, wide_bias)
        self._ctx = (uq, n_u, send, recv, rows_recv, wts)
        return wide_out, deep_in

    # ---- backward + update ------------------------------------------------------------------------
    def update(self, delta, gx):
        """delta: [B,1] logit gradient (x sens), gx: [B, F*D] deep-input gradient (x sens, fp32 or fp16)."""
        k, g = self.k, self.world
        uq, n_u, send, recv, rows_recv, wts = self._ctx
        n = uq.n
        mask = wts.reshape(-1)
        if self._gs is None or self._gs[0].shape[0] != n:
            self._gs = (torch.empty((n, self.dim), dtype=torch.float32, device=delta.device),
                        torch.empty((n, 1), dtype=torch.float32, device=delta.device))
        gs_deep = k.segment_sum(gx.view(n, self.dim), mask, uq, dim=self.dim, out=self._gs[0])  # first n_u rows valid
        gs_wide = k.segment_sum(delta, mask, uq, dim=1, out=self._gs[1])
        n_r = sum(recv)
        rg_deep = torch.empty((max(n_r, 1), self.dim), dtype=torch.float32, device=delta.device)
        rg_wide = torch.empty((max(n_r, 1), 1), dtype=torch.float32, device=delta.device)
        if g > 1:
            _a2a(rg_deep[:n_r], gs_deep[:n_u], recv, send, self.group)
            _a2a(rg_wide[:n_r], gs_wide[:n_u], recv, send, self.group)
        else:
            rg_deep[:n_r].copy_(gs_deep[:n_u])
            rg_wide[:n_r].copy_(gs_wide[:n_u])
        if n_r == 0:
            k.adam_begin_step(self.adam_hyper)
            return
        # owner side: the same row can arrive from several ranks -> dedup again, then fused row updates
        uq2 = self._unique(rows_recv, self.deep, "bwd")
        k.sparse_ftrl(self.wide, self.acc, self.lin, self.ftrl_hyper, rg_wide[:n_r], None, uq2)
        k.adam_begin_step(self.adam_hyper)
        k.sparse_lazy_adam(self.deep, self.m, self.v, self.adam_hyper, rg_deep[:n_r], None, uq2)

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


class ShardedWideDeepStep:
    """Wide&Deep training step with row-sharded tables and a data-parallel DenseLayer stack.  Same call
    surface as cells.TrainStepWrap (`__call__`, `capture`, `replay`); the step runs eagerly because the
    all-to-all split sizes are read back every step."""

    def __init__(self, batch_size, vocab_size, emb_dim, hidden, device, seed=1, sens=1024.0, fields=39,
                 use_mixed_precision=True, group=None, kernels=_cuda_ops):
        from .nn import DenseStack
        self.k = kernels
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.sens = float(sens)
        self.fields, self.emb_dim = fields, emb_dim
        self.mixed = use_mixed_precision
        self.tables = ShardedWideDeepTables(vocab_size, emb_dim, device, group=group, seed=seed, sens=sens,
                                            kernels=kernels)
        gen = torch.Generator(device=device)
        gen.manual_seed(seed)                      # identical DenseLayer replicas on every rank
        dims = [fields * emb_dim] + list(hidden) + [1]
        self.dense = DenseStack(dims, use_mixed_precision, device, generator=gen, weight_init="normal",
                                bias_init="normal", extra=1)
        self.wide_b = self.dense.extra
        self.wide_b.normal_(0.0, 0.01, generator=gen)
        self.dense_hyper = kernels.adam_hyper(3.5e-4, eps=1e-8, loss_scale=sens, device=device)
        self.dense_m = torch.zeros_like(self.dense.flat)
        self.dense_v = torch.zeros_like(self.dense.flat)
        self._static = None

    def __call__(self, ids, wts, label):
        k = self.k
        b = ids.shape[0]
        wide_out, deep_in = self.tables.lookup(ids, wts, self.wide_b,
                                               torch.float16 if self.mixed else torch.float32)
        logit = wide_out + self.dense.forward(deep_in)
        log_loss = torch.clamp(logit, min=0) - logit * label + torch.log1p(torch.exp(-logit.abs()))
        loss = log_loss.mean()
        delta = (torch.sigmoid(logit) - label) * (self.sens / b)
        gx = self.dense.backward(delta)
        self.dense.extra_grad.copy_(delta.sum().reshape(1))
        self.tables.update(delta, gx)
        if self.world > 1:                          # DistributedGradReducer(mean): wide_and_deep.py:455-470
            dist.all_reduce(self.dense.flat_grad, group=self.group)
            self.dense.flat_grad.div_(self.world)
        k.adam_begin_step(self.dense_hyper)
        k.adam_dense(self.dense.flat, self.dense_m, self.dense_v, self.dense_hyper, self.dense.flat_grad)
        return loss, loss

    def capture(self, ids, wts, label, warmup=3):
        self._static = (ids.clone(), wts.clone(), label.clone())
        for _ in range(warmup + 1):
            self(*self._static)
        return self._static

    def replay(self, ids=None, wts=None, label=None):
        if ids is not None:
            self._static[0].copy_(ids, non_blocking=True)
            self._static[1].copy_(wts, non_blocking=True)
            self._static[2].copy_(label, non_blocking=True)
        return self(*self._static)


def build_sharded_wide_deep(batch_size, vocab_size, emb_dim, hidden, device, seed=1):
    return ShardedWideDeepStep(batch_size, vocab_size, emb_dim, hidden, device, seed=seed)
