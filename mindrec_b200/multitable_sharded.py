"""BASELINE config 5: the multitable Wide&Deep model's big table row-sharded over the GPUs of one box, plus a
hash-sharded MapParameter for dynamic features, in ONE training step.

    emb128_embedding [rows, 128] + wide_emb128_w [rows, 1]      models/wide_and_deep_multitable/src/wide_and_deep.py:154,189,
                                                                 291-300,366-372 — looked up by the same ids
    HashEmbeddingLookup(128) on a MapParameter                   mindspore_rec/ops/embedding.py:85-206 (int64 keys, admission
                                                                 by permit_filter_value, eviction by evict_filter_value)
    DenseLayers 5 x 1024 + logit, data parallel                  :195-246, train_and_eval_distribute.py:118-119,135-138
    FTRL(lr 0.1, l1 = l2 = 5e-4, accum 0.1) on the "wide" names, Adam(lr 3e-3, eps 1e-6) on the rest, sens 1000 (:525-535)

Semantics stated: the reference gathers with a non-sparse `P.Gather`, which makes every table gradient dense and every
optimizer step touch all rows (mindrec_b200.multitable reproduces exactly that at the reference's 650 000 rows).  At
the 10^8..10^9 rows of this config that is 3.5 KB x rows of HBM traffic per step, so the sharded step takes the
ROW-SPARSE semantics of the reference's own large-table mode (`sparse=True` -> LazyAdam / sparse FTRL,
models/wide_deep/src/wide_and_deep.py:415-430): only looked-up rows and their moments move.  Rows per GPU are fixed
(weak scaling): the table holds rows_per_gpu x G rows.

Exchange: the device-driven protocol of mindrec_b200.peer_sharded for both lookups — owner = key mod G for the table,
owner = hash(key) mod G for the MapParameter; keys, rows and gradients travel as NVLink peer stores into CUDA-IPC
inboxes at offsets computed on the device, and the mean all-reduce of the DenseLayer gradients is a kernel over peer
memory too (peer_sharded.PeerAllReduce); nothing is read back to the host, so the whole step is ONE CUDA graph.
"""
import torch
import torch.distributed as dist

from . import _lib, ops
from .nn import DenseStack
from .peer_sharded import PeerAllReduce, PeerShardedHashEmbedding, PeerShardedTables, _Fork, _join

EMB128 = 128


class ShardedMultitableStep:
    """One training step of config 5.  `__call__(ids, keys, label)` runs it eagerly, `capture` / `replay` as CUDA graphs.

        ids    int32 [B, n_table_fields]   rows of the sharded emb128 table (and of its wide vector)
        keys   int64 [B, n_hash_fields]    dynamic-feature keys in [0, 2^key_bits)
        label  float32 [B, 1]
    """

    def __init__(self, batch_size, rows_total, device, group=None, n_table_fields=26, n_hash_fields=26,
                 deep_dim_list=(1024, 1024, 1024, 1024, 1024), hash_capacity=1 << 22, key_bits=40,
                 permit_filter_value=2, evict_filter_value=8, evict_every=4, seed=1, sens=1000.0,
                 use_mixed_precision=True, adam_lr=3e-3, ftrl_lr=0.1, cap_rows=None, init_std=0.01):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = dev = torch.device(device)
        self.b, self.ft, self.fh = batch_size, n_table_fields, n_hash_fields
        self.sens = float(sens)
        self.mixed = use_mixed_precision
        self.evict_every = int(evict_every)
        g = self.world
        # emb128_embedding + wide_emb128_w, row-sharded; LazyAdam / FTRL with the multitable hyper-parameters
        self.tables = PeerShardedTables(rows_total, EMB128, batch_size * n_table_fields, dev, group=group, seed=seed,
                                        sens=sens, cap_rows=cap_rows, adam=(adam_lr, 1e-6), ftrl=(ftrl_lr, 5e-4, 5e-4, 0.1),
                                        init_std=init_std)
        # the dynamic features: one MapParameter per rank, LazyAdam by slot
        self.hash = PeerShardedHashEmbedding(EMB128, batch_size * n_hash_fields, dev, group=group, key_bits=key_bits,
                                             capacity=hash_capacity, seed=seed, learning_rate=adam_lr, eps=1e-6,
                                             loss_scale=sens * g, permit_filter_value=permit_filter_value,
                                             evict_filter_value=evict_filter_value, cap_rows=cap_rows)
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)                                  # identical DenseLayer replicas on every rank
        self.in_dim = (n_table_fields + n_hash_fields) * EMB128
        dims = [self.in_dim] + list(deep_dim_list) + [1]
        n_dense = DenseStack.numel(dims)
        # the DenseLayer gradients and the wide_bias gradient share one peer-memory buffer: one all-reduce per step,
        # `_reduce` is this rank's contribution, `_reduced` the sum over the ranks
        self._ar = PeerAllReduce(n_dense + 4, dev, group)
        self._reduce, self._reduced = self._ar.src, self._ar.dst
        self.dense = DenseStack(dims, use_mixed_precision, dev, generator=gen, weight_init="normal", bias_init="normal",
                                storage=(torch.zeros(n_dense, dtype=torch.float32, device=dev), self._reduce[:n_dense]))
        self.dense_hyper = ops.adam_hyper(adam_lr, eps=1e-6, loss_scale=sens * g, device=dev)
        self.dense_m, self.dense_v = torch.zeros_like(self.dense.flat), torch.zeros_like(self.dense.flat)
        # wide_bias: a "wide" name, so FTRL (dense kernel on one element); its gradient rides in the all-reduce
        self.wide_bias = torch.empty(1, dtype=torch.float32, device=dev).normal_(0.0, 0.01, generator=gen)
        self.bias_acc, self.bias_lin = torch.full_like(self.wide_bias, 0.1), torch.zeros_like(self.wide_bias)
        self.bias_hyper = ops.ftrl_hyper(ftrl_lr, l1=5e-4, l2=5e-4, loss_scale=sens * g, device=dev)
        f16 = torch.float16 if use_mixed_precision else torch.float32
        b = batch_size
        self._io = dict(
            ids=torch.zeros((b, n_table_fields), dtype=torch.int32, device=dev),
            keys=torch.zeros((b, n_hash_fields), dtype=torch.int64, device=dev),
            label=torch.zeros((b, 1), dtype=torch.float32, device=dev),
            ones=torch.ones((b, n_table_fields), dtype=torch.float32, device=dev),
            x_table=torch.empty((b, n_table_fields * EMB128), dtype=f16, device=dev),
            x_hash=torch.empty((b, n_hash_fields * EMB128), dtype=f16, device=dev),
            wide_out=torch.empty((b, 1), dtype=torch.float32, device=dev),
            loss_out=(torch.empty((b, 1), dtype=torch.float32, device=dev), torch.empty(1, dtype=torch.float32, device=dev),
                      torch.empty((b, 1), dtype=torch.float32, device=dev),
                      torch.empty((b, 1) if use_mixed_precision else (0,), dtype=torch.float16, device=dev),
                      torch.empty(1, dtype=torch.float32, device=dev)))
        self._sens_t = torch.tensor([self.sens], dtype=torch.float32, device=dev)
        self._nan = torch.tensor(float("nan"), dtype=torch.float32, device=dev)
        self._zero = torch.zeros((), dtype=torch.float32, device=dev)
        self._main_stream = torch.cuda.Stream(device=dev, priority=-1)       # see PeerShardedWideDeepStep
        self._dense_stream = torch.cuda.Stream(device=dev)
        self._t_stream = torch.cuda.Stream(device=dev, priority=-1)          # the table's branch beside the MapParameter's
        self._graphs = None
        self._loss = torch.zeros((), dtype=torch.float32, device=dev)
        self._steps = 0
        self._last = None
        self.launches_per_step = None

    # ---- the step (one CUDA graph) -----------------------------------------------------------------------------
    def _forward_exchange(self):
        """plan -> key exchange -> row exchange -> expand.  The table (T) and the MapParameter (H) go through the four
        phases in lock step, always T before H, with every SIGNAL issued on one stream: every rank signals in the same
        order, so a wait can only ever be held up by a peer's earlier work.  What follows a wait of T and is local
        (expand; in the backward the row updates) runs on a forked branch next to H's NVLink phases — those waits depend
        only on signals the peers issue BEFORE their H work.  The dim-1 twins of the table's kernels (wide vector) and
        the owner-side dedup run on another forked branch: they contain no waits."""
        io, t, h = self._io, self.tables.rk, self.hash.rk
        main = torch.cuda.current_stream()
        side = self.tables.owner_stream
        t.p_plan_publish(io["ids"])
        h.p_plan_publish(io["keys"])
        t.wait(0)
        t.p_keys()
        h.wait(0)
        h.p_keys()
        t.wait(1)
        side.wait_stream(main)
        with torch.cuda.stream(side):                        # under serve / expand / DenseLayers
            t.p_owner_dedup()
        t.p_serve(side=side)
        # the table's expand (HBM / L2 bound) runs on a forked branch while the MapParameter's rows are still crossing
        # NVLink: T's wait 2 only depends on the peers' T serves, issued before their H serves
        tb = self._t_stream
        tb.wait_stream(main)
        with torch.cuda.stream(tb):
            t.wait(2)
            t.p_expand(io["ids"].shape, io["ones"], self.wide_bias, io["x_table"], io["wide_out"], side=side)
        h.wait(1)
        h.p_serve()
        h.wait(2)
        h.p_expand(io["x_hash"])
        main.wait_stream(tb)
        main.wait_stream(side)

    def _dense_update(self):
        """All-reduce over peer memory (DistributedGradReducer(mean): the 1/G is in the loss scale) -> dense Adam ->
        FTRL on wide_bias."""
        self._ar.run(self.tables.rk.err)
        self._dense_adam()

    def _dense_adam(self):
        n = self.dense.flat.numel()
        ops.adam_begin_step(self.dense_hyper)
        ops.adam_dense(self.dense.flat, self.dense_m, self.dense_v, self.dense_hyper, self._reduced[:n])
        ops.ftrl_dense(self.wide_bias, self.bias_acc, self.bias_lin, self.bias_hyper, self._reduced[n:n + 1])

    def _body(self):
        io, t, h = self._io, self.tables.rk, self.hash.rk
        main = torch.cuda.current_stream()
        side = self.tables.owner_stream
        self._forward_exchange()
        # DenseLayers: the two lookups' outputs feed layer 0 as column blocks (no concatenation, no split of gx)
        deep_out = self.dense.forward((io["x_table"], io["x_hash"]))
        _, loss, delta, delta16, dsum = ops.sigmoid_xent(io["wide_out"], deep_out, io["label"], self._sens_t,
                                                         out=io["loss_out"])
        n = self.dense.flat.numel()
        self._reduce[n:n + 1].copy_(dsum)                    # d loss / d wide_bias = sum(delta)

        def fork_allreduce():                                # under the input-gradient GEMMs of layer 0
            self._dense_stream.wait_stream(main)
            with torch.cuda.stream(self._dense_stream):
                self._ar.run(t.err)

        g_table, g_hash = self.dense.backward(delta16 if delta16.numel() else delta, on_weight_grads=fork_allreduce)
        self._dense_stream.wait_stream(main)                 # those GEMMs were the last readers of the weights
        with torch.cuda.stream(self._dense_stream):
            self._dense_adam()
            err = t.err[0] | h.err[0]                        # an exchange error poisons the loss the caller reads
            torch.add(loss[0], torch.where(err != 0, self._nan, self._zero), out=self._loss)
        self._last = {"delta": delta, "g_table": g_table, "g_hash": g_hash}      # inspection (eager calls, tests)
        # gradient exchange (same lock-step order as the forward); the table's row updates — HBM bound — run on a forked
        # branch while the MapParameter's gradients are still crossing NVLink
        t.p_grads(delta, g_table, side=side)
        tb = self._t_stream
        tb.wait_stream(main)
        with torch.cuda.stream(tb):
            t.wait(3)
            with _Fork(side):                                # latency-bound FTRL beside the LazyAdam rows
                t.p_update_wide()
            t.p_update_deep()
            _join(side)
        h.p_grads(g_hash)
        h.wait(3)
        h.p_update()
        main.wait_stream(tb)
        main.wait_stream(self._dense_stream)
        return self._loss

    def _one_step(self):
        if self._graphs is not None:
            self._graphs.replay()
        else:
            self._body()
        self._steps += 1
        if self.evict_every and self._steps % self.evict_every == 0:
            self.hash.rk.table.evict()                       # eviction sweep of this rank's MapParameter (eager)
        return self._loss

    def _load(self, ids, keys, label):
        io = self._io
        io["ids"].copy_(ids, non_blocking=True)
        io["keys"].copy_(keys, non_blocking=True)
        io["label"].copy_(label, non_blocking=True)

    def __call__(self, ids, keys, label):
        self._load(ids, keys, label)
        loss = self._one_step()
        return loss, loss

    def capture(self, ids, keys, label, warmup=2, graph=True):
        self._load(ids, keys, label)
        for _ in range(warmup):
            self._one_step()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        if graph:
            gr = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(gr, stream=self._main_stream):
                self._body()
            self.launches_per_step = _lib.launch_count() - n0
            self._graphs = gr
            torch.cuda.synchronize()
            dist.barrier(group=self.group)

    def replay(self, ids, keys, label):
        self._load(ids, keys, label)
        loss = self._one_step()
        return loss, loss

    # ---- inspection ---------------------------------------------------------------------------------
    def error_flags(self):
        """bit 0: a peer wait timed out, bit 1: an inbox overflowed, bit 2: the MapParameter overflowed."""
        return self.tables.error_flags() | self.hash.error_flags()

    def exchange_stats(self):
        """Unique keys of this rank, rows this rank owns in the last step, NVLink bytes leaving this GPU in one step
        (keys out + rows served out + gradients out, for the table (row = 128 + 1 floats) and the MapParameter)."""
        out = {}
        total = 0
        for name, rk, key_b, row_b in (("table", self.tables.rk, 4, (EMB128 + 1) * 4), ("hash", self.hash.rk, 8, EMB128 * 4)):
            bnd = rk.bounds.tolist()
            n_u, own = bnd[self.world], bnd[self.rank + 1] - bnd[self.rank]
            n_r = int(rk.n_r.item())
            nbytes = (n_u - own) * key_b + (n_r - own) * row_b + (n_u - own) * row_b
            out[name] = {"unique_keys": n_u, "rows_owned": n_r, "nvlink_bytes_out": nbytes}
            total += nbytes
        out["nvlink_bytes_out"] = total
        return out

    def close(self):
        self.tables.close()
        self.hash.close()
        self._ar.close()
