"""Model cells of the reference, restated over the aot kernels.

    WideDeepModel / NetWithLossClass / TrainStepWrap / PredictWithSigmoid
        models/wide_deep/src/wide_and_deep.py:136-519
The embedding lookup, mask multiply, wide reduce, sparse gradient dedup and the optimizer updates run in
libmindrec_b200.so; the DenseLayer stack uses library GEMMs (torch.mm -> cuBLAS), as SURVEY 2b scopes it.
Because there is no graph compiler here, TrainStepWrap states forward, backward and update explicitly; a
step touches only pre-allocated buffers and can be captured in a CUDA graph (`TrainStepWrap.capture`).
"""
import contextlib

import torch

from . import ops
from .nn import Adam, DenseStack, EmbeddingLookup, FTRL, LazyAdam, Parameter, RowTensor


class StepProfile:
    """Optional per-phase CUDA-event timing of an eager step (bench.py's breakdown / roofline source)."""

    def __init__(self):
        self.records = []

    @contextlib.contextmanager
    def range(self, name):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        yield
        b.record()
        self.records.append((name, a, b))

    def totals(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self.records:
            out.setdefault(name, []).append(a.elapsed_time(b))
        return out


@contextlib.contextmanager
def _null_range(name):
    yield


class WideDeepConfig:
    """Hyper-parameters of models/wide_deep/default_config.yaml:16-44 (same names)."""

    def __init__(self, batch_size=16000, field_size=39, vocab_size=200000, emb_dim=80,
                 deep_layer_dim=(1024, 512, 256, 128), deep_layer_act="relu", keep_prob=1.0,
                 dropout_flag=False, l2_coef=8e-5, emb_init="normal", weight_bias_init=("normal", "normal"),
                 use_mixed_precision=True, sparse=False, dynamic_embedding=False, parameter_server=False,
                 vocab_cache_size=0, seed=1, hash_capacity=1 << 20, hash_auto_grow=True, interleave_adam_state=False):
        self.batch_size = batch_size
        self.field_size = field_size
        self.vocab_size = vocab_size
        self.emb_dim = emb_dim
        self.deep_layer_dim = list(deep_layer_dim)
        self.deep_layer_act = deep_layer_act
        self.keep_prob = keep_prob
        self.dropout_flag = dropout_flag
        self.l2_coef = l2_coef
        self.emb_init = emb_init
        self.weight_bias_init = tuple(weight_bias_init)
        self.use_mixed_precision = use_mixed_precision
        self.sparse = sparse
        self.dynamic_embedding = dynamic_embedding
        self.parameter_server = parameter_server
        self.vocab_cache_size = vocab_cache_size
        self.seed = seed
        # dynamic_embedding only: initial slot count of the two MapParameters and whether they grow by themselves
        self.hash_capacity = hash_capacity
        self.hash_auto_grow = hash_auto_grow
        # deep table stored as one [V,3,D] array of w | m | v records (LazyAdam only): see nn.EmbeddingLookup
        self.interleave_adam_state = interleave_adam_state


class WideDeepModel:
    """wide_and_deep.py:136-316.  construct(ids, wts) -> (logit[B,1], embedding_table)."""

    def __init__(self, config, device="cuda"):
        if config.deep_layer_act != "relu":
            raise ValueError("only the reference default deep_layer_act='relu' is implemented")
        if config.dropout_flag:
            raise ValueError("dropout_flag=True is not on the benchmarked path (keep_prob 1.0)")
        self.config = config
        self.batch_size = config.batch_size
        self.field_size = config.field_size
        self.emb_dim = config.emb_dim
        self.device = torch.device(device)
        self.dynamic = bool(config.dynamic_embedding)
        self.cached = (not self.dynamic) and int(config.vocab_cache_size) > 0
        gen = torch.Generator(device=self.device)
        gen.manual_seed(config.seed)
        if self.cached:
            # embedding-cache mode (wide_and_deep.py:176-230 with vocab_cache_size > 0): the tables live in pinned host
            # memory, the device holds vocab_cache_size rows behind an id -> slot map (mindrec_b200.cache); from here
            # on the model addresses rows by slot exactly as in dynamic-embedding mode
            from .cache import CachedEmbeddingLookup
            kw = dict(param_init=config.emb_init, sparse=config.sparse, device=self.device, generator=gen)
            self.wide_embeddinglookup = CachedEmbeddingLookup(config.vocab_size, 1, config.vocab_cache_size,
                                                              name="wide_embeddinglookup.embedding_table", **kw)
            self.deep_embeddinglookup = CachedEmbeddingLookup(config.vocab_size, config.emb_dim, config.vocab_cache_size,
                                                              name="deep_embeddinglookup.embedding_table", **kw)
            self.dynamic = True
        elif self.dynamic:
            # wide_and_deep.py:268-274: HashEmbeddingLookup(embedding_size=emb_dim) + HashEmbeddingLookup(embedding_size=1)
            # over two MapParameters (int32 keys, admitted on first sight, never evicted: the constructor defaults)
            from .hash import HashEmbeddingLookup
            kw = dict(param_init=config.emb_init, capacity=config.hash_capacity, device=self.device,
                      auto_grow=config.hash_auto_grow)
            self.deep_embeddinglookup = HashEmbeddingLookup(config.emb_dim, seed=config.seed, **kw)
            self.wide_embeddinglookup = HashEmbeddingLookup(1, seed=config.seed + 1, **kw)
            self.deep_embeddinglookup.embedding_table.name = "deep_embeddinglookup.embedding_table"
            self.wide_embeddinglookup.embedding_table.name = "wide_embeddinglookup.embedding_table"
        else:
            self.wide_embeddinglookup = EmbeddingLookup(config.vocab_size, 1, param_init=config.emb_init,
                                                        sparse=config.sparse, device=self.device,
                                                        name="wide_embeddinglookup.embedding_table", generator=gen)
            self.deep_embeddinglookup = EmbeddingLookup(config.vocab_size, config.emb_dim,
                                                        param_init=config.emb_init, sparse=config.sparse,
                                                        device=self.device,
                                                        name="deep_embeddinglookup.embedding_table", generator=gen,
                                                        interleave_state=config.interleave_adam_state)
        self.embedding_table = self.deep_embeddinglookup.embedding_table
        dims = [self.field_size * self.emb_dim] + list(config.deep_layer_dim) + [1]
        w_init, b_init = config.weight_bias_init
        self.dense = DenseStack(dims, config.use_mixed_precision, self.device, generator=gen,
                                weight_init=w_init, bias_init=b_init, extra=1)
        # "Wide_b" (capital W): lives in the Adam group, wide_and_deep.py:161-163,405-413
        self.wide_b = Parameter(self.dense.extra, name="Wide_b")
        if config.emb_init == "normal":
            self.wide_b.data.normal_(0.0, 0.01, generator=gen)
        self.slots_w = self.slots_d = None          # dynamic mode: the step's slot indices (the RowTensor indices)
        self._wide_out = None
        self._deep_in = None
        self._wide_stream = None
        self.defer_add = False       # set by NetWithLossClass: `out = wide_out + deep_out` happens in mrec_sigmoid_xent

    def trainable_params(self):
        return [self.wide_embeddinglookup.embedding_table, self.deep_embeddinglookup.embedding_table,
                Parameter(self.dense.flat, name="dense_layers+Wide_b")]

    def __call__(self, id_hldr, wt_hldr):
        return self.construct(id_hldr, wt_hldr)

    def construct(self, id_hldr, wt_hldr):
        b, f, d = id_hldr.shape[0], self.field_size, self.emb_dim
        if self._wide_out is None or self._wide_out.shape[0] != b:
            self._wide_out = torch.empty((b, 1), dtype=torch.float32, device=self.device)
            # fp16 when the DenseLayers run in mixed precision: the Cast is fused into the gather store
            self._deep_in = torch.empty((b, f * d), device=self.device,
                                        dtype=torch.float16 if self.config.use_mixed_precision else torch.float32)
        # wide_and_deep.py:300,303,305-306: gather(dim 1) * mask, ReduceSum(axis 1) + Wide_b.  The wide term is
        # needed only by the loss: it runs on its own stream (a parallel graph branch) beside the deep path.
        main = torch.cuda.current_stream()
        if self._wide_stream is None:
            self._wide_stream = torch.cuda.Stream(device=self.device)
        if self.dynamic:
            # HashEmbeddingLookup.construct (embedding.py:184-206): key -> slot (find-or-insert; new rows initialised),
            # then the same fused gathers read the arenas by slot (slot C = the default row of unadmitted keys)
            wt, dt = self.wide_embeddinglookup, self.deep_embeddinglookup
            for emb in (wt, dt):
                if emb.auto_grow:
                    emb.embedding_table.maybe_grow(id_hldr.numel())
        self._wide_stream.wait_stream(main)
        with torch.cuda.stream(self._wide_stream):
            if self.dynamic:
                self.slots_w = wt.embedding_table.lookup_slots(id_hldr).view(b, f)
                ops.gather_reduce(wt.embedding_table.values, self.slots_w, wt_hldr, self.wide_b.data, out=self._wide_out)
            else:
                ops.gather_reduce(self.wide_embeddinglookup.embedding_table.data, id_hldr, wt_hldr,
                                  self.wide_b.data, out=self._wide_out)
        # wide_and_deep.py:302,308-309: gather(dim D) * mask -> [B, F*D]
        if self.dynamic:
            self.slots_d = dt.embedding_table.lookup_slots(id_hldr).view(b, f)
            ops.gather_masked(dt.embedding_table.values, self.slots_d, wt_hldr, out=self._deep_in)
        else:
            ops.gather_masked(self.deep_embeddinglookup.embedding_table.kernel_arg, id_hldr, wt_hldr,
                              out=self._deep_in)
        deep_out = self.dense.forward(self._deep_in)       # :310-314
        main.wait_stream(self._wide_stream)
        self.wide_out, self.deep_out = self._wide_out, deep_out
        if self.defer_add:                                  # :315 is fused into the loss kernel
            return None, self.embedding_table
        return self._wide_out + deep_out, self.embedding_table


class NetWithLossClass:
    """wide_and_deep.py:319-362: (wide_loss, deep_loss); the l2 term exists only in dense mode."""

    def __init__(self, network, config):
        self.network = network
        self.no_l2loss = bool(config.parameter_server) or bool(config.sparse)
        self.l2_coef = config.l2_coef
        self.logit = None
        self._out = None
        self.sens_t = torch.ones(1, dtype=torch.float32, device=network.device)   # TrainStepWrap sets sens

    def __call__(self, batch_ids, batch_wts, label):
        return self.construct(batch_ids, batch_wts, label)

    def construct(self, batch_ids, batch_wts, label):
        net = self.network
        net.defer_add = True
        _, embedding_table = net(batch_ids, batch_wts)
        net.defer_add = False
        b = batch_ids.shape[0]
        if self._out is None or self._out[0].shape[0] != b:
            dev = batch_ids.device
            half = net.config.use_mixed_precision
            self._out = (torch.empty((b, 1), dtype=torch.float32, device=dev), torch.empty(1, dtype=torch.float32, device=dev),
                         torch.empty((b, 1), dtype=torch.float32, device=dev),
                         torch.empty((b, 1) if half else (0,), dtype=torch.float16, device=dev),
                         torch.empty(1, dtype=torch.float32, device=dev))
        # logit = wide + deep (:315), mean SigmoidCrossEntropyWithLogits (:354-355) and the sens-scaled seed
        # of the two backward passes (:479-486), in one kernel
        self.logit, loss, self.delta, self.delta16, self.delta_sum = ops.sigmoid_xent(
            net.wide_out, net.deep_out, label, self.sens_t, out=self._out)
        wide_loss = loss[0]
        if self.no_l2loss:
            deep_loss = wide_loss
        else:
            tbl = embedding_table.values[:embedding_table.capacity] if net.dynamic else embedding_table.data
            l2_loss_v = tbl.square().sum() / 2
            deep_loss = wide_loss + self.l2_coef * l2_loss_v
        return wide_loss, deep_loss


class TrainStepWrap:
    """wide_and_deep.py:376-492: FTRL on the "wide" group, Adam / LazyAdam on the rest, loss scale `sens`.

    LazyAdam is chosen exactly when the reference chooses it (`:415-419`): sparse with auto-parallel or
    parameter server, or dynamic embedding; pass lazy_adam=True to request the row-sparse semantics on a
    single device (what BASELINE config 2 measures — nn.Adam there would stream the whole 34 M-row table)."""

    def __init__(self, network, sens=1024.0, parameter_server=False, sparse=False, cache_enable=False,
                 dynamic_embedding=False, is_auto_parallel=False, lazy_adam=None):
        self.network = network
        model = network.network
        self.model = model
        self.sens = float(sens)
        network.sens_t.fill_(self.sens)
        self.sparse = sparse
        self.dynamic = bool(getattr(model, "dynamic", False))
        if self.dynamic and not (dynamic_embedding or (cache_enable and getattr(model, "cached", False))):
            raise ValueError("the model was built with dynamic_embedding=True (or vocab_cache_size > 0): pass "
                             "dynamic_embedding=True (cache_enable=True) here too")
        if getattr(model, "cached", False):
            dynamic_embedding = True            # rows are addressed by cache slot: the MapParameter optimizer path
        if self.dynamic and not network.no_l2loss:
            # the reference's launcher pairs --dynamic_embedding=True with --sparse=True
            # (scripts/run_dynamic_embed_standalone_train_for_gpu.sh:24-30): no dense l2 gradient on a MapParameter
            raise ValueError("dynamic_embedding needs sparse=True (no full-table l2 term on a hash table)")
        if lazy_adam is None:
            lazy_adam = (sparse and is_auto_parallel) or (sparse and parameter_server) or dynamic_embedding
        self.lazy_adam = bool(lazy_adam)
        # weights_w: names containing "wide" (lower case) -> only the wide table
        self.weights_w = [model.wide_embeddinglookup.embedding_table]
        self.weights_d = [model.deep_embeddinglookup.embedding_table,
                          Parameter(model.dense.flat, name="dense_layers+Wide_b")]
        opt_cls = LazyAdam if self.lazy_adam else Adam
        self.optimizer_d = opt_cls(self.weights_d, learning_rate=3.5e-4, eps=1e-8, loss_scale=sens)
        self.optimizer_w = FTRL(learning_rate=5e-2, params=self.weights_w, l1=1e-8, l2=1e-8,
                                initial_accum=1.0, loss_scale=sens)
        if not network.no_l2loss:
            # dense mode: d(deep_loss)/dWd carries l2_coef * Wd on every row (wide_and_deep.py:359-360)
            self.optimizer_d.hyper[8] = network.l2_coef
        self._uq = None
        self._uq_w = None
        self._graph = None
        self._static = None
        self.profile = None
        self._staged = None
        self.overlap = True          # fork dedup / FTRL onto a side stream (see construct)
        self._side = None

    def __call__(self, batch_ids, batch_wts, label):
        return self.construct(batch_ids, batch_wts, label)

    def construct(self, batch_ids, batch_wts, label):
        """One training step.  Two pieces of work are forked onto a side stream (parallel branches of the
        captured graph): the dedup of the ids, which depends on nothing but the ids and hides under the
        DenseLayer GEMMs, and the latency-bound FTRL update of the wide table, which runs beside the
        bandwidth-bound LazyAdam update of the deep table."""
        model = self.model
        rng = self.profile.range if self.profile is not None else _null_range
        b = batch_ids.shape[0]
        n = batch_ids.numel()
        if self.dynamic:
            return self._construct_dynamic(batch_ids, batch_wts, label)
        main = torch.cuda.current_stream()
        side = self._side_stream() if self.overlap else main
        if self._uq is None or self._uq.n != n:
            self._uq = ops.UniqueResult(n, batch_ids.dtype, batch_ids.device)
        side.wait_stream(main)
        with torch.cuda.stream(side), rng("unique"):
            uq = ops.unique(batch_ids, table_like=model.embedding_table.kernel_arg, result=self._uq)
        with rng("forward"):
            loss_w, loss_d = self.network(batch_ids, batch_wts, label)
        # (mixed precision only: the input-gradient GEMM then reads the fp16 shadow of the weights, so the dense Adam may
        # rewrite the fp32 masters underneath it)
        fork_dense = self.overlap and self.lazy_adam and bool(model.dense.convert_dtype)

        def dense_update():
            # every DenseLayer gradient has been issued: the dense Adam runs on a forked branch underneath the
            # input-gradient GEMM of layer 0 instead of after the sparse update
            side2 = self._side_stream(1)
            side2.wait_stream(main)
            with torch.cuda.stream(side2):
                self.optimizer_d.begin_step()
                self.optimizer_d.apply(1, model.dense.flat_grad)

        with rng("dense_backward"):
            delta = self.network.delta                                            # sens*(sigmoid-y)/B, [B,1]
            seed = self.network.delta16 if self.network.delta16.numel() else delta
            model.dense.extra_grad.copy_(self.network.delta_sum)                  # Wide_b gradient
            gx = model.dense.backward(seed, on_weight_grads=dense_update if fork_dense else None)    # [B, F*D]
        mask = batch_wts.reshape(-1)
        grads_w = [RowTensor(batch_ids, delta, mask, uq)]
        grads_d = [RowTensor(batch_ids, gx.view(n, model.emb_dim), mask, uq), model.dense.flat_grad]
        main.wait_stream(side)                 # dedup done
        side.wait_stream(main)                 # gradients ready
        with torch.cuda.stream(side), rng("ftrl_wide"):
            self.optimizer_w(grads_w)
        with rng("adam_deep"):
            if fork_dense:
                main.wait_stream(self._side_stream(1))                            # begin_step (and the dense Adam) done
                self.optimizer_d.apply(0, grads_d[0])
            else:
                self.optimizer_d(grads_d)
        main.wait_stream(side)
        return loss_w, loss_d

    def _construct_dynamic(self, batch_ids, batch_wts, label):
        """dynamic_embedding=True (wide_and_deep.py:268-274,415-430): the two tables are MapParameters, the sparse
        gradients are addressed by the slots the forward lookups returned (each table has its own key -> slot map, so
        each gets its own dedup), LazyAdam / FTRL update rows and sibling-arena state by slot."""
        model = self.model
        rng = self.profile.range if self.profile is not None else _null_range
        n = batch_ids.numel()
        main = torch.cuda.current_stream()
        side = self._side_stream() if self.overlap else main
        with rng("forward"):
            loss_w, loss_d = self.network(batch_ids, batch_wts, label)
        if self._uq is None or self._uq.n != n:
            self._uq = ops.UniqueResult(n, torch.int32, batch_ids.device)
            self._uq_w = ops.UniqueResult(n, torch.int32, batch_ids.device)
        tw, td = model.wide_embeddinglookup.embedding_table, model.deep_embeddinglookup.embedding_table
        side.wait_stream(main)
        with torch.cuda.stream(side), rng("unique"):        # under the DenseLayer backward
            uq_w = ops.unique(model.slots_w, table_like=tw.values[:tw.capacity], result=self._uq_w, ws_tag="unique_wide")
            uq_d = ops.unique(model.slots_d, table_like=td.values[:td.capacity], result=self._uq)
        with rng("dense_backward"):
            delta = self.network.delta
            seed = self.network.delta16 if self.network.delta16.numel() else delta
            gx = model.dense.backward(seed)
            model.dense.extra_grad.copy_(self.network.delta_sum)
        mask = batch_wts.reshape(-1)
        main.wait_stream(side)
        side.wait_stream(main)
        with torch.cuda.stream(side), rng("ftrl_wide"):
            self.optimizer_w([RowTensor(model.slots_w, delta, mask, uq_w)])
        with rng("adam_deep"):
            self.optimizer_d([RowTensor(model.slots_d, gx.view(n, model.emb_dim), mask, uq_d), model.dense.flat_grad])
        main.wait_stream(side)
        return loss_w, loss_d

    def _side_stream(self, i=0):
        if self._side is None:
            self._side = [torch.cuda.Stream(device=self.model.device) for _ in range(2)]
        return self._side[i]

    # ---- CUDA-graph replay of the whole step -------------------------------------------------------
    def capture(self, batch_ids, batch_wts, label, warmup=3):
        """Capture construct() on static input buffers.  Afterwards `replay(ids, wts, label)` launches the captured
        step.  The static inputs exist TWICE (one captured graph each, sharing one memory pool): while a step runs on
        one slot, the next batch is copied into the other on a copy stream, so no input copy sits on the step's critical
        path."""
        if self.dynamic and (self.model.wide_embeddinglookup.auto_grow or self.model.deep_embeddinglookup.auto_grow):
            raise RuntimeError("dynamic_embedding with hash_auto_grow=True cannot be captured (growth re-allocates the "
                               "tables): size hash_capacity up front and set hash_auto_grow=False")
        self._slots = [(batch_ids.clone(), batch_wts.clone(), label.clone()) for _ in range(2)]
        self._static = self._slots[0]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.construct(*self._static)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._graphs, self._outs, pool = [], [], None
        for slot in self._slots:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                out = self.construct(*slot)
            pool = pool or g.pool()
            self._graphs.append(g)
            self._outs.append(out)
        self._graph, self._static_out = self._graphs[0], self._outs[0]
        self._cur = 0
        self._staged = None
        self._copy_stream = torch.cuda.Stream(device=batch_ids.device)
        return self._static

    def replay(self, batch_ids=None, batch_wts=None, label=None, next_batch=None):
        """Launch the captured step on (batch_ids, batch_wts, label).

        next_batch (optional, pinned-host or device tensors): the batch the caller will pass next.  It is copied into
        the OTHER input slot on a copy stream while this step computes (5 MB over PCIe for 16000 x 39 from the host);
        passing the same tensors to the next call then costs no copy at all.  A batch that was not announced is copied
        into the current slot on the compute stream."""
        main = torch.cuda.current_stream()
        slot = self._cur
        if batch_ids is not None:
            if self._staged is not None and self._staged[0] is batch_ids:
                slot = self._staged[2]
                main.wait_event(self._staged[1])
            else:
                for d, s_ in zip(self._slots[slot], (batch_ids, batch_wts, label)):
                    d.copy_(s_, non_blocking=True)
        self._staged = None
        self._cur = slot
        if next_batch is not None:
            other = slot ^ 1
            self._copy_stream.wait_stream(main)       # the last step that read this slot has been enqueued
            with torch.cuda.stream(self._copy_stream):
                for d, s_ in zip(self._slots[other], next_batch):
                    d.copy_(s_, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            self._staged = (next_batch[0], ev, other)
        self._graphs[slot].replay()
        return self._outs[slot]


class PredictWithSigmoid:
    """wide_and_deep.py:495-519."""

    def __init__(self, network):
        self.network = network

    def __call__(self, batch_ids, batch_wts, labels):
        logits, _ = self.network(batch_ids, batch_wts)
        return logits, torch.sigmoid(logits), labels
