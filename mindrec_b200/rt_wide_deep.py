"""Wide&Deep training step WITHOUT PyTorch: `mindrec_b200.runtime` buffers / streams / graph capture, the aot kernels
through `mindrec_b200.ops`, the DenseLayer GEMMs through the library's own cuBLASLt binding (`runtime.gemm`).

Same step as cells.TrainStepWrap(sparse=True, lazy_adam=True) — models/wide_deep/src/wide_and_deep.py:293-316 (forward),
:349-362 (loss), :376-492 (FTRL on the "wide" table, LazyAdam on the deep table, Adam on the DenseLayers + Wide_b,
loss scale `sens`) — and the same flat DenseLayer parameter layout as nn.DenseStack (per layer W[k,m] then b[m], then
Wide_b), so that state can be exchanged with the torch cells and with the oracle.  BASELINE north_star: "Python host
code calls a .so through the aot C-ABI, with no PyTorch"; `tests/test_rt_wide_deep_gpu.py` runs it in a process that
asserts torch was never imported.

    dev = runtime.Device(0)
    step = WideDeepRT(dev, batch=16000, fields=39, vocab=33_762_616, emb_dim=80, hidden=(1024, 512, 256, 128))
    step.capture()
    loss = step.train_step(ids, wts, label)          # numpy (pinned) or DeviceBuffer inputs; one graph launch
"""
import numpy as np

from . import ops, runtime


class WideDeepRT:
    def __init__(self, dev, batch, fields, vocab, emb_dim, hidden, mixed=True, sens=1024.0, seed=1, init_std=0.01,
                 id_dtype="int32"):
        if emb_dim % 4:
            raise ValueError("emb_dim must be a multiple of 4")
        self.dev, self.b, self.f, self.v, self.d = dev, int(batch), int(fields), int(vocab), int(emb_dim)
        self.mixed, self.sens = bool(mixed), float(sens)
        self.dims = [self.f * self.d] + list(hidden) + [1]
        if len(self.dims) < 3:
            raise ValueError("at least one hidden DenseLayer")
        nl = len(self.dims) - 1
        act = "float16" if mixed else "float32"
        b, n = self.b, self.b * self.f
        # ---- tables + optimizer state (wide_and_deep.py:420-430: LazyAdam lr 3.5e-4 eps 1e-8; FTRL lr 5e-2, l1 = l2 = 1e-8,
        # initial_accum 1.0) -------------------------------------------------------------------------------------------
        self.wide, self.deep = dev.empty((self.v, 1)), dev.empty((self.v, self.d))
        self._init_table(self.wide, init_std, seed * 7 + 1)
        self._init_table(self.deep, init_std, seed * 7 + 2)
        self.acc, self.lin = dev.full((self.v, 1), 1.0), dev.zeros((self.v, 1))
        self.m, self.vv = dev.zeros((self.v, self.d)), dev.zeros((self.v, self.d))
        self.adam_deep = ops.adam_hyper(3.5e-4, eps=1e-8, loss_scale=self.sens, device=dev)
        self.ftrl_wide = ops.ftrl_hyper(5e-2, l1=1e-8, l2=1e-8, loss_scale=self.sens, device=dev)
        # ---- DenseLayers: one flat fp32 parameter / gradient / moment buffer (+ Wide_b: the Adam group quirk) --
        n_flat = sum(self.dims[i] * self.dims[i + 1] + self.dims[i + 1] for i in range(nl)) + 1
        rng = np.random.default_rng(seed)
        self.flat = dev.from_numpy(rng.normal(0.0, 0.01, n_flat).astype(np.float32))
        self.flat_grad, self.flat_m, self.flat_v = dev.zeros(n_flat), dev.zeros(n_flat), dev.zeros(n_flat)
        self.flat16 = dev.empty(n_flat, "float16") if mixed else None
        self.adam_dense = ops.adam_hyper(3.5e-4, eps=1e-8, loss_scale=self.sens, device=dev)
        self.w, self.bias, self.gw, self.gb = [], [], [], []
        src = self.flat16 if mixed else self.flat
        o = 0
        for i in range(nl):
            k, m_ = self.dims[i], self.dims[i + 1]
            self.w.append(src[o:o + k * m_].view(k, m_))
            self.gw.append(self.flat_grad[o:o + k * m_].view(k, m_))
            o += k * m_
            self.bias.append(src[o:o + m_])
            self.gb.append(self.flat_grad[o:o + m_])
            o += m_
        self.wide_b, self.wide_b_grad = self.flat[o:o + 1], self.flat_grad[o:o + 1]
        # ---- static inputs, activations, gradients: nothing is allocated inside a step (graph capture) ------------------
        # two input slots (one captured graph each): while a step runs on one, the next batch is copied into the other
        self.slots = [(dev.zeros((b, self.f), id_dtype), dev.zeros((b, self.f)), dev.zeros((b, 1))) for _ in range(2)]
        self.cur = 0
        self.acts = [dev.empty((b, self.dims[0]), act)] + [dev.empty((b, self.dims[i]), act) for i in range(1, nl)]
        self.deep_out, self.wide_out = dev.empty((b, 1)), dev.empty((b, 1))
        self.loss_out = (dev.empty((b, 1)), dev.zeros(1), dev.empty((b, 1)), dev.empty((b, 1) if mixed else (0,), "float16"),
                         dev.empty(1))
        self.g = [dev.empty((b, self.dims[i]), act) for i in range(nl)]      # gradient wrt the input of layer i
        self.sens_t = dev.tensor([self.sens])
        self.uq = ops.UniqueResult(n, id_dtype, dev)
        self.side, self.side2, self.copy_stream = runtime.Stream(), runtime.Stream(), runtime.Stream()
        self.graphs = None
        self._staged = None
        self.launches_per_step = None

    ids = property(lambda self: self.slots[self.cur][0])
    wts = property(lambda self: self.slots[self.cur][1])
    label = property(lambda self: self.slots[self.cur][2])

    def _init_table(self, t, std, seed):
        """normal(0, std) rows: one random block of 2^16 rows tiled over the table (synthetic weights; tests load theirs)."""
        rows = t.shape[0]
        blk = min(rows, 1 << 16)
        block = self.dev.from_numpy(np.random.default_rng(seed).normal(0.0, std, (blk,) + t.shape[1:]).astype(np.float32))
        for r0 in range(0, rows, blk):
            r1 = min(rows, r0 + blk)
            t[r0:r1].copy_(block[:r1 - r0])

    def load_state(self, wide=None, deep=None, flat=None):
        """Overwrite tables / the flat DenseLayer buffer (numpy arrays in nn.DenseStack's layout) — tests, checkpoints."""
        if wide is not None:
            self.wide.copy_(np.asarray(wide, np.float32).reshape(self.v, 1))
        if deep is not None:
            self.deep.copy_(np.asarray(deep, np.float32).reshape(self.v, self.d))
        if flat is not None:
            self.flat.copy_(np.asarray(flat, np.float32).reshape(-1))

    # ---- one step on the static inputs ---------------------------------------------------------------------------
    def _body(self):
        dev, nl, mixed = self.dev, len(self.dims) - 1, self.mixed
        main = runtime.Stream(dev.current_stream_handle())
        n = self.b * self.f
        # forked: the dedup of the ids (needs nothing but the ids; under the GEMMs) and the wide forward
        self.side.wait_stream(main)
        with dev.use_stream(self.side):
            ops.unique(self.ids, table_like=self.deep, result=self.uq)
        self.side2.wait_stream(main)
        with dev.use_stream(self.side2):
            ops.gather_reduce(self.wide, self.ids, self.wts, self.wide_b, out=self.wide_out)
        if mixed:
            ops.cast_f32_f16(self.flat, out=self.flat16)     # Cast(weight, float16): wide_and_deep.py:119-122
        ops.gather_masked(self.deep, self.ids, self.wts, out=self.acts[0])
        for i in range(nl - 1):                              # MatMul + BiasAdd + ReLU
            runtime.gemm(self.acts[i], self.w[i], self.acts[i + 1], bias=self.bias[i], relu=True)
        ops.dense_head_fwd(self.acts[nl - 1], self.w[nl - 1].view(-1), self.bias[nl - 1], out=self.deep_out)
        main.wait_stream(self.side2)
        _, loss, delta, delta16, dsum = ops.sigmoid_xent(self.wide_out, self.deep_out, self.label, self.sens_t,
                                                         out=self.loss_out)
        # ---- backward -------------------------------------------------------------------------------------------
        self.wide_b_grad.copy_(dsum)
        g = ops.dense_head_bwd(delta16 if mixed else delta, self.acts[nl - 1], self.w[nl - 1].view(-1), True,
                               self.gw[nl - 1].view(-1), self.gb[nl - 1], self.gb[nl - 2], out=self.g[nl - 1])
        for i in range(nl - 2, -1, -1):
            if i != nl - 2:                                  # (the head kernel did layer nl-2's ReluGrad + BiasAddGrad)
                g = ops.relu_bwd_bias(g, self.acts[i + 1], self.gb[i])
            runtime.gemm(self.acts[i], g, self.gw[i], trans_a=True)                      # weight gradient, fp32 out
            if i == 0 and mixed:
                # every DenseLayer gradient has been issued, and the last GEMM reads the fp16 shadow of the weights: the
                # dense Adam rewrites the fp32 masters on a forked branch underneath it
                self.side2.wait_stream(main)
                with dev.use_stream(self.side2):
                    ops.adam_begin_step(self.adam_dense)
                    ops.adam_dense(self.flat, self.flat_m, self.flat_v, self.adam_dense, self.flat_grad)
            g = runtime.gemm(g, self.w[i], self.g[i], trans_b=True)                      # input gradient
        gx = g.view(n, self.d)
        mask = self.wts.view(-1)
        # ---- updates: FTRL on the wide rows beside LazyAdam on the deep rows, then Adam on the DenseLayers -----------
        main.wait_stream(self.side)                          # dedup done
        self.side.wait_stream(main)                          # gradients ready
        with dev.use_stream(self.side):
            ops.sparse_ftrl(self.wide, self.acc, self.lin, self.ftrl_wide, delta, mask, self.uq)
        ops.adam_begin_step(self.adam_deep)
        ops.sparse_lazy_adam(self.deep, self.m, self.vv, self.adam_deep, gx, mask, self.uq)
        if mixed:
            main.wait_stream(self.side2)
        else:
            ops.adam_begin_step(self.adam_dense)
            ops.adam_dense(self.flat, self.flat_m, self.flat_v, self.adam_dense, self.flat_grad)
        main.wait_stream(self.side)
        return loss

    def capture(self, warmup=1):
        """Warm up (library handles, workspaces, cuBLASLt heuristics) on the current static inputs, then record the step
        into one CUDA graph.  The warm-up steps TRAIN (like cells.TrainStepWrap.capture)."""
        from . import _lib
        for _ in range(max(1, warmup)):
            self._body()
        self.dev.synchronize()
        keep, graphs = self.cur, []
        for slot in (0, 1):                                  # the same step, reading the other input slot
            self.cur = slot
            n0 = _lib.launch_count()
            with self.dev.capture() as g:
                self._body()
            self.launches_per_step = _lib.launch_count() - n0
            graphs.append(g)
        self.cur, self.graphs = keep, graphs
        return self

    def set_inputs(self, ids, wts, label):
        for dst, src in zip(self.slots[self.cur], (ids, wts, label)):
            dst.copy_(src)

    def train_step(self, ids=None, wts=None, label=None, next_batch=None):
        """One step on (ids, wts, label): numpy arrays (pin them and set dev.async_host_copies for overlap) or
        DeviceBuffers.  next_batch: the batch the caller will pass next — it is copied into the other input slot on a copy
        stream while this step runs; passing the same objects to the next call then costs no copy.  Returns the loss as
        a DeviceBuffer [1] (`.item()` reads it)."""
        dev = self.dev
        main = runtime.Stream(dev.current_stream_handle())
        if ids is not None:
            if self._staged is not None and self._staged[0] is ids:
                self.cur = self._staged[2]
                main.wait_event(self._staged[1])
            else:
                self.set_inputs(ids, wts, label)
        self._staged = None
        if next_batch is not None:
            other = self.cur ^ 1
            self.copy_stream.wait_stream(main)               # the last step that read that slot has been enqueued
            with dev.use_stream(self.copy_stream):
                for dst, src in zip(self.slots[other], next_batch):
                    dst.copy_(src)
            ev = runtime.Event()
            ev.record(self.copy_stream)
            self._staged = (next_batch[0], ev, other)
        if self.graphs is not None:
            self.graphs[self.cur].launch()
        else:
            self._body()
        return self.loss_out[1]
