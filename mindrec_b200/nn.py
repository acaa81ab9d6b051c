"""Host-side mirror of the reference's operator surface for the hot path.

Names, constructor arguments and error behaviour follow the reference / the MindSpore cells it calls:

    EmbeddingLookup      mindspore.nn.EmbeddingLookup as used at models/wide_deep/src/wide_and_deep.py:234-290
    Adam / LazyAdam      mindspore.nn.optim (wide_and_deep.py:420-422,435-437; deepfm.py:272)
    FTRL                 mindspore.nn.optim.FTRL (wide_and_deep.py:423-430)
    RowTensor            the sparse gradient of SparseGatherV2

All arithmetic runs in libmindrec_b200.so through the aot C-ABI (mindrec_b200.ops); torch only owns the
device buffers.
"""
import math

import torch

from . import ops


class Parameter:
    """A named device buffer (mindspore.Parameter stand-in)."""

    def __init__(self, data, name, requires_grad=True, packed=None):
        self.data = data
        self.name = name
        self.requires_grad = requires_grad
        # interleaved layout (EmbeddingLookup(interleave_state=True)): `packed` is wmv[V,3,D], each row's weights |
        # Adam m | Adam v back to back; `data` is then the strided view packed[:, 0, :]
        self.packed = packed

    @property
    def kernel_arg(self):
        """What the aot kernels are handed: the contiguous table, or the interleaved record array."""
        return self.packed if self.packed is not None else self.data

    @property
    def shape(self):
        return tuple(self.data.shape)

    def __repr__(self):
        return "Parameter(name=%s, shape=%s)" % (self.name, self.shape)


class RowTensor:
    """Sparse gradient of a gather: lookup position n contributes mask[n] * values[n // div] to row
    indices[n].  `uq` caches the dedup (mrec_unique) so that several tables looked up with the same ids
    (wide + deep) sort them once."""

    def __init__(self, indices, values, mask=None, uq=None):
        self.indices = indices
        self.values = values
        self.mask = mask
        self.uq = uq


def _init_table(shape, param_init, device, generator=None):
    if isinstance(param_init, torch.Tensor):
        assert tuple(param_init.shape) == tuple(shape)
        return param_init.to(device=device, dtype=torch.float32).contiguous()
    t = torch.empty(shape, dtype=torch.float32, device=device)
    if param_init == "normal":
        t.normal_(0.0, 0.01, generator=generator)  # initializer("normal") = N(0, 0.01^2)
    elif param_init in ("zeros", "zero"):
        t.zero_()
    elif param_init in ("ones", "one"):
        t.fill_(1.0)
    elif isinstance(param_init, (int, float)):
        t.fill_(float(param_init))
    else:
        raise ValueError("unsupported param_init %r" % (param_init,))
    return t


class EmbeddingLookup:
    """nn.EmbeddingLookup(vocab_size, embedding_size, param_init='normal', target='CPU',
    slice_mode='batch_slice', manual_shapes=None, max_norm=None, sparse=True, vocab_cache_size=0).

    Only target='DEVICE' exists here (there is no CPU path); slice modes other than batch_slice are
    provided by mindrec_b200.sharded.ShardedEmbedding."""

    BATCH_SLICE = "batch_slice"
    FIELD_SLICE = "field_slice"
    TABLE_ROW_SLICE = "table_row_slice"
    TABLE_COLUMN_SLICE = "table_column_slice"

    def __init__(self, vocab_size, embedding_size, param_init="normal", target="DEVICE",
                 slice_mode="batch_slice", manual_shapes=None, max_norm=None, sparse=True,
                 vocab_cache_size=0, device="cuda", name="embedding_table", generator=None, interleave_state=False):
        if not isinstance(vocab_size, int) or vocab_size <= 0:
            raise ValueError("For 'EmbeddingLookup', 'vocab_size' must be a positive int, got %r" % (vocab_size,))
        if not isinstance(embedding_size, int) or embedding_size <= 0:
            raise ValueError("For 'EmbeddingLookup', 'embedding_size' must be a positive int, got %r" % (embedding_size,))
        if target not in ("CPU", "DEVICE"):
            raise ValueError("For 'EmbeddingLookup', 'target' must be 'CPU' or 'DEVICE', got %r" % (target,))
        if target == "CPU":
            raise RuntimeError("mindrec_b200.EmbeddingLookup has no CPU path: use target='DEVICE'")
        if not isinstance(sparse, bool):
            raise TypeError("For 'EmbeddingLookup', 'sparse' must be bool")
        if slice_mode != self.BATCH_SLICE:
            raise ValueError("slice_mode %r: use mindrec_b200.sharded.ShardedEmbedding" % (slice_mode,))
        self.vocab_size = vocab_size
        self.embedding_size = embedding_size
        self.sparse = sparse
        self.max_norm = max_norm
        if interleave_state:
            # one [V, 3, D] array: w | m | v of a row are one 3*D*4-byte record, so the LazyAdam row update is one
            # random DRAM access per row instead of three (mrec_sparse_lazy_adam's interleaved form); the gathers read
            # array 0 at the record pitch.  m, v start at zero.
            if embedding_size % 4:
                raise ValueError("interleave_state needs embedding_size % 4 == 0")
            packed = torch.zeros((vocab_size, 3, embedding_size), dtype=torch.float32, device=device)
            packed[:, 0, :] = _init_table((vocab_size, embedding_size), param_init, device, generator)
            self.embedding_table = Parameter(packed[:, 0, :], name=name, packed=packed)
        else:
            self.embedding_table = Parameter(_init_table((vocab_size, embedding_size), param_init, device, generator),
                                             name=name)

    def __call__(self, indices):
        return self.construct(indices)

    def construct(self, indices):
        out = ops.gather(self.embedding_table.kernel_arg, indices)
        if self.max_norm is not None:
            norm = out.norm(dim=-1, keepdim=True).clamp_min(1e-12)
            out = out * torch.clamp(self.max_norm / norm, max=1.0)
        return out


# ------------------------------------------------------------------------------------------------
# optimizers
# ------------------------------------------------------------------------------------------------
def _is_map(p):
    """A mindrec_b200.hash.MapParameter (duck-typed: nn must not import hash)."""
    return hasattr(p, "tkeys") and hasattr(p, "add_arena")


class _Optimizer:
    """Parameters are `Parameter`s or `MapParameter`s.  On a MapParameter the rows are addressed by SLOT (the
    RowTensor's indices are the slots returned by the lookup), the state lives in sibling arenas registered with
    add_arena — so new keys start from fresh state (zeros / initial_accum) and a growth of the table moves it along —
    and the shared default row (slot C) is outside the updated range (upstream: the MapTensor branch of
    LazyAdam / FTRL, get(keys) -> row math -> put(keys))."""

    def __init__(self, params, loss_scale):
        self.parameters = list(params)
        if not self.parameters:
            raise ValueError("Optimizer got an empty parameter list")
        self.loss_scale = float(loss_scale)
        p0 = self.parameters[0]
        self.device = p0.device if _is_map(p0) else p0.data.device

    @staticmethod
    def _state_arenas(p, fills):
        first = len(p._arenas)
        for f in fills:
            p.add_arena(f)
        return tuple(range(first, first + len(fills)))

    @staticmethod
    def _rows(p, arena_ids):
        """(w, state...) as [C, D] views of the CURRENT arenas of a MapParameter."""
        c = p.capacity
        return [p.values[:c]] + [p.arena(i)[:c] for i in arena_ids]

    def _dedup(self, p, g):
        if g.uq is None:
            g.uq = ops.unique(g.indices, table_like=p.values[:p.capacity] if _is_map(p) else p.kernel_arg)
        return g.uq

    def __call__(self, grads):
        if len(grads) != len(self.parameters):
            raise ValueError("expected %d gradients, got %d" % (len(self.parameters), len(grads)))
        self.step(grads)
        return True


class Adam(_Optimizer):
    """nn.Adam(params, learning_rate=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, use_locking=False,
    use_nesterov=False, weight_decay=0.0, loss_scale=1.0).  RowTensor gradients are applied with the
    dense-equivalent rule (every row's moments decay) unless `lazy` (LazyAdam)."""
    lazy = False

    def __init__(self, params, learning_rate=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, use_locking=False,
                 use_nesterov=False, weight_decay=0.0, loss_scale=1.0):
        super().__init__(params, loss_scale)
        if not 0.0 < beta1 < 1.0 or not 0.0 < beta2 < 1.0:
            raise ValueError("beta1/beta2 must be in (0, 1)")
        if eps <= 0:
            raise ValueError("eps must be > 0")
        if use_nesterov or weight_decay != 0.0:
            raise NotImplementedError("use_nesterov / weight_decay are not used on the reference's hot path")
        self.hyper = ops.adam_hyper(learning_rate, beta1, beta2, eps, loss_scale, device=self.device)
        self.moment1, self.moment2, self._map_state = [], [], {}
        for i, p in enumerate(self.parameters):
            if _is_map(p):
                self._map_state[i] = self._state_arenas(p, (0.0, 0.0))
                self.moment1.append(None)
                self.moment2.append(None)
            elif getattr(p, "packed", None) is not None:      # interleaved record: the moments are views of it
                if not self.lazy:
                    raise ValueError("an interleaved table (interleave_state=True) is updated by LazyAdam only")
                self.moment1.append(p.packed[:, 1, :])
                self.moment2.append(p.packed[:, 2, :])
            else:
                self.moment1.append(torch.zeros_like(p.data))
                self.moment2.append(torch.zeros_like(p.data))

    def _row_flags(self, p):
        flags = getattr(p, "_row_flags", None)
        if flags is None:
            flags = torch.zeros(p.data.shape[0], dtype=torch.uint8, device=p.data.device)
            p._row_flags = flags
        return flags

    def step(self, grads):
        self.begin_step()
        for i, g in enumerate(grads):
            self.apply(i, g)

    def begin_step(self):
        """Advance the step count / bias-corrected learning rate (device side).  `step` = begin_step + apply(i, g) for
        every parameter; callers that overlap the updates of different parameters on forked streams call the two
        halves themselves."""
        ops.adam_begin_step(self.hyper)

    def apply(self, i, g):
        p, m, v = self.parameters[i], self.moment1[i], self.moment2[i]
        if i in self._map_state:
            if not isinstance(g, RowTensor) or not self.lazy:
                raise TypeError("a MapParameter takes a RowTensor gradient and LazyAdam (wide_and_deep.py:415-422)")
            w, m, v = self._rows(p, self._map_state[i])
            ops.sparse_lazy_adam(w, m, v, self.hyper, g.values, g.mask, self._dedup(p, g))
        elif isinstance(g, RowTensor):
            uq = self._dedup(p, g)
            if self.lazy and getattr(p, "packed", None) is not None:
                ops.sparse_lazy_adam(p.packed, None, None, self.hyper, g.values, g.mask, uq)
            elif self.lazy:
                ops.sparse_lazy_adam(p.data, m, v, self.hyper, g.values, g.mask, uq)
            else:
                ops.adam_rowsparse_dense_equiv(p.data, m, v, self.hyper, g.values, g.mask, uq,
                                               self._row_flags(p))
        else:
            ops.adam_dense(p.data, m, v, self.hyper, g)


class LazyAdam(Adam):
    """nn.LazyAdam: with RowTensor gradients only the looked-up rows move."""
    lazy = True


class FTRL(_Optimizer):
    """nn.FTRL(params, initial_accum=0.1, learning_rate=0.001, lr_power=-0.5, l1=0.0, l2=0.0,
    use_locking=False, loss_scale=1.0, weight_decay=0.0)."""

    def __init__(self, params, initial_accum=0.1, learning_rate=0.001, lr_power=-0.5, l1=0.0, l2=0.0,
                 use_locking=False, loss_scale=1.0, weight_decay=0.0):
        super().__init__(params, loss_scale)
        if initial_accum < 0:
            raise ValueError("initial_accum must be >= 0")
        if learning_rate <= 0:
            raise ValueError("learning_rate must be > 0")
        if lr_power > 0:
            raise ValueError("lr_power must be <= 0")
        if l1 < 0 or l2 < 0:
            raise ValueError("l1/l2 must be >= 0")
        self.hyper = ops.ftrl_hyper(learning_rate, l1, l2, lr_power, loss_scale, device=self.device)
        self.accum, self.linear, self._map_state = [], [], {}
        for i, p in enumerate(self.parameters):
            if _is_map(p):          # new keys: accum = initial_accum, linear = 0
                self._map_state[i] = self._state_arenas(p, (float(initial_accum), 0.0))
                self.accum.append(None)
                self.linear.append(None)
            else:
                self.accum.append(torch.full_like(p.data, float(initial_accum)))
                self.linear.append(torch.zeros_like(p.data))

    def step(self, grads):
        for i, (p, a, l, g) in enumerate(zip(self.parameters, self.accum, self.linear, grads)):
            if i in self._map_state:
                if not isinstance(g, RowTensor):
                    raise TypeError("a MapParameter takes a RowTensor gradient")
                w, a, l = self._rows(p, self._map_state[i])
                ops.sparse_ftrl(w, a, l, self.hyper, g.values, g.mask, self._dedup(p, g))
                continue
            if isinstance(g, RowTensor):
                ops.sparse_ftrl(p.data, a, l, self.hyper, g.values, g.mask, self._dedup(p, g))
            else:
                ops.ftrl_dense(p.data, a, l, self.hyper, g)


# ------------------------------------------------------------------------------------------------
# DenseLayer stack (library GEMMs: not a product kernel, SURVEY 2b last rows)
# ------------------------------------------------------------------------------------------------
class DenseStack:
    """The reference's DenseLayer chain (wide_and_deep.py:72-133): MatMul + BiasAdd + ReLU, optionally in
    fp16 (`convert_dtype`), with an explicit backward.  All weights / grads live in ONE flat fp32 buffer so
    that the dense Adam update is a single mrec_adam_dense launch.  `extra` reserves trailing scalars in
    the same buffer (W&D's Wide_b lands in the Adam group)."""

    @staticmethod
    def numel(dims, extra=0):
        return sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(len(dims) - 1)) + extra

    def __init__(self, dims, convert_dtype, device, generator=None, weight_init="normal", bias_init="zero",
                 last_activation=False, extra=0, storage=None):
        self.dims = list(dims)
        self.convert_dtype = convert_dtype
        self.last_activation = last_activation
        n = self.numel(dims, extra)
        if storage is None:
            self.flat = torch.zeros(n, dtype=torch.float32, device=device)
            self.flat_grad = torch.zeros(n, dtype=torch.float32, device=device)
        else:  # slices of a larger flat parameter / gradient buffer owned by the model
            self.flat, self.flat_grad = storage
            assert self.flat.numel() == n and self.flat_grad.numel() == n
        self.weights, self.biases, self.gw, self.gb = [], [], [], []
        o = 0
        for i in range(len(dims) - 1):
            k, m = dims[i], dims[i + 1]
            self.weights.append(self.flat[o:o + k * m].view(k, m)); self.gw.append(self.flat_grad[o:o + k * m].view(k, m)); o += k * m
            self.biases.append(self.flat[o:o + m]); self.gb.append(self.flat_grad[o:o + m]); o += m
        self.extra = self.flat[o:o + extra]
        self.extra_grad = self.flat_grad[o:o + extra]
        for w in self.weights:
            if weight_init == "normal":
                w.normal_(0.0, 0.01, generator=generator)
            elif weight_init == "uniform":
                bound = 1.0 / math.sqrt(w.shape[0])
                w.uniform_(-bound, bound, generator=generator)
        for b in self.biases:
            if bias_init == "normal":
                b.normal_(0.0, 0.01, generator=generator)
        self._acts = None
        self._flat16 = None
        self._mm_f32 = None
        self._head = False

    def _wgrad(self, h_in, g, out):
        """gw = h_in^T g for fp16 operands, fp32 accumulate.  With out_dtype the GEMM writes fp32 straight into the
        flat gradient buffer; older torch builds round to fp16 and cast."""
        if self._mm_f32 is None:
            try:
                torch.mm(h_in.t(), g, out_dtype=torch.float32, out=out)
                self._mm_f32 = True
                return
            except (TypeError, RuntimeError):
                self._mm_f32 = False
        if self._mm_f32:
            torch.mm(h_in.t(), g, out_dtype=torch.float32, out=out)
        else:
            out.copy_(torch.mm(h_in.t(), g))

    def refresh_half(self):
        """fp16 shadow of the weights (the per-layer Cast(weight, float16) of the reference, done once
        per step for the whole flat buffer)."""
        if self._flat16 is None:
            self._flat16 = torch.empty_like(self.flat, dtype=torch.float16)
            self.w16, self.b16 = [], []
            o = 0
            for i in range(len(self.dims) - 1):
                k, m = self.dims[i], self.dims[i + 1]
                self.w16.append(self._flat16[o:o + k * m].view(k, m)); o += k * m
                self.b16.append(self._flat16[o:o + m]); o += m
        self._flat16.copy_(self.flat)

    def _first_layer_parts(self, xs, w, b, relu):
        """Layer 0 on an input given as column blocks [B, k_j] (sum k_j = dims[0]) WITHOUT concatenating them: the
        blocks meet row blocks of the weight matrix, out = relu(b + sum_j x_j @ W[o_j : o_j + k_j])."""
        o, h = 0, None
        for x in xs:
            k = x.shape[1]
            h = torch.addmm(b, x, w[o:o + k]) if h is None else h.addmm_(x, w[o:o + k])
            o += k
        assert o == self.dims[0], "input blocks do not add up to the first layer's width"
        return torch.relu_(h) if relu else h

    def forward(self, x):
        """x: [B, dims[0]] fp32, or fp16 when convert_dtype (mrec_gather_masked can emit it directly) — or a tuple of
        column blocks of it (two lookups feeding one tower: no concatenation copy; backward then returns a tuple).
        Returns the stack output in fp32."""
        nl = len(self.weights)
        parts = isinstance(x, (tuple, list))
        if parts and (nl < 2 or (self.dims[-1] == 1 and nl == 1)):
            raise ValueError("an input in column blocks needs at least one hidden layer")
        acts = [x]
        # a one-unit output layer without activation runs as mrec_dense_head_fwd / _bwd instead of skinny GEMMs
        head = self.dims[-1] == 1 and not self.last_activation
        self._head = head
        if self.convert_dtype:
            self.refresh_half()
            if parts:
                h = acts[0] = tuple(t if t.dtype == torch.float16 else t.half() for t in x)
            else:
                h = x if x.dtype == torch.float16 else x.half()
                acts[0] = h
            for i in range(nl):
                if i == 0 and parts:
                    h = self._first_layer_parts(h, self.w16[0], self.b16[0], True)
                elif i + 1 < nl or self.last_activation:
                    h = torch._addmm_activation(self.b16[i], h, self.w16[i], use_gelu=False)
                elif head:
                    h = ops.dense_head_fwd(h, self.w16[i].view(-1), self.b16[i])
                else:
                    h = torch.addmm(self.b16[i], h, self.w16[i])
                acts.append(h)
            self._acts = acts
            return h if head else h.float()
        h = x
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            if i == 0 and parts:
                a = self._first_layer_parts(h, w, b, False)
            elif head and i + 1 == nl:
                a = ops.dense_head_fwd(h, w.view(-1), b)
            else:
                a = torch.addmm(b, h, w)
            if i + 1 < nl or self.last_activation:
                a = torch.relu_(a)
            h = a
            acts.append(h)
        self._acts = acts
        return h

    def backward(self, g_out, input_grad_dtype=None, on_weight_grads=None):
        """g_out: gradient wrt the stack output.  Fills flat_grad (fp32) and returns the gradient wrt the
        input — fp16 when convert_dtype (the sparse optimizers read fp16 rows directly).  on_weight_grads(): called
        once every weight / bias gradient has been issued, BEFORE the input-gradient GEMM of layer 0 — a data-parallel
        caller forks its gradient all-reduce there, underneath that GEMM."""
        acts = self._acts
        nl = len(self.weights)
        g = g_out.half() if (self.convert_dtype and g_out.dtype != torch.float16) else g_out
        premasked = False
        for i in range(nl - 1, -1, -1):
            h_in, h_out = acts[i], acts[i + 1]
            if self._head and i + 1 == nl:
                # output unit: rank-1 input gradient, weight / bias gradients and the previous layer's
                # ReluGrad + BiasAddGrad in one kernel
                w = self.w16[i] if self.convert_dtype else self.weights[i]
                g = ops.dense_head_bwd(g, h_in, w.view(-1), i > 0, self.gw[i].view(-1), self.gb[i],
                                       self.gb[i - 1] if i > 0 else None)
                premasked = True
                continue
            masked = i + 1 < nl or self.last_activation
            if premasked:
                premasked = False
            else:                  # ReluGrad + BiasAddGrad in one mrec_relu_bwd_bias pass, fp32 sums in place
                g = ops.relu_bwd_bias(g, h_out if masked else None, self.gb[i])
            w = self.w16[i] if self.convert_dtype else self.weights[i]
            xs = h_in if isinstance(h_in, tuple) else (h_in,)      # layer 0 may be fed in column blocks
            o = 0
            for x in xs:                             # weight gradients: row blocks of gw
                k = x.shape[1]
                if self.convert_dtype:
                    self._wgrad(x, g, self.gw[i][o:o + k])
                else:
                    torch.mm(x.t(), g, out=self.gw[i][o:o + k])
                o += k
            if i == 0 and on_weight_grads is not None:
                on_weight_grads()
            o, gxs = 0, []
            for x in xs:                             # input gradients: row blocks of w, one per block
                k = x.shape[1]
                gxs.append(torch.mm(g, w[o:o + k].t()))
                o += k
            g = tuple(gxs) if isinstance(h_in, tuple) else gxs[0]
        return g
