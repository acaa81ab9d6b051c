"""Functional wrappers over the aot entry points (one Python function per exported symbol).

Each wrapper only allocates outputs / workspaces (torch is the allocator here, MindSpore would be under
``ops.Custom``) and forwards to the C-ABI; no arithmetic happens in Python.  Workspaces are cached per
(symbol, size, device) so steady-state steps allocate nothing and can be captured in a CUDA graph.
"""
import ctypes

from . import _lib

_ws_cache = {}
HYPER_LEN = 16


# ---- allocation: the wrappers only ever allocate OUTPUTS and WORKSPACES next to their inputs.  The allocator is picked
# from the input's device: a mindrec_b200.runtime.Device (plain CUDA allocations, no torch in the process) or a torch
# device (tests, bench, the model cells).  Nothing below imports torch unless it is handed a torch tensor.
def _dt(dtype):
    """Canonical dtype name ('float32', 'int32', ...) of a torch dtype or a name."""
    return dtype if isinstance(dtype, str) else str(dtype).replace("torch.", "")


def _alloc(device, shape, dtype="float32", zero=False):
    shape = tuple(shape) if isinstance(shape, (tuple, list)) else (int(shape),)
    if getattr(device, "__mrec_rt__", False):
        return device.zeros(shape, _dt(dtype)) if zero else device.empty(shape, _dt(dtype))
    import torch
    td = getattr(torch, _dt(dtype))
    return torch.zeros(shape, dtype=td, device=device) if zero else torch.empty(shape, dtype=td, device=device)


def _from_list(device, values, dtype="float32"):
    if getattr(device, "__mrec_rt__", False):
        return device.tensor(values, dtype)
    import torch
    return torch.tensor(values, dtype=getattr(torch, dtype), device=device)


_ws_retired = []


def _ws(tag, nbytes, device):
    """Cached workspace for (tag, device, CURRENT STREAM).  Keyed by the stream because two calls that share a tag may
    run concurrently on forked streams (parallel branches of a captured graph): calls on one stream are ordered, calls
    on different streams get different buffers.  A buffer that has to grow is replaced, and the old one is kept alive
    for the life of the process: captured CUDA graphs and kernels still queued hold raw pointers to it."""
    if getattr(device, "__mrec_rt__", False):
        stream = device.current_stream_handle()
    else:
        import torch
        stream = torch.cuda.current_stream(device).cuda_stream if torch.cuda.is_available() else 0
    key = (tag, device, stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _ws_retired.append(buf)
        buf = _alloc(device, max(int(nbytes), 256), "uint8")
        _ws_cache[key] = buf
    return buf


def _dummy(device):
    key = ("dummy", device)
    buf = _ws_cache.get(key)
    if buf is None:
        buf = _alloc(device, 1, "int32", zero=True)
        _ws_cache[key] = buf
    return buf


def _size_fn(name):
    f = getattr(_lib.lib(), name)
    f.restype = ctypes.c_size_t
    f.argtypes = [ctypes.c_int64, ctypes.c_int]
    return f


# ------------------------------------------------------------------------------------------------
# K1 gather
# ------------------------------------------------------------------------------------------------
def gather(table, ids, out=None, oob_flag=None):
    """out[..., :] = table[ids[...], :]  (mrec_gather; nn.EmbeddingLookup / P.Gather axis 0).  A [V,K,D] table is the
    interleaved record layout: array 0 of every row is read."""
    dim = table.shape[-1] if table.dim() >= 2 else 1
    if out is None:
        out = _alloc(table.device, tuple(ids.shape) + (dim,), "float32")
    args = [table, ids, out] + ([oob_flag] if oob_flag is not None else [])
    _lib.aot_call("mrec_gather", args)
    return out


def gather_masked(table, ids, mask, out=None, oob_flag=None, out_dtype="float32"):
    """out[b, f*D:(f+1)*D] = table[ids[b,f]] * mask[b,f]  (gather + Mul + Reshape fused).
    A float16 `out` additionally fuses the Cast at the head of a mixed-precision DenseLayer."""
    dim = table.shape[-1] if table.dim() >= 2 else 1
    if out is None:
        out = _alloc(table.device, (ids.shape[0], ids.numel() // ids.shape[0] * dim), out_dtype)
    args = [table, ids, mask, out] + ([oob_flag] if oob_flag is not None else [])
    _lib.aot_call("mrec_gather_masked", args)
    return out


def gather_reduce(table, ids, mask, bias, out=None, oob_flag=None):
    """out[b] = sum_f table[ids[b,f]] * mask[b,f] + bias  for a dim-1 table (wide / linear term)."""
    if out is None:
        out = _alloc(table.device, (ids.shape[0], 1), "float32")
    args = [table, ids, mask, bias, out] + ([oob_flag] if oob_flag is not None else [])
    _lib.aot_call("mrec_gather_reduce", args)
    return out


def gather_pool(table, ids, mask, out=None, oob_flag=None):
    """out[b] = mean over the S slots of table[ids[b,s]] * mask[b,s]  (multi-hot fields of the multitable model)."""
    if out is None:
        out = _alloc(table.device, (ids.shape[0], table.shape[1]), "float32")
    args = [table, ids, mask, out] + ([oob_flag] if oob_flag is not None else [])
    _lib.aot_call("mrec_gather_pool", args)
    return out


# ------------------------------------------------------------------------------------------------
# K2 unique
# ------------------------------------------------------------------------------------------------
class UniqueResult:
    """Outputs of mrec_unique (all padded to N; `count` is a device scalar).  packed=True (int32 keys) carves all
    of them out of ONE buffer (`flat`), so a whole result can be copied with a single device copy."""
    __slots__ = ("uniq", "inverse", "count", "perm", "seg_start", "seg_of", "n", "flat")

    def __init__(self, n, dtype, device, packed=False):
        self.n = n
        self.flat = None
        if packed and _dt(dtype) == "int32":
            a = (n + 3) // 4 * 4                              # every field starts 16-byte aligned
            b = (n + 1 + 3) // 4 * 4
            self.flat = _alloc(device, 4 * a + b + 4, "int32", zero=True)
            f = self.flat
            self.uniq, self.inverse, self.perm, self.seg_of = f[0:n], f[a:a + n], f[2 * a:2 * a + n], f[3 * a:3 * a + n]
            self.seg_start = f[4 * a:4 * a + n + 1]
            self.count = f[4 * a + b:4 * a + b + 1]
            return
        self.uniq = _alloc(device, n, dtype)
        self.inverse = _alloc(device, n, "int32")
        self.count = _alloc(device, 1, "int32", zero=True)
        self.perm = _alloc(device, n, "int32")
        self.seg_start = _alloc(device, n + 1, "int32")
        self.seg_of = _alloc(device, n, "int32")

    def copy_from(self, other):
        if self.flat is not None and other.flat is not None and self.flat.numel() == other.flat.numel():
            self.flat.copy_(other.flat)
        else:
            for d, s_ in zip(self.outputs(), other.outputs()):
                d.copy_(s_)

    def outputs(self):
        return [self.uniq, self.inverse, self.count, self.perm, self.seg_start, self.seg_of]

    def sliced(self, n):
        """A view of these buffers for a dedup of n <= self.n keys (data-dependent sizes without reallocation)."""
        v = UniqueResult.__new__(UniqueResult)
        v.n = n
        v.flat = None
        v.uniq, v.inverse, v.count = self.uniq[:n], self.inverse[:n], self.count
        v.perm, v.seg_start, v.seg_of = self.perm[:n], self.seg_start[:n + 1], self.seg_of[:n]
        return v


def unique(ids, table_like=None, result=None, ws_tag="unique", n_valid=None):
    """Ascending unique + inverse + stable sort permutation + segment map.

    With `table_like` (any tensor whose dim 0 is the table's row count V) only ceil(log2(V+1)) key bits
    are sorted and ids outside [0, V) collapse onto the value V (mrec_unique_bounded).  `ws_tag` names the
    cached workspace: calls that may overlap on different streams must use different tags.
    """
    flat = ids.reshape(-1)
    n = flat.numel()
    if result is None:
        result = UniqueResult(n, flat.dtype, flat.device)
    nbytes = _size_fn("mrec_unique_workspace_bytes")(n, flat.element_size())
    ws = _ws(ws_tag, nbytes, flat.device)
    if table_like is None:
        _lib.aot_call("mrec_unique", [flat] + result.outputs() + [ws])
    elif n_valid is None:
        _lib.aot_call("mrec_unique_bounded", [flat, table_like] + result.outputs() + [ws])
    else:
        _lib.aot_call("mrec_unique_bounded", [flat, table_like, n_valid] + result.outputs() + [ws])
    return result


def unique_first(ids):
    """First-occurrence-order unique (the order of upstream's CPU Unique): (uniq[N], inverse[N], count[1])."""
    flat = ids.reshape(-1)
    n = flat.numel()
    uniq = _alloc(flat.device, n, flat.dtype)
    inverse = _alloc(flat.device, n, "int32")
    count = _alloc(flat.device, 1, "int32", zero=True)
    nbytes = _size_fn("mrec_unique_first_workspace_bytes")(n, flat.element_size())
    ws = _ws("unique_first", nbytes, flat.device)
    _lib.aot_call("mrec_unique_first", [flat, uniq, inverse, count, ws])
    return uniq, inverse, count


# ------------------------------------------------------------------------------------------------
# K3-K5 segment-sum + sparse optimizers
# ------------------------------------------------------------------------------------------------
_EMPTY = {}


def _empty_mask(device):
    m = _EMPTY.get(device)
    if m is None:
        m = _alloc(device, 0, "float32")
        _EMPTY[device] = m
    return m


def _opt_ws(n, dim, device):
    # one workspace per row width: updates of different tables may run concurrently on forked streams
    nbytes = _size_fn("mrec_sparse_opt_workspace_bytes")(n, dim)
    return _ws("sparse_opt_%d" % dim, nbytes, device)


def segment_sum(g, mask, uq, dim=None, out=None):
    """gsum[u] = sum over positions n of segment u of mask[n] * g[n // div]  (sorted order, no atomics)."""
    dim = dim if dim is not None else (g.shape[-1] if g.dim() >= 2 else 1)
    if out is None:
        out = _alloc(g.device, (uq.n, dim), "float32", zero=True)
    mask = _empty_mask(g.device) if mask is None else mask.reshape(-1)
    _lib.aot_call("mrec_segment_sum", [g, mask, uq.perm, uq.seg_start, uq.seg_of, out,
                                       _opt_ws(uq.n, dim, g.device)])
    return out


def segment_sum_to_peers(g, mask, uq, my_bounds, inbox_off, peer_ptrs, cap_like, err, dim=None):
    """Fused segment-sum + gradient push: the sum of segment u goes straight into the inbox of the rank that owns u's
    key (peer-mapped pointers), without a local gsum buffer (mrec_segment_sum_to_peers)."""
    dim = dim if dim is not None else g.shape[-1]
    mask = _empty_mask(g.device) if mask is None else mask.reshape(-1)
    nbytes = _size_fn("mrec_segment_sum_workspace_bytes")(uq.n, dim)
    _lib.aot_call("mrec_segment_sum_to_peers", [g, mask, uq.perm, uq.seg_start, uq.seg_of, my_bounds, inbox_off, peer_ptrs,
                                                cap_like, err, _ws("segsum_peers_%d" % dim, nbytes, g.device)])


def segment_sum_scatter_add(table, g, mask, uq):
    """table[uq.uniq[u]] += segment sum u (dense gradient of a non-sparse Gather; accumulates over calls)."""
    dim = table.shape[1] if table.dim() == 2 else 1
    mask = _empty_mask(table.device) if mask is None else mask.reshape(-1)
    _lib.aot_call("mrec_segment_sum_scatter_add", [g, mask, uq.uniq, uq.perm, uq.seg_start, uq.seg_of, table,
                                                   _opt_ws(uq.n, dim, table.device)])
    return table


def sparse_lazy_adam(w, m, v, hyper, g, mask, uq, n_valid=None):
    """Fused segment-sum + LazyAdam row update on the rows named by uq.uniq (in place).
    n_valid (device int32[1]): only the first n_valid sorted positions are real (static inbox).
    m = v = None: `w` is the interleaved record array wmv[V,3,D] (weights | m | v of a row back to back)."""
    dim = w.shape[-1] if w.dim() >= 2 else 1
    if m is None:
        if not (w.dim() == 3 and w.shape[1] == 3):
            raise ValueError("sparse_lazy_adam without m, v needs the interleaved array wmv[V,3,D]")
        m = v = _empty_mask(w.device)
    mask = _empty_mask(w.device) if mask is None else mask.reshape(-1)
    nv = [] if n_valid is None else [n_valid]
    _lib.aot_call("mrec_sparse_lazy_adam", [w, m, v, hyper, g, mask, uq.uniq, uq.perm, uq.seg_start,
                                            uq.seg_of] + nv + [_dummy(w.device), _opt_ws(uq.n, dim, w.device)])


def adam_rowsparse_dense_equiv(w, m, v, hyper, g, mask, uq, row_flags):
    """nn.Adam with a RowTensor gradient: dense-equivalent update of every row (in place)."""
    dim = w.shape[1] if w.dim() == 2 else 1
    mask = _empty_mask(w.device) if mask is None else mask.reshape(-1)
    _lib.aot_call("mrec_adam_rowsparse", [w, m, v, hyper, g, mask, uq.uniq, uq.perm, uq.seg_start,
                                          uq.seg_of, row_flags, _dummy(w.device),
                                          _opt_ws(uq.n, dim, w.device)])


def sparse_ftrl(w, accum, linear, hyper, g, mask, uq, n_valid=None):
    """Fused segment-sum + FTRL row update on the rows named by uq.uniq (in place)."""
    dim = w.shape[1] if w.dim() == 2 else 1
    mask = _empty_mask(w.device) if mask is None else mask.reshape(-1)
    nv = [] if n_valid is None else [n_valid]
    _lib.aot_call("mrec_sparse_ftrl", [w, accum, linear, hyper, g, mask, uq.uniq, uq.perm, uq.seg_start,
                                       uq.seg_of] + nv + [_dummy(w.device), _opt_ws(uq.n, dim, w.device)])


def adam_hyper(lr, beta1=0.9, beta2=0.999, eps=1e-8, loss_scale=1.0, l2=0.0, device="cuda"):
    """Device hyper block for Adam / LazyAdam: [lr, b1, b2, eps, b1^t, b2^t, lr_t, 1/loss_scale, l2, 0...]."""
    return _from_list(device, [lr, beta1, beta2, eps, 1.0, 1.0, 0.0, 1.0 / loss_scale, l2] + [0.0] * 7)


def ftrl_hyper(lr, l1=0.0, l2=0.0, lr_power=-0.5, loss_scale=1.0, device="cuda"):
    """Device hyper block for FTRL: [lr, l1, l2, lr_power, 1/loss_scale, 0...]."""
    return _from_list(device, [lr, l1, l2, lr_power, 1.0 / loss_scale] + [0.0] * 11)


def adam_begin_step(hyper):
    _lib.aot_call("mrec_adam_begin_step", [hyper, _dummy(hyper.device)])


def adam_dense(w, m, v, hyper, g):
    _lib.aot_call("mrec_adam_dense", [w, m, v, hyper, g, _dummy(w.device)])


def ftrl_dense(w, accum, linear, hyper, g):
    _lib.aot_call("mrec_ftrl_dense", [w, accum, linear, hyper, g, _dummy(w.device)])


# ------------------------------------------------------------------------------------------------
# K7 FM second-order interaction, K8 DCN cross stack
# ------------------------------------------------------------------------------------------------
def fm_fwd(vx, out=None):
    """fm[b] = 0.5 * sum_d[(sum_f vx)^2 - sum_f vx^2]; vx is [B,F,D] (already masked)."""
    if out is None:
        out = _alloc(vx.device, (vx.shape[0], 1), "float32")
    _lib.aot_call("mrec_fm_fwd", [vx, out])
    return out


def fm_bwd(vx, gout, out=None, addend=None):
    """dvx[b,f,d] = gout[b] * (S[b,d] - vx[b,f,d]) (+ addend[b,f,d], fp32 or fp16, fused into the store)."""
    if out is None:
        out = _alloc(vx.device, vx.shape, vx.dtype)
    _lib.aot_call("mrec_fm_bwd", [vx, gout] + ([addend] if addend is not None else []) + [out])
    return out


def _cross_ws(layers, dp, device):
    f = getattr(_lib.lib(), "mrec_cross_workspace_bytes")
    f.restype = ctypes.c_size_t
    f.argtypes = [ctypes.c_int64, ctypes.c_int]
    return _ws("cross", f(layers, dp), device)


def cross_fwd(x0, w, b, y=None, p=None):
    """The whole cross stack: y = x_L, p[B,L] = saved dots x0.w_l (needed by cross_bwd).  w, b: [L, D']."""
    layers = w.numel() // x0.shape[1]
    if y is None:
        y = _alloc(x0.device, x0.shape, x0.dtype)
    if p is None:
        p = _alloc(x0.device, (x0.shape[0], layers), "float32")
    _lib.aot_call("mrec_cross_fwd", [x0, w, b, y, p, _cross_ws(layers, x0.shape[1], x0.device)])
    return y, p


def cross_bwd(x0, dy, w, b, p, dx=None, dw=None, db=None):
    """Backward of the cross stack: (dx0 total, dw[L,D'], db[L,D'])."""
    layers = w.numel() // x0.shape[1]
    dx = _alloc(x0.device, x0.shape, x0.dtype) if dx is None else dx
    dw = _alloc(x0.device, (layers, x0.shape[1]), "float32") if dw is None else dw
    db = _alloc(x0.device, (layers, x0.shape[1]), "float32") if db is None else db
    _lib.aot_call("mrec_cross_bwd", [x0, dy, w, b, p, dx, dw, db, _cross_ws(layers, x0.shape[1], x0.device)])
    return dx, dw, db


def shard_bounds(uniq, count, edges, out=None):
    """bounds[r] = #{i < count : uniq[i] < edges[r]} (device-side lower_bound, no host sync)."""
    if out is None:
        out = _alloc(uniq.device, edges.numel(), "int32")
    _lib.aot_call("mrec_shard_bounds", [uniq, count, edges, out])
    return out


def shard_remap(ids, table_like, owners_like, out=None):
    """key -> (key mod G) * R + key div G with G, R = owners_like.shape (out-of-range keys -> G * R)."""
    if out is None:
        out = _alloc(ids.device, ids.shape, ids.dtype)
    _lib.aot_call("mrec_shard_remap", [ids, table_like, owners_like, out])
    return out


def sigmoid_xent(a, b, label, sens, out=None, half=False):
    """Fused logit = a + b, mean sigmoid cross-entropy, delta = sens * (sigmoid(logit) - label) / B, sum(delta).
    `out` = (logit, loss, delta, delta16, delta_sum) preallocated, or None."""
    n = a.numel()
    dev = a.device
    if out is None:
        out = (_alloc(dev, (n, 1), "float32"), _alloc(dev, 1, "float32"),
               _alloc(dev, (n, 1), "float32"),
               _alloc(dev, (n, 1) if half else (0,), "float16"),
               _alloc(dev, 1, "float32"))
    b = _empty_mask(dev) if b is None else b
    _lib.aot_call("mrec_sigmoid_xent", [a, b, label, sens, out[0], out[1], out[2], out[3], out[4]])
    return out


# Workspaces of the DenseLayer glue kernels are cached per (device, width): calls that share a width must be
# stream-ordered (they are: one DenseStack runs its layers on one stream).
_dense_ws = {}


def relu_bwd_bias(g, y, gb, out=None):
    """gz = g * (y > 0) written over g (or into `out`), gb[N] = column sums of gz in fp32 (ReluGrad + BiasAddGrad of a
    DenseLayer in one pass).  y=None: plain BiasAddGrad, g untouched.  Returns gz."""
    n_cols = g.shape[1]
    key = (g.device, n_cols)
    ws = _dense_ws.get(key)
    if ws is None:
        lib = _lib.lib()
        lib.mrec_relu_bwd_bias_workspace_bytes.restype = ctypes.c_size_t
        lib.mrec_relu_bwd_bias_workspace_bytes.argtypes = [ctypes.c_int64]
        ws = _dense_ws[key] = _alloc(g.device, lib.mrec_relu_bwd_bias_workspace_bytes(n_cols), "uint8", zero=True)
    gz = g if out is None else out
    if y is None:
        _lib.aot_call("mrec_relu_bwd_bias", [g, _empty_like_dtype(g), _empty_like_dtype(g), gb, ws])
        return g
    _lib.aot_call("mrec_relu_bwd_bias", [g, y, gz, gb, ws])
    return gz


def dense_head_fwd(h, w, bias, out=None):
    """out[b] = sum_k h[b,k] * w[k] + bias  (the one-unit output DenseLayer; fp32 accumulate and result)."""
    if out is None:
        out = _alloc(h.device, (h.shape[0], 1), "float32")
    _lib.aot_call("mrec_dense_head_fwd", [h, w, bias, out])
    return out


_head_ws = {}


def dense_head_bwd(delta, h, w, masked, gw, gb_head, gb_prev=None, out=None):
    """Backward of the one-unit output layer fused with the previous layer's ReluGrad + BiasAddGrad:
    gh = (delta x w) * (h > 0 if masked), gw = delta^T h, gb_head = sum(delta), gb_prev = column sums of gh."""
    k = h.shape[1]
    key = (h.device, k)
    ws = _head_ws.get(key)
    if ws is None:
        lib = _lib.lib()
        lib.mrec_dense_head_workspace_bytes.restype = ctypes.c_size_t
        lib.mrec_dense_head_workspace_bytes.argtypes = [ctypes.c_int64]
        ws = _head_ws[key] = _alloc(h.device, lib.mrec_dense_head_workspace_bytes(k), "uint8", zero=True)
    if out is None:
        out = _alloc(h.device, h.shape, h.dtype)
    flag = _relu_flag(h.device) if masked else _empty_mask(h.device)
    _lib.aot_call("mrec_dense_head_bwd", [delta, h, w, flag, out, gw, gb_head,
                                          gb_prev if gb_prev is not None else _empty_mask(h.device), ws])
    return out


_relu_flags = {}


def _relu_flag(device):
    f = _relu_flags.get(device)
    if f is None:
        f = _relu_flags[device] = _from_list(device, [1.0])
    return f


_empties = {}


def _empty_like_dtype(t):
    key = (t.device, t.dtype)
    e = _empties.get(key)
    if e is None:
        e = _empties[key] = _alloc(t.device, (0, 0), t.dtype)
    return e


_mode_like = {}


def gather_to_peers(table, rows, peer_ptrs, dst_off, src_off, dirty=None, mode=0):
    """Owner-side gather fused with the NVLink peer store of every row into its requester's landing buffer.
    dirty (int32 bitmap, one bit per table row) + mode: 1 = only rows whose bit is clear, 2 = only rows whose bit is set."""
    if dirty is None or mode == 0:
        _lib.aot_call("mrec_gather_to_peers", [table, rows, peer_ptrs, dst_off, src_off, _dummy(table.device)])
        return
    key = (table.device, mode)
    ml = _mode_like.get(key)
    if ml is None:
        ml = _mode_like[key] = _alloc(table.device, (mode, 0), "float32")
    _lib.aot_call("mrec_gather_to_peers", [table, rows, peer_ptrs, dst_off, src_off, dirty, ml, _dummy(table.device)])


def bitmap_set(rows, count, bitmap):
    """bitmap bit r |= 1 for r in rows[:count] (device-side count)."""
    _lib.aot_call("mrec_bitmap_set", [rows, count, bitmap])


def shard_offsets(bounds_all, ctrl, dst_off, src_off, inbox_off, n_r):
    _lib.aot_call("mrec_shard_offsets", [bounds_all, ctrl, dst_off, src_off, inbox_off, n_r])


def shard_remap_hash(keys, owners_like, bits_like, out=None):
    """key -> (owner << B) | key with owner = hash(key) mod G (G, B = dim 0 of the two shape carriers)."""
    if out is None:
        out = _alloc(keys.device, keys.shape, "int64")
    _lib.aot_call("mrec_shard_remap_hash", [keys, owners_like, bits_like, out])
    return out


def fill_tail(buf, n_valid, value):
    """buf[i] = value for i >= n_valid (device-side count)."""
    _lib.aot_call("mrec_fill_tail", [n_valid, value, buf])
    return buf


def push_rows_to_peers(rows, my_bounds, inbox_off, peer_ptrs, cap_like, mod_like, err):
    _lib.aot_call("mrec_push_rows_to_peers", [rows, my_bounds, inbox_off, peer_ptrs, cap_like, mod_like, err])


def peer_signal(payload, payload_ptrs, flag_ptrs, epoch):
    _lib.aot_call("mrec_peer_signal", [payload, payload_ptrs, flag_ptrs, epoch, _dummy(epoch.device)])


def cast_f32_f16(src, out=None):
    """out = fp16(src) for a flat fp32 buffer (the per-step Cast of the mixed-precision DenseLayers' weights)."""
    if out is None:
        out = _alloc(src.device, src.shape, "float16")
    _lib.aot_call("mrec_cast_f32_f16", [src, out])
    return out


def zero_row(buf, idx):
    """buf[idx[0]] = 0 (device-side index; a no-op when it lies outside buf)."""
    _lib.aot_call("mrec_zero_row", [idx, buf])


def peer_allreduce(src_ptrs, dst_ptrs, ctrl, dst):
    """dst (on every rank) = sum over ranks of the source buffers, added in rank order by the slice owners."""
    _lib.aot_call("mrec_peer_allreduce", [src_ptrs, dst_ptrs, ctrl, dst])


def peer_wait(flags, epoch, err, max_cycles_log2=None):
    """Spin (bounded) until every flag slot reached epoch; a time-out sets bit 0 of err instead of hanging."""
    if max_cycles_log2 is None:
        _lib.aot_call("mrec_peer_wait", [flags, epoch, err])
    else:
        lim = _from_list(flags.device, [max_cycles_log2], "int32")
        _lib.aot_call("mrec_peer_wait", [flags, epoch, lim, err])
