"""CPU stand-ins for the kernels peer_sharded.PeerRank drives (test infrastructure only, never on the product path).

Same call surface as mindrec_b200.ops, arithmetic from the numpy oracle, peer stores done through the REAL
addresses the protocol computes (CPU tensors have real `data_ptr()`s; stores go through ctypes), so the pointer
tables built by `PeerRank.connect` — base + byte offsets per rank and phase — are exercised for real.
tests/test_peer_protocol_cpu.py monkeypatches `peer_sharded.ops` with this module.
"""
import ctypes
import types

import numpy as np
import torch

from mindrec_b200.ops import UniqueResult, adam_hyper, ftrl_hyper   # pure torch: no library call  # noqa: F401
from oracle import ref_numpy as R

F32 = np.float32


def _np(t):
    return t.detach().numpy()


def _at(addr, n, ctype, dtype):
    """numpy view of n elements at a raw address (a peer buffer as the protocol addresses it)."""
    return np.frombuffer((ctype * n).from_address(int(addr)), dtype=dtype)


# ---- plan ----------------------------------------------------------------------------------------------------------
def shard_remap(ids, table_like, owners_like, out=None):
    v, g, r = table_like.shape[0], owners_like.shape[0], owners_like.shape[1]
    k = _np(ids).astype(np.int64)
    ok = (k >= 0) & (k < v)
    res = np.where(ok, (k % g) * r + k // g, g * r)
    if out is None:
        out = torch.empty_like(ids)
    out.copy_(torch.from_numpy(res.astype(_np(ids).dtype)).view(out.shape))
    return out


def unique(ids, table_like=None, result=None, ws_tag="unique", n_valid=None):
    flat = _np(ids).reshape(-1)
    n = flat.size
    nv = n if n_valid is None else max(0, min(n, int(n_valid[0])))
    bound = table_like.shape[0] if table_like is not None else None
    uniq, inverse, perm, seg_start = R.unique_sorted(flat[:nv], bound=bound)
    if result is None:
        result = UniqueResult(n, ids.dtype, ids.device)
    u = uniq.size
    _np(result.uniq)[:u] = uniq
    _np(result.inverse)[:nv] = inverse
    _np(result.perm)[:nv] = perm
    _np(result.seg_start)[:u + 1] = seg_start
    seg_of = np.empty(nv, dtype=np.int32)
    seg_of[:] = inverse[perm] if nv else 0
    _np(result.seg_of)[:nv] = seg_of
    result.count[0] = u
    return result


def shard_bounds(uniq, count, edges, out=None):
    u = _np(uniq)[:int(count[0])]
    res = np.searchsorted(u, _np(edges), side="left").astype(np.int32)
    if out is None:
        out = torch.empty(edges.numel(), dtype=torch.int32)
    out.copy_(torch.from_numpy(res))
    return out


def shard_offsets(bounds_all, ctrl, dst_off, src_off, inbox_off, n_r):
    me, g = int(ctrl[0]), int(ctrl[1])
    d, s, i, n = R.shard_exchange_offsets(_np(bounds_all).reshape(g, g + 1), me)
    dst_off.copy_(torch.from_numpy(d.astype(np.int32)))
    src_off.copy_(torch.from_numpy(s.astype(np.int32)))
    inbox_off.copy_(torch.from_numpy(i.astype(np.int32)))
    n_r[0] = n


# ---- peer stores ---------------------------------------------------------------------------------------------------
def push_rows_to_peers(rows, my_bounds, inbox_off, peer_ptrs, cap_like, mod_like, err):
    b = _np(my_bounds)
    g = peer_ptrs.numel()
    src = _np(rows).reshape(rows.shape[0], -1)
    width = src.shape[1]
    cap, mod = cap_like.shape[0], mod_like.shape[0]
    ct, dt = (ctypes.c_int32, np.int32) if rows.dtype == torch.int32 else \
             (ctypes.c_int64, np.int64) if rows.dtype == torch.int64 else (ctypes.c_float, np.float32)
    for o in range(g):
        seg = src[b[o]:b[o + 1]]
        if mod > 0:
            seg = seg % mod
        for j in range(seg.shape[0]):
            drow = int(inbox_off[o]) + j
            if drow >= cap:
                err[0] = int(err[0]) | 2
                continue
            dst = _at(int(peer_ptrs[o]) + drow * width * ctypes.sizeof(ct), width, ct, dt)
            dst[:] = seg[j]


def segment_sum_to_peers(g, mask, uq, my_bounds, inbox_off, peer_ptrs, cap_like, err, dim=None):
    """Stand-in of the fused kernel: segment sums, then the rows of the valid segments through the real addresses."""
    dim = dim if dim is not None else g.shape[-1]
    gs = torch.zeros((uq.n, dim), dtype=torch.float32)
    segment_sum(g, mask, uq, dim=dim, out=gs)
    push_rows_to_peers(gs, my_bounds, inbox_off, peer_ptrs, cap_like, torch.empty((0, 0)), err)


def gather_to_peers(table, rows, peer_ptrs, dst_off, src_off, dirty=None, mode=0):
    t = _np(table)
    d = t.shape[1]
    g = peer_ptrs.numel()
    r = _np(rows)
    bits = _np(dirty).view(np.uint32) if dirty is not None else None
    for s in range(g):
        lo, hi = int(src_off[s]), min(int(src_off[s + 1]), r.size)      # like the kernel: never past the inbox
        for j in range(max(0, hi - lo)):
            row = int(r[lo + j])
            ok = 0 <= row < t.shape[0]
            if mode and ok and bool((int(bits[row >> 5]) >> (row & 31)) & 1) != (mode == 2):
                continue                                                 # the other half of the split serve moves this row
            dst = _at(int(peer_ptrs[s]) + (int(dst_off[s]) + j) * d * 4, d, ctypes.c_float, np.float32)
            dst[:] = t[row] if ok else 0.0


def bitmap_set(rows, count, bitmap):
    b = _np(bitmap).view(np.uint32)
    for row in _np(rows)[:int(count[0])].tolist():
        if 0 <= row < b.size * 32:
            b[row >> 5] |= np.uint32(1 << (row & 31))


def peer_signal(payload, payload_ptrs, flag_ptrs, epoch):
    e = int(epoch[0]) + 1
    p = _np(payload)
    for s in range(flag_ptrs.numel()):
        if p.size:
            _at(int(payload_ptrs[s]), p.size, ctypes.c_int32, np.int32)[:] = p
        _at(int(flag_ptrs[s]), 1, ctypes.c_int32, np.int32)[0] = e
    epoch[0] = e


def peer_wait(flags, epoch, err, max_cycles_log2=None):
    if bool((flags < int(epoch[0])).any()):          # on the device this spins; phase-major emulation never should
        err[0] = int(err[0]) | 1


# ---- lookups / updates ---------------------------------------------------------------------------------------------
def gather_masked(table, ids, mask, out=None, oob_flag=None, out_dtype=torch.float32):
    res = R.gather_masked(_np(table), _np(ids), _np(mask))
    out.copy_(torch.from_numpy(np.asarray(res, dtype=F32)).view(out.shape))
    return out


def gather_reduce(table, ids, mask, bias, out=None, oob_flag=None):
    res = R.gather_reduce(_np(table), _np(ids), _np(mask), _np(bias))
    out.copy_(torch.from_numpy(np.asarray(res, dtype=F32)).view(out.shape))
    return out


def _segment_rows(g, mask, uq, dim, n_valid=None):
    n = uq.n if n_valid is None else max(0, min(uq.n, int(n_valid[0])))
    u = int(uq.count[0])
    rows = _np(g).reshape(-1, dim)                     # the padding past n_valid is uninitialised: never cast it
    div = max(1, uq.n // max(1, rows.shape[0])) if rows.shape[0] and uq.n % rows.shape[0] == 0 else 1
    perm, seg_of = _np(uq.perm)[:n].astype(np.int64), _np(uq.seg_of)[:n].astype(np.int64)
    vals = rows[perm // div].astype(np.float64)
    if mask is not None and mask.numel():
        vals = vals * _np(mask).reshape(-1).astype(np.float64)[perm][:, None]
    out = np.zeros((u, dim))
    np.add.at(out, seg_of, vals)
    return out, u


def segment_sum(g, mask, uq, dim=None, out=None):
    dim = dim if dim is not None else (g.shape[-1] if g.dim() >= 2 else 1)
    res, u = _segment_rows(g, mask, uq, dim)
    _np(out).reshape(-1, dim)[:u] = res.astype(F32)
    return out


def adam_begin_step(hyper):
    h = _np(hyper)
    h[4] = F32(h[4] * h[1])
    h[5] = F32(h[5] * h[2])
    h[6] = F32(h[0] * np.sqrt(F32(1) - h[5]) / (F32(1) - h[4]))


def sparse_lazy_adam(w, m, v, hyper, g, mask, uq, n_valid=None):
    dim = w.shape[1]
    gsum, u = _segment_rows(g, mask, uq, dim, n_valid)
    h = _np(hyper)
    st = types.SimpleNamespace(beta1=float(h[1]), beta2=float(h[2]), eps=float(h[3]), lr_t=h[6], grad_scale=h[7])
    R.lazy_adam_sparse(_np(w), _np(m), _np(v), _np(uq.uniq)[:u], gsum, st)


def sparse_ftrl(w, accum, linear, hyper, g, mask, uq, n_valid=None):
    dim = w.shape[1]
    gsum, u = _segment_rows(g, mask, uq, dim, n_valid)
    h = _np(hyper)
    st = types.SimpleNamespace(lr=float(h[0]), l1=float(h[1]), l2=float(h[2]), lr_power=float(h[3]), grad_scale=h[4])
    R.ftrl_sparse(_np(w), _np(accum), _np(linear), _np(uq.uniq)[:u], gsum, st)


# ---- hash-sharded variant (PeerHashRank) -------------------------------------------------------------------------
def shard_remap_hash(keys, owners_like, bits_like, out=None):
    """Any owner function works for the protocol; the stand-in uses key mod G (the CUDA kernel uses high mix64 bits)."""
    g, bits = owners_like.shape[0], bits_like.shape[0]
    k = _np(keys).astype(np.int64).reshape(-1)
    ok = (k >= 0) & ((k >> bits) == 0)
    res = np.where(ok, ((k % g) << bits) | k, g << bits)
    if out is None:
        out = torch.empty(keys.shape, dtype=torch.int64)
    out.copy_(torch.from_numpy(res).view(out.shape))
    return out


def fill_tail(buf, n_valid, value):
    _np(buf)[max(0, int(n_valid[0])):] = _np(value)[0]
    return buf


def gather(table, ids, out=None, oob_flag=None):
    t, i = _np(table), _np(ids).reshape(-1).astype(np.int64)
    ok = (i >= 0) & (i < t.shape[0])
    res = np.where(ok[:, None], t[np.where(ok, i, 0)], 0.0).astype(F32)
    if out is None:
        out = torch.empty(tuple(ids.shape) + (t.shape[1],), dtype=torch.float32)
    out.copy_(torch.from_numpy(res).view(out.shape))
    return out


class FakeMapParameter:
    """Slot table with the surface PeerHashRank uses: slots handed out in arrival order, rows initialised by a
    deterministic function of the KEY (like the CUDA table's Philox-by-key), arenas indexed by slot."""

    def __init__(self, key_dtype=torch.int64, value_shape=1, default_value="normal", permit_filter_value=1,
                 evict_filter_value=None, capacity=1 << 10, device="cpu", seed=0, **_):
        self.dim = value_shape if isinstance(value_shape, int) else value_shape[0]
        self.capacity = capacity
        self.values = torch.zeros((capacity + 1, self.dim), dtype=torch.float32)
        self.slot_of = {}
        self._arenas = []
        self.seed = seed
        self.overflowed = False

    def add_arena(self, fill=0.0, dim=None):
        a = torch.full((self.capacity + 1, dim or self.dim), float(fill), dtype=torch.float32)
        self._arenas.append(a)
        return a

    def _init_row(self, key):
        rng = np.random.default_rng([self.seed, int(key) & 0xffffffff, int(key) >> 32])
        return (rng.standard_normal(self.dim) * 0.01).astype(F32)

    def lookup_slots(self, key, insert_default_value=True):
        flat = _np(key).reshape(-1).astype(np.int64)
        slots = np.full(flat.size, self.capacity, dtype=np.int32)
        for i, k in enumerate(flat.tolist()):
            if k < 0:                                   # reserved keys (the blanked inbox tail) read the default row
                continue
            s = self.slot_of.get(k)
            if s is None and insert_default_value:
                if len(self.slot_of) >= self.capacity:
                    self.overflowed = True
                    continue
                s = self.slot_of[k] = len(self.slot_of)
                self.values[s] = torch.from_numpy(self._init_row(k))
                for a in self._arenas:
                    a[s] = a[self.capacity]
            if s is not None:
                slots[i] = s
        return torch.from_numpy(slots)

    def get_data(self):
        ks = sorted(self.slot_of)
        k = torch.tensor(ks, dtype=torch.int64)
        v = self.values[[self.slot_of[x] for x in ks]] if ks else torch.zeros((0, self.dim))
        return k, v

    def __len__(self):
        return len(self.slot_of)


def zero_row(buf, idx):
    r = int(idx[0])
    if 0 <= r < buf.shape[0]:
        buf[r].zero_()
