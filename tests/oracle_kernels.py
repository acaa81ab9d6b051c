"""CPU stand-in for mindrec_b200.ops backed by the numpy oracle — TEST ONLY.

The multi-process (gloo, world_size 2) tests exercise the HOST logic of mindrec_b200.sharded — key remapping,
bucket bounds, split sizes, the three all-to-alls, owner-side dedup — on machines without a GPU.  The product
code never imports this module; production always binds mindrec_b200.ops (CUDA, no fallback)."""
import numpy as np
import torch

from oracle import ref_numpy as R


class _Uq:
    pass


class UniqueResult:
    """Shape-compatible with ops.UniqueResult as far as sharded.py uses it (a capacity and a slice)."""

    def __init__(self, n, dtype=None, device=None, packed=False):
        self.n = n

    def sliced(self, n):
        return UniqueResult(n)


def _np(t):
    return t.detach().cpu().numpy()


def adam_hyper(lr, beta1=0.9, beta2=0.999, eps=1e-8, loss_scale=1.0, l2=0.0, device="cpu"):
    return torch.tensor([lr, beta1, beta2, eps, 1.0, 1.0, 0.0, 1.0 / loss_scale, l2] + [0.0] * 7, dtype=torch.float32)


def ftrl_hyper(lr, l1=0.0, l2=0.0, lr_power=-0.5, loss_scale=1.0, device="cpu"):
    return torch.tensor([lr, l1, l2, lr_power, 1.0 / loss_scale] + [0.0] * 11, dtype=torch.float32)


def unique(ids, table_like=None, result=None, ws_tag=None):
    assert result is None or result.n == ids.numel()
    flat = _np(ids).reshape(-1)
    bound = table_like.shape[0] if table_like is not None else None
    uniq, inverse, perm, seg_start = R.unique_sorted(flat, bound)
    n, u = flat.size, uniq.size
    uq = _Uq()
    uq.n = n
    uq.uniq = torch.zeros(n, dtype=ids.dtype); uq.uniq[:u] = torch.from_numpy(uniq.astype(flat.dtype))
    uq.inverse = torch.from_numpy(inverse)
    uq.count = torch.tensor([u], dtype=torch.int32)
    uq.perm = torch.from_numpy(perm)
    uq.seg_start = torch.zeros(n + 1, dtype=torch.int32); uq.seg_start[:u + 1] = torch.from_numpy(seg_start)
    uq.seg_of = torch.from_numpy(np.repeat(np.arange(u), np.diff(seg_start)).astype(np.int32))
    return uq


def shard_bounds(uniq, count, edges, out=None):
    u = int(count.item())
    return torch.from_numpy(np.searchsorted(_np(uniq)[:u], _np(edges), side="left").astype(np.int32))


def gather(table, ids, out=None, oob_flag=None):
    return torch.from_numpy(R.gather(_np(table), _np(ids)))


def gather_masked(table, ids, mask, out=None, oob_flag=None, out_dtype=torch.float32):
    res = torch.from_numpy(R.gather_masked(_np(table), _np(ids), _np(mask)))
    if out is not None:
        out.copy_(res)
        return out
    return res.to(out_dtype)


def gather_reduce(table, ids, mask, bias, out=None, oob_flag=None):
    res = torch.from_numpy(R.gather_reduce(_np(table), _np(ids), _np(mask), _np(bias)))
    if out is not None:
        out.copy_(res)
        return out
    return res


def segment_sum(g, mask, uq, dim=None, out=None):
    gn = _np(g).astype(np.float32)
    n = uq.n
    div = n // gn.reshape(-1, dim).shape[0]
    u = int(uq.count.item())
    s = R.segment_sum(gn.reshape(-1, dim), _np(uq.inverse), u, None if mask is None else _np(mask), div=div)
    res = torch.zeros((n, dim), dtype=torch.float32)
    res[:u] = torch.from_numpy(s.astype(np.float32))
    return res


def _state(h):
    st = R.AdamState(float(h[0]), float(h[1]), float(h[2]), float(h[3]))
    st.grad_scale = np.float32(h[7]); st.lr_t = np.float32(h[6])
    return st


def adam_begin_step(h):
    h[4] *= h[1]; h[5] *= h[2]
    h[6] = h[0] * torch.sqrt(1 - h[5]) / (1 - h[4])


def sparse_lazy_adam(w, m, v, hyper, g, mask, uq):
    u = int(uq.count.item())
    gs = segment_sum(g, mask, uq, dim=w.shape[1])[:u]
    wn, mn, vn = _np(w), _np(m), _np(v)
    R.lazy_adam_sparse(wn, mn, vn, _np(uq.uniq)[:u], _np(gs), _state(hyper))
    w.copy_(torch.from_numpy(wn)); m.copy_(torch.from_numpy(mn)); v.copy_(torch.from_numpy(vn))


def sparse_ftrl(w, acc, lin, hyper, g, mask, uq):
    u = int(uq.count.item())
    gs = segment_sum(g, mask, uq, dim=w.shape[1])[:u]
    st = R.FtrlState(float(hyper[0]), float(hyper[1]), float(hyper[2]), float(hyper[3]))
    st.grad_scale = np.float32(hyper[4])
    wn, an, ln = _np(w), _np(acc), _np(lin)
    R.ftrl_sparse(wn, an, ln, _np(uq.uniq)[:u], _np(gs), st)
    w.copy_(torch.from_numpy(wn)); acc.copy_(torch.from_numpy(an)); lin.copy_(torch.from_numpy(ln))


def adam_dense(w, m, v, hyper, g):
    wn, mn, vn = _np(w), _np(m), _np(v)
    R.adam_dense(wn, mn, vn, _np(g), _state(hyper))
    w.copy_(torch.from_numpy(wn)); m.copy_(torch.from_numpy(mn)); v.copy_(torch.from_numpy(vn))


def shard_remap(ids, table_like, owners_like, out=None):
    g, r = owners_like.shape[0], owners_like.shape[1]
    v = table_like.shape[0]
    ok = (ids >= 0) & (ids < v)
    km = (ids % g) * r + torch.div(ids, g, rounding_mode="floor")
    return torch.where(ok, km, torch.full_like(ids, g * r))


def sigmoid_xent(a, b, label, sens, out=None, half=False):
    x = a.double() + (b.double() if b is not None and b.numel() else 0)
    n = x.numel()
    loss = torch.from_numpy(np.array([R.sigmoid_xent(_np(x), _np(label)).mean()], dtype=np.float32))
    delta = (float(sens[0]) * (torch.sigmoid(x) - label.double()) / n).float()
    res = (x.float(), loss, delta, delta.half() if half else torch.empty(0, dtype=torch.float16), delta.sum().reshape(1))
    if out is not None:
        for o, r in zip(out, res):
            if o.numel():
                o.copy_(r.reshape(o.shape))
        return out
    return res


def dense_head_fwd(h, w, bias, out=None):
    return (h.double() @ w.double().view(-1, 1) + bias.double()).float()


def dense_head_bwd(delta, h, w, masked, gw, gb_head, gb_prev=None, out=None):
    gh = delta.double() @ w.double().view(1, -1)
    if masked:
        gh = gh * (h > 0)
    gw.copy_((delta.double().t() @ h.double()).view(-1).float())
    gb_head.copy_(delta.double().sum().reshape(gb_head.shape).float())
    if gb_prev is not None:
        gb_prev.copy_(gh.sum(0).float())
    return gh.to(h.dtype)


def relu_bwd_bias(g, y, gb, out=None):
    gz = g if y is None else g * (y > 0)
    gb.copy_(gz.double().sum(0).float())
    return gz
