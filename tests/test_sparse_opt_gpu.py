"""K3-K5 parity: deterministic segment-sum and fused sparse LazyAdam / FTRL vs the numpy oracle
(fp32, 1e-5 relative as BASELINE.json's north_star states)."""
import numpy as np
import pytest
import torch

from mindrec_b200 import ops
from oracle import ref_numpy as R

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-7


def _zipf_ids(rng, b, f, vocab, dense_fields=3):
    ids = (rng.zipf(1.05, size=(b, f)) % vocab).astype(np.int32)
    ids[:, :dense_fields] = np.arange(dense_fields)  # one id per dense field: B-long segments
    return ids


@pytest.mark.parametrize("dim", [80, 128, 16, 1, 27])
@pytest.mark.parametrize("b", [5, 700])
def test_segment_sum(cuda, dim, b):
    rng = np.random.default_rng(dim * 1000 + b)
    f, vocab = 39, 5000
    ids = _zipf_ids(rng, b, f, vocab)
    g = rng.standard_normal((b * f, dim)).astype(np.float32)
    mask = rng.random(b * f).astype(np.float32)
    uq = ops.unique(torch.from_numpy(ids).to(cuda))
    out = ops.segment_sum(torch.from_numpy(g).to(cuda), torch.from_numpy(mask).to(cuda), uq, dim=dim)
    uniq, inverse, _, _ = R.unique_sorted(ids)
    ref = R.segment_sum(g, inverse, uniq.size, mask)
    scale = np.abs(ref).max()
    np.testing.assert_allclose(out[:uniq.size].cpu().numpy(), ref, rtol=RTOL, atol=RTOL * scale)
    # conservation: sum of segment sums == sum of masked rows
    np.testing.assert_allclose(out[:uniq.size].double().sum(0).cpu().numpy(),
                               (g.astype(np.float64) * mask[:, None]).sum(0), rtol=1e-4, atol=1e-3)


def test_segment_sum_is_run_to_run_deterministic(cuda):
    rng = np.random.default_rng(0)
    ids = _zipf_ids(rng, 3000, 39, 20000)
    g = torch.from_numpy(rng.standard_normal((3000 * 39, 80)).astype(np.float32)).to(cuda)
    uq = ops.unique(torch.from_numpy(ids).to(cuda))
    a = ops.segment_sum(g, None, uq).clone()
    b = ops.segment_sum(g, None, uq).clone()
    assert torch.equal(a, b)


def test_segment_sum_broadcast_rows_div(cuda):
    """wide path: one logit gradient per sample, broadcast over its F lookups (div = F)."""
    rng = np.random.default_rng(1)
    b, f, vocab = 900, 39, 3000
    ids = _zipf_ids(rng, b, f, vocab)
    g = rng.standard_normal((b, 1)).astype(np.float32)
    mask = rng.random(b * f).astype(np.float32)
    uq = ops.unique(torch.from_numpy(ids).to(cuda))
    out = ops.segment_sum(torch.from_numpy(g).to(cuda), torch.from_numpy(mask).to(cuda), uq, dim=1)
    uniq, inverse, _, _ = R.unique_sorted(ids)
    ref = R.segment_sum(g, inverse, uniq.size, mask, div=f)
    np.testing.assert_allclose(out[:uniq.size].cpu().numpy(), ref, rtol=RTOL, atol=RTOL * np.abs(ref).max())


@pytest.mark.parametrize("dim", [80, 128, 1, 6])
def test_sparse_lazy_adam_three_steps(cuda, dim):
    """Two oracles run side by side: `tight` is fed the kernel's own fp32 segment sums (mrec_segment_sum
    shares the fused kernel's summation order), so it isolates the LazyAdam row math at 1e-5; `e2e` sums
    in float64 and is compared at a bound that allows for fp32 cancellation in the 600-term hot segments
    (|err| <= 1e-5 * sum|terms| can exceed 1e-5 * |sum| there)."""
    rng = np.random.default_rng(dim)
    vocab, b, f = 4000, 600, 39
    w = (rng.standard_normal((vocab, dim)) * 0.01).astype(np.float32)
    m = np.zeros_like(w)
    v = np.zeros_like(w)
    tw, tm, tv = w.copy(), m.copy(), v.copy()
    dw, dm, dv = (torch.from_numpy(x.copy()).to(cuda) for x in (w, m, v))
    st = R.AdamState(3.5e-4, eps=1e-8, loss_scale=1024.0)
    hyper = ops.adam_hyper(3.5e-4, eps=1e-8, loss_scale=1024.0, device=cuda)
    for step in range(3):
        ids = _zipf_ids(rng, b, f, vocab + 50)  # some ids out of range: must be skipped
        g = (rng.standard_normal((b * f, dim)) * 1024).astype(np.float32)
        mask = rng.random(b * f).astype(np.float32)
        dg, dmask = torch.from_numpy(g).to(cuda), torch.from_numpy(mask).to(cuda)
        uq = ops.unique(torch.from_numpy(ids).to(cuda), table_like=dw)
        gsum_gpu = ops.segment_sum(dg, dmask, uq, dim=dim).cpu().numpy()
        ops.adam_begin_step(hyper)
        ops.sparse_lazy_adam(dw, dm, dv, hyper, dg, dmask, uq)
        uniq, inverse, _, _ = R.unique_sorted(ids, bound=vocab)
        st.begin_step()
        R.lazy_adam_sparse(w, m, v, uniq, R.segment_sum(g, inverse, uniq.size, mask), st)
        R.lazy_adam_sparse(tw, tm, tv, uniq, gsum_gpu[:uniq.size], st)
    np.testing.assert_allclose(hyper[4:7].cpu().numpy(),
                               [st.beta1_power, st.beta2_power, st.lr_t], rtol=1e-6)
    for got, ref in ((dw, tw), (dm, tm), (dv, tv)):
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=RTOL, atol=RTOL * np.abs(ref).max())
    for got, ref in ((dw, w), (dm, m), (dv, v)):
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=1e-4, atol=1e-4 * np.abs(ref).max())


@pytest.mark.parametrize("dim", [1, 64])
def test_sparse_ftrl_three_steps(cuda, dim):
    """As for LazyAdam: `tight` feeds the oracle's row math the kernel's own fp32 segment sums and holds FTRL to
    1e-5; `e2e` sums in float64 and carries the fp32 cancellation of the 600-term hot segments."""
    rng = np.random.default_rng(100 + dim)
    vocab, b, f = 4000, 600, 39
    w = (rng.standard_normal((vocab, dim)) * 0.01).astype(np.float32)
    acc = np.full_like(w, 1.0)
    lin = np.zeros_like(w)
    tw, ta, tl = w.copy(), acc.copy(), lin.copy()
    dw, da, dl = (torch.from_numpy(x.copy()).to(cuda) for x in (w, acc, lin))
    st = R.FtrlState(5e-2, l1=1e-8, l2=1e-8, loss_scale=1024.0)
    hyper = ops.ftrl_hyper(5e-2, l1=1e-8, l2=1e-8, loss_scale=1024.0, device=cuda)
    for step in range(3):
        ids = _zipf_ids(rng, b, f, vocab)
        if dim == 1:
            g = (rng.standard_normal((b, 1)) * 1024).astype(np.float32)
            div = f
        else:
            g = (rng.standard_normal((b * f, dim)) * 1024).astype(np.float32)
            div = 1
        mask = rng.random(b * f).astype(np.float32)
        dg, dmask = torch.from_numpy(g).to(cuda), torch.from_numpy(mask).to(cuda)
        uq = ops.unique(torch.from_numpy(ids).to(cuda), table_like=dw)
        gsum_gpu = ops.segment_sum(dg, dmask, uq, dim=dim).cpu().numpy()
        ops.sparse_ftrl(dw, da, dl, hyper, dg, dmask, uq)
        uniq, inverse, _, _ = R.unique_sorted(ids, bound=vocab)
        R.ftrl_sparse(w, acc, lin, uniq, R.segment_sum(g, inverse, uniq.size, mask, div=div), st)
        R.ftrl_sparse(tw, ta, tl, uniq, gsum_gpu[:uniq.size], st)
    # 1e-5 relative (north_star) with one stated exception: `lin + g - sigma * w` is a subtraction of fp32 numbers
    # that cancels on a few rows in a thousand (|lin_new| << |lin| + |g|); ANY fp32 ApplyFtrl, the reference's
    # included, is then off by ~1e-7 of the operands, which is more than 1e-5 of the small result.  Those rows are
    # held to 1e-6 of the tensor's scale instead (fp32 emulation of the formula on the host: 2e-7).
    for got, ref in ((dw, tw), (da, ta), (dl, tl)):
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=RTOL, atol=1e-6 * np.abs(ref).max())
    for got, ref in ((dw, w), (da, acc), (dl, lin)):
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=2e-5, atol=2e-5 * np.abs(ref).max())


def test_lazy_adam_all_rows_touched_equals_dense_adam(cuda):
    """Property: LazyAdam with every row touched once == dense Adam."""
    rng = np.random.default_rng(9)
    vocab, dim = 3000, 80
    w0 = torch.from_numpy((rng.standard_normal((vocab, dim)) * 0.01).astype(np.float32)).to(cuda)
    g = torch.from_numpy(rng.standard_normal((vocab, dim)).astype(np.float32)).to(cuda)
    ids = torch.randperm(vocab, device=cuda).to(torch.int32)
    ws, ms, vs = w0.clone(), torch.zeros_like(w0), torch.zeros_like(w0)
    wd, md, vd = w0.clone(), torch.zeros_like(w0), torch.zeros_like(w0)
    hs = ops.adam_hyper(1e-3, device=cuda)
    hd = ops.adam_hyper(1e-3, device=cuda)
    uq = ops.unique(ids, table_like=ws)
    ops.adam_begin_step(hs)
    ops.sparse_lazy_adam(ws, ms, vs, hs, g, None, uq)
    g_dense = torch.empty_like(g)
    g_dense[ids.long()] = g
    ops.adam_begin_step(hd)
    ops.adam_dense(wd, md, vd, hd, g_dense)
    assert torch.equal(ws, wd) and torch.equal(ms, md) and torch.equal(vs, vd)


def test_dense_adam_and_ftrl_match_oracle(cuda):
    rng = np.random.default_rng(11)
    n = 100003
    w = (rng.standard_normal(n) * 0.1).astype(np.float32)
    g = rng.standard_normal(n).astype(np.float32)
    m, v = np.zeros_like(w), np.zeros_like(w)
    dw, dm, dv = (torch.from_numpy(x.copy()).to(cuda) for x in (w, m, v))
    st = R.AdamState(5e-4, eps=5e-8, loss_scale=1000.0)
    hyper = ops.adam_hyper(5e-4, eps=5e-8, loss_scale=1000.0, device=cuda)
    for _ in range(2):
        ops.adam_begin_step(hyper)
        ops.adam_dense(dw, dm, dv, hyper, torch.from_numpy(g).to(cuda))
        st.begin_step()
        R.adam_dense(w, m, v, g, st)
    np.testing.assert_allclose(dw.cpu().numpy(), w, rtol=RTOL, atol=1e-7)
    w2 = (rng.standard_normal(n) * 0.1).astype(np.float32)
    acc, lin = np.full_like(w2, 0.1), np.zeros_like(w2)
    dw2, da, dl = (torch.from_numpy(x.copy()).to(cuda) for x in (w2, acc, lin))
    fst = R.FtrlState(0.1, l1=5e-4, l2=5e-4)
    fh = ops.ftrl_hyper(0.1, l1=5e-4, l2=5e-4, device=cuda)
    ops.ftrl_dense(dw2, da, dl, fh, torch.from_numpy(g).to(cuda))
    R.ftrl_dense(w2, acc, lin, g, fst)
    np.testing.assert_allclose(dw2.cpu().numpy(), w2, rtol=RTOL, atol=1e-6 * np.abs(w2).max())   # see the sparse FTRL test
    np.testing.assert_allclose(da.cpu().numpy(), acc, rtol=RTOL)


@pytest.mark.parametrize("dim,b,use_mask", [(80, 300, True), (6, 300, True), (80, 64, False), (16, 1, True), (8, 937, True)])
def test_fp16_gradient_rows_equal_cast_then_fp32(cuda, dim, b, use_mask):
    """fp16 gradient rows fuse the Cast-to-fp32 of the DenseLayer bprop: same bits as casting first
    (D % 8 == 0 takes the 16-byte-chunk tile walk, other widths the generic one)."""
    rng = np.random.default_rng(21 + dim)
    vocab, f = 3000, 39
    ids = torch.from_numpy(_zipf_ids(rng, b, f, vocab)).to(cuda)
    g16 = torch.from_numpy(rng.standard_normal((b * f, dim)).astype(np.float16)).to(cuda)
    mask = torch.from_numpy(rng.random(b * f).astype(np.float32)).to(cuda) if use_mask else None
    w0 = torch.from_numpy((rng.standard_normal((vocab, dim)) * 0.01).astype(np.float32)).to(cuda)
    uq = ops.unique(ids, table_like=w0)
    a = ops.segment_sum(g16, mask, uq, dim=dim).clone()
    b32 = ops.segment_sum(g16.float(), mask, uq, dim=dim).clone()
    u = int(uq.count.item())
    assert torch.equal(a[:u], b32[:u])
    res = []
    for g in (g16, g16.float()):
        w, m, v = w0.clone(), torch.zeros_like(w0), torch.zeros_like(w0)
        hyper = ops.adam_hyper(1e-3, device=cuda)
        ops.adam_begin_step(hyper)
        ops.sparse_lazy_adam(w, m, v, hyper, g, mask, uq)
        res.append(w)
    assert torch.equal(res[0], res[1])


@pytest.mark.parametrize("dim", [80, 128, 12])
@pytest.mark.parametrize("half", [False, True])
def test_interleaved_records_equal_split_arrays(cuda, dim, half):
    """The interleaved layout wmv[V,3,D] (one w | m | v record per row) is a storage choice: two LazyAdam steps leave
    bit-identical weights and moments, and the gathers read the same rows, as with three [V,D] arrays."""
    rng = np.random.default_rng(50 + dim)
    vocab, b, f = 5000, 700, 39
    w0 = torch.from_numpy((rng.standard_normal((vocab, dim)) * 0.01).astype(np.float32)).to(cuda)
    w, m, v = w0.clone(), torch.zeros_like(w0), torch.zeros_like(w0)
    wmv = torch.zeros((vocab, 3, dim), device=cuda)
    wmv[:, 0, :] = w0
    h1, h2 = ops.adam_hyper(1e-3, loss_scale=8.0, device=cuda), ops.adam_hyper(1e-3, loss_scale=8.0, device=cuda)
    for step in range(2):
        ids = torch.from_numpy(_zipf_ids(rng, b, f, vocab + 20)).to(cuda)
        g = torch.from_numpy(rng.standard_normal((b * f, dim)).astype(np.float32)).to(cuda)
        g = g.half() if half else g
        mask = torch.from_numpy(rng.random(b * f).astype(np.float32)).to(cuda)
        uq = ops.unique(ids, table_like=w)
        ops.adam_begin_step(h1)
        ops.sparse_lazy_adam(w, m, v, h1, g, mask, uq)
        ops.adam_begin_step(h2)
        ops.sparse_lazy_adam(wmv, None, None, h2, g, mask, ops.unique(ids, table_like=wmv))
        assert torch.equal(wmv[:, 0, :], w) and torch.equal(wmv[:, 1, :], m) and torch.equal(wmv[:, 2, :], v)
        wts = mask.view(b, f)
        assert torch.equal(ops.gather(wmv, ids), ops.gather(w, ids))
        assert torch.equal(ops.gather_masked(wmv, ids, wts), ops.gather_masked(w, ids, wts))
        assert torch.equal(ops.gather_masked(wmv, ids, wts, out_dtype=torch.float16),
                           ops.gather_masked(w, ids, wts, out_dtype=torch.float16))
