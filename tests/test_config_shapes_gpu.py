"""Parity at the BENCHMARKED shapes ("weak" 2): the ops bench.py times are checked here on the
batches bench.py feeds them.

Config 2 (BASELINE.json configs[1]): 16000 x 39 ids from CriteoSynth (Criteo-Kaggle cardinalities, Zipf 1.05) — the
13 dense fields give thirteen 16000-long segments (the long-chain CTAs of rows_update_kernel), U ~ 120 k unique
rows, D = 80 (256-bit row pass) and D = 128, fp16 and fp32 gradient rows.  The table is folded to 1 M rows so that
the numpy oracle's copy fits in host memory; the fold keeps the dense ids and the Zipf head.

Config 5: 425 984 int64 Zipf keys over 2^40, D = 128, permit_filter_value = 2, evict_filter_value = 3, six steps
against the dict model of MapParameter.

Tolerances are north_star's: bit-exact indices / keys / gathered rows, 1e-5 relative for sums and optimizer rows.
"""
import numpy as np
import pytest
import torch

from mindrec_b200 import _lib, hash as H, ops, synth
from oracle import ref_numpy as R

pytestmark = pytest.mark.gpu
RTOL = 1e-5
B, F = 16000, 39
V_FOLD = 1_000_003


def _config2_batch(seed):
    gen = synth.CriteoSynth(B, cards=synth.CARD_KAGGLE, alpha=1.05, seed=20260101 + seed)
    ids, wts, _ = gen.next()
    ids = ids.copy()
    ids[:, synth.N_DENSE:] = synth.N_DENSE + ids[:, synth.N_DENSE:] % (V_FOLD - synth.N_DENSE)
    return ids, wts


def _grads(rng, n, dim, half, signed):
    g = rng.standard_normal((n, dim)).astype(np.float32)
    if not signed:
        g = np.abs(g) + 0.5      # sums that do not cancel: "1e-5 relative" is then meaningful end to end
    if half:
        g = g.astype(np.float16).astype(np.float32)     # exactly representable in both dtypes
    return g


def _dev(x, cuda, half=False):
    t = torch.from_numpy(np.ascontiguousarray(x)).to(cuda)
    return t.half() if half else t


@pytest.mark.parametrize("half", [True, False], ids=["g16", "g32"])
@pytest.mark.parametrize("dim", [80, 128])
def test_config2_segment_sum(cuda, dim, half):
    rng = np.random.default_rng(dim + half)
    ids, wts = _config2_batch(0)
    n = ids.size
    uniq, inverse, perm, seg_start = R.unique_sorted(ids, bound=V_FOLD)
    assert np.diff(seg_start).max() == B            # the dense fields: 16000-long segments (long-chain path)
    assert 100_000 < uniq.size < 140_000
    uq = ops.unique(_dev(ids, cuda), table_like=torch.empty((V_FOLD, 0), device=cuda))
    assert int(uq.count.item()) == uniq.size
    np.testing.assert_array_equal(uq.perm.cpu().numpy(), perm)
    mask = wts.reshape(-1)
    for signed in (False, True):
        g = _grads(rng, n, dim, half, signed)
        out = ops.segment_sum(_dev(g, cuda, half), _dev(mask, cuda), uq, dim=dim)[:uniq.size].cpu().numpy()
        ref = R.segment_sum(g, inverse, uniq.size, mask)
        if signed:
            # a sum of up to 16000 signed terms can cancel: the bound is 1e-5 of the sum of magnitudes
            mag = R.segment_sum(np.abs(g), inverse, uniq.size, mask)
            assert np.all(np.abs(out - ref) <= RTOL * mag + 1e-30)
        else:
            np.testing.assert_allclose(out, ref, rtol=RTOL, atol=0)


@pytest.mark.parametrize("half", [True, False], ids=["g16", "g32"])
@pytest.mark.parametrize("dim", [80, 128])
def test_config2_sparse_lazy_adam(cuda, dim, half):
    """Two steps of mrec_sparse_lazy_adam (segsum_stage_kernel + rows_update_kernel: exactly 2 launches per call) on
    the config-2 batches.  `e2e`: non-cancelling gradients, float64-summed oracle, 1e-5 relative on w, m, v.
    `tight`: signed gradients, the oracle's row math fed the kernel's own fp32 segment sums (mrec_segment_sum shares
    the fused kernel's summation order), 1e-5 relative."""
    rng = np.random.default_rng(7 * dim + half)
    w0 = (rng.standard_normal((V_FOLD, dim)) * 0.01).astype(np.float32)
    table_like = torch.empty((V_FOLD, 0), device=cuda)
    for signed in (False, True):
        w, m, v = w0.copy(), np.zeros_like(w0), np.zeros_like(w0)
        dw, dm, dv = _dev(w, cuda), _dev(m, cuda), _dev(v, cuda)
        st = R.AdamState(3.5e-4, eps=1e-8, loss_scale=1024.0)
        hyper = ops.adam_hyper(3.5e-4, eps=1e-8, loss_scale=1024.0, device=cuda)
        for step in range(2):
            ids, wts = _config2_batch(10 + step)
            mask = wts.reshape(-1)
            g = _grads(rng, ids.size, dim, half, signed) * np.float32(1024.0)
            dg, dmask = _dev(g, cuda, half), _dev(mask, cuda)
            uq = ops.unique(_dev(ids, cuda), table_like=table_like)
            uniq, inverse, _, _ = R.unique_sorted(ids, bound=V_FOLD)
            if signed:
                gsum = ops.segment_sum(dg, dmask, uq, dim=dim)[:uniq.size].cpu().numpy()
            else:
                gsum = R.segment_sum(g, inverse, uniq.size, mask)
            ops.adam_begin_step(hyper)
            n0 = _lib.launch_count()
            ops.sparse_lazy_adam(dw, dm, dv, hyper, dg, dmask, uq)
            assert _lib.launch_count() - n0 == 2
            st.begin_step()
            R.lazy_adam_sparse(w, m, v, uniq, gsum, st)
        # elementwise 1e-5; `b1 * m + (1 - b1) * g` with signed g cancels on a few elements, which any fp32 kernel
        # resolves only to ~1e-7 of the operands: those are held to 1e-6 of the tensor's scale
        for got, ref in ((dw, w), (dm, m), (dv, v)):
            np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=RTOL, atol=1e-6 * np.abs(ref).max())


def test_config2_sparse_ftrl_wide(cuda):
    """FTRL on the dim-1 wide table with the logit gradient broadcast over the 39 lookups of a sample (div = F)."""
    rng = np.random.default_rng(3)
    w = (rng.standard_normal((V_FOLD, 1)) * 0.01).astype(np.float32)
    acc, lin = np.ones_like(w), np.zeros_like(w)
    dw, da, dl = _dev(w, cuda), _dev(acc, cuda), _dev(lin, cuda)
    st = R.FtrlState(5e-2, l1=1e-8, l2=1e-8, loss_scale=1024.0)
    hyper = ops.ftrl_hyper(5e-2, l1=1e-8, l2=1e-8, loss_scale=1024.0, device=cuda)
    table_like = torch.empty((V_FOLD, 0), device=cuda)
    for step in range(2):
        ids, wts = _config2_batch(20 + step)
        mask = wts.reshape(-1)
        delta = ((np.abs(rng.standard_normal((B, 1))) + 0.5) * 1024.0 / B).astype(np.float32)
        uq = ops.unique(_dev(ids, cuda), table_like=table_like)
        uniq, inverse, _, _ = R.unique_sorted(ids, bound=V_FOLD)
        ops.sparse_ftrl(dw, da, dl, hyper, _dev(delta, cuda), _dev(mask, cuda), uq)
        R.ftrl_sparse(w, acc, lin, uniq, R.segment_sum(delta, inverse, uniq.size, mask, div=F), st)
    # 1e-5 relative; rows whose `lin + g - sigma * w` cancels in fp32 are held to 1e-6 of the tensor's scale (see
    # tests/test_sparse_opt_gpu.py::test_sparse_ftrl_three_steps)
    for got, ref in ((dw, w), (da, acc), (dl, lin)):
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=RTOL, atol=1e-6 * np.abs(ref).max())


def _row_of_key(keys, dim):
    """A deterministic fp32 row per key (computed on the host for both sides)."""
    base = ((keys % 1000003).astype(np.float32) * np.float32(1e-6))[:, None]
    return (base + np.arange(dim, dtype=np.float32)[None, :] * np.float32(1e-3)).astype(np.float32)


def test_config5_hash_permit_evict(cuda):
    """MapParameter at the config-5 shape: 16384 x 26 int64 Zipf keys over 2^40 per step, D = 128, admission on the
    second sighting, eviction after 3 unseen steps; every step's returned rows and the resident key set are exact
    against the dict model.  After each lookup the rows of the resident looked-up keys are overwritten by slot with
    a per-key pattern (what an optimizer does), so later lookups return distinguishable rows."""
    dim, n = 128, 16384 * 26
    rng = np.random.default_rng(5)
    mp = H.MapParameter(key_dtype=torch.int64, value_shape=dim, default_value="zeros", permit_filter_value=2,
                        evict_filter_value=3, capacity=1 << 22, device=cuda)
    model = R.MapParameterModel(dim, default_value=0.0, permit_filter_value=2, evict_filter_value=3)
    for it in range(6):
        keys = ((rng.zipf(1.05, size=n) - 1) % (1 << 40)).astype(np.int64)
        dkeys = _dev(keys, cuda)
        slots = mp.lookup_slots(dkeys)
        got = ops.gather(mp.values, slots).cpu().numpy()
        np.testing.assert_array_equal(got, model.get(keys))
        rows = _row_of_key(keys, dim)
        _lib.aot_call("mrec_hash_scatter_rows", [mp.values, slots, _dev(rows, cuda), ops._dummy(cuda)])
        uk, first = np.unique(keys, return_index=True)
        for k, i in zip(uk.tolist(), first.tolist()):
            if k in model.rows:
                model.rows[k] = rows[i]
        if it % 2 == 1:
            mp.evict()
            model.evict()
        np.testing.assert_array_equal(mp.get_keys().cpu().numpy(), model.keys())
        assert len(mp) == model.keys().size
    assert torch.equal(mp.values[mp.capacity], torch.zeros(dim, device=cuda))       # default row untouched
    st = mp.state.tolist()
    assert st[3] == 0                                   # no overflow
    k, v = mp.get_data()
    ref_rows = np.stack([model.rows[int(x)] for x in k.cpu().numpy()])
    np.testing.assert_array_equal(v.cpu().numpy(), ref_rows)
