"""CPU tests of the oracle itself (no GPU): the numpy restatement is pinned by algebraic identities,
finite differences and the frozen vectors in tests/golden/; the C restatement must agree with it."""
import os

import numpy as np
import pytest

from oracle import ref_c as C
from oracle import ref_numpy as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_unique_docs_example_both_orders():
    ids = np.array([1, 2, 5, 2], dtype=np.int32)
    uniq, inv, perm, seg = R.unique_sorted(ids)
    assert uniq.tolist() == [1, 2, 5] and inv.tolist() == [0, 1, 2, 1]
    assert perm.tolist() == [0, 1, 3, 2] and seg.tolist() == [0, 1, 3, 4]
    u2, i2 = R.unique_first(ids)
    assert u2.tolist() == [1, 2, 5] and i2.tolist() == [0, 1, 2, 1]
    u3, i3 = R.unique_first(np.array([7, 3, 7, 1, 3], dtype=np.int64))
    assert u3.tolist() == [7, 3, 1] and i3.tolist() == [0, 1, 0, 2, 1]


def test_unique_invariants_random():
    rng = np.random.default_rng(0)
    ids = rng.integers(0, 300, size=5000).astype(np.int32)
    uniq, inv, perm, seg = R.unique_sorted(ids)
    assert (uniq[inv] == ids).all() and (np.diff(uniq) > 0).all()
    assert sorted(set(ids.tolist())) == uniq.tolist()
    assert (ids[perm][seg[:-1]] == uniq).all()


def test_gather_out_of_range_is_zero_row():
    tab = np.arange(12, dtype=np.float32).reshape(4, 3)
    out = R.gather(tab, np.array([[0, 4], [-1, 3]]))
    assert out.shape == (2, 2, 3)
    assert (out[0, 0] == tab[0]).all() and not out[0, 1].any() and not out[1, 0].any()


def test_segment_sum_conservation():
    rng = np.random.default_rng(1)
    g = rng.standard_normal((1000, 8)).astype(np.float32)
    ids = rng.integers(0, 50, 1000)
    uniq, inv, _, _ = R.unique_sorted(ids)
    s = R.segment_sum(g, inv, uniq.size)
    np.testing.assert_allclose(s.sum(0), g.astype(np.float64).sum(0), rtol=1e-12)


def test_fm_identity_against_pairwise_sum():
    vx = np.random.default_rng(2).standard_normal((6, 39, 16))
    np.testing.assert_allclose(R.fm_forward(vx), R.fm_pairwise(vx), rtol=1e-10)


def test_fm_backward_finite_difference():
    rng = np.random.default_rng(3)
    vx = rng.standard_normal((3, 5, 4))
    g = rng.standard_normal((3, 1))
    ana = R.fm_backward(vx, g)
    eps = 1e-6
    for idx in [(0, 0, 0), (1, 3, 2), (2, 4, 3)]:
        p, m = vx.copy(), vx.copy()
        p[idx] += eps
        m[idx] -= eps
        num = ((R.fm_forward(p) - R.fm_forward(m)) * g).sum() / (2 * eps)
        np.testing.assert_allclose(ana[idx], num, rtol=1e-5)


def test_cross_backward_finite_difference():
    rng = np.random.default_rng(4)
    x0 = rng.standard_normal((4, 10))
    w = rng.standard_normal((6, 10)) * 0.3
    b = rng.standard_normal((6, 10)) * 0.1
    gy = rng.standard_normal((4, 10))
    dx, dw, db = R.cross_backward(x0, w, b, gy)
    eps = 1e-6

    def f(x0_, w_, b_):
        return (R.cross_forward(x0_, w_, b_)[0] * gy).sum()
    for (i, j) in [(0, 0), (3, 9), (2, 5)]:
        p, m = x0.copy(), x0.copy(); p[i, j] += eps; m[i, j] -= eps
        np.testing.assert_allclose(dx[i, j], (f(p, w, b) - f(m, w, b)) / (2 * eps), rtol=1e-4)
    for (l, j) in [(0, 0), (5, 9), (2, 4)]:
        p, m = w.copy(), w.copy(); p[l, j] += eps; m[l, j] -= eps
        np.testing.assert_allclose(dw[l, j], (f(x0, p, b) - f(x0, m, b)) / (2 * eps), rtol=1e-4)
        p, m = b.copy(), b.copy(); p[l, j] += eps; m[l, j] -= eps
        np.testing.assert_allclose(db[l, j], (f(x0, w, p) - f(x0, w, m)) / (2 * eps), rtol=1e-4)


def test_ftrl_closed_form_without_regularisation():
    """l1 = l2 = 0, w0 = 0, one step: w = -g * lr / sqrt(a0 + g^2) * ... (closed form of B7)."""
    g = np.array([[0.5], [-2.0]], dtype=np.float32)
    w = np.zeros((2, 1), np.float32); acc = np.full((2, 1), 0.1, np.float32); lin = np.zeros((2, 1), np.float32)
    st = R.FtrlState(0.05)
    R.ftrl_dense(w, acc, lin, g, st)
    expect = -g / (np.sqrt(0.1 + g * g) / 0.05)
    np.testing.assert_allclose(w, expect, rtol=1e-6)
    np.testing.assert_allclose(acc, 0.1 + g * g, rtol=1e-6)


def test_adam_first_step_moves_by_lr():
    """Step 1 of Adam: m/(sqrt(v)) = sign(g) * (1-b1)/sqrt(1-b2); with bias correction |dw| ~= lr."""
    w = np.zeros((1, 4), np.float32); m = np.zeros_like(w); v = np.zeros_like(w)
    st = R.AdamState(1e-3)
    st.begin_step()
    R.adam_dense(w, m, v, np.array([[1.0, -3.0, 10.0, -0.1]]), st)
    np.testing.assert_allclose(np.abs(w), 1e-3, rtol=1e-4)


def test_lazy_adam_leaves_other_rows_alone_and_matches_dense_when_all_touched():
    rng = np.random.default_rng(5)
    w0 = rng.standard_normal((20, 4)).astype(np.float32)
    g = rng.standard_normal((20, 4)).astype(np.float32)
    st = R.AdamState(1e-2); st.begin_step()
    w1, m1, v1 = w0.copy(), np.zeros_like(w0), np.zeros_like(w0)
    R.lazy_adam_sparse(w1, m1, v1, np.arange(20), g, st)
    w2, m2, v2 = w0.copy(), np.zeros_like(w0), np.zeros_like(w0)
    R.adam_dense(w2, m2, v2, g, st)
    assert (w1 == w2).all()
    w3 = w0.copy(); m3 = np.zeros_like(w0); v3 = np.zeros_like(w0)
    R.lazy_adam_sparse(w3, m3, v3, np.array([3, 7]), g[:2], st)
    untouched = np.setdiff1d(np.arange(20), [3, 7])
    assert (w3[untouched] == w0[untouched]).all() and not (w3[[3, 7]] == w0[[3, 7]]).any()


def test_map_parameter_model_permit_and_evict():
    mp = R.MapParameterModel(2, default_value=1.0, permit_filter_value=2, evict_filter_value=2)
    mp.get([5, 5, 9])            # step 1: first sighting of 5 and 9 -> not resident
    assert mp.keys().tolist() == []
    mp.get([5])                  # step 2: second sighting of 5 -> resident
    assert mp.keys().tolist() == [5]
    mp.put([5], [[3.0, 4.0]])
    assert mp.get([5, 9]).tolist() == [[3.0, 4.0], [1.0, 1.0]]   # step 3: 9 becomes resident
    mp.get([1]); mp.get([1]); mp.get([1])  # steps 4..6: 5 and 9 unseen for 3 steps
    mp.evict()
    assert mp.keys().tolist() == [1]


def test_c_port_agrees_with_numpy_oracle():
    rng = np.random.default_rng(6)
    vocab, dim, b, f = 500, 16, 64, 39
    tab = (rng.standard_normal((vocab, dim)) * 0.01).astype(np.float32)
    ids = rng.integers(-2, vocab + 3, size=(b, f)).astype(np.int32)
    mask = rng.random((b, f)).astype(np.float32)
    np.testing.assert_array_equal(C.gather_masked(tab, ids, mask), R.gather_masked(tab, ids, mask))
    wt = (rng.standard_normal((vocab, 1)) * 0.01).astype(np.float32)
    np.testing.assert_allclose(C.gather_reduce(wt, ids, mask, 0.5), R.gather_reduce(wt, ids, mask, [0.5]), rtol=1e-6)
    cu, ci, cp, cs = C.unique(ids, vocab)
    ru, ri, rp, rs = R.unique_sorted(ids, bound=vocab)
    for a, bb in ((cu, ru), (ci, ri), (cp, rp), (cs, rs)):
        np.testing.assert_array_equal(a, bb)
    g = rng.standard_normal((b * f, dim)).astype(np.float32)
    np.testing.assert_allclose(C.segment_sum(g, dim, 1, mask.reshape(-1), cp, cs),
                               R.segment_sum(g, ri, ru.size, mask), rtol=1e-6, atol=1e-6)
    vx = rng.standard_normal((8, 39, 16)).astype(np.float32)
    np.testing.assert_allclose(C.fm_fwd(vx), R.fm_forward(vx), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(C.fm_bwd(vx, np.ones(8)), R.fm_backward(vx, np.ones(8)), rtol=1e-5, atol=1e-5)
    x0 = rng.standard_normal((8, 40)).astype(np.float32)
    w = (rng.standard_normal((6, 40)) * 0.1).astype(np.float32)
    bb = (rng.standard_normal((6, 40)) * 0.1).astype(np.float32)
    y, s = C.cross_fwd(x0, w, bb)
    ry, _, rs_ = R.cross_forward(x0, w, bb)
    np.testing.assert_allclose(y, ry, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(s, rs_, rtol=1e-5, atol=1e-5)


def test_c_port_full_step_agrees_with_numpy_oracle():
    from mindrec_b200 import synth
    vocab, dim = 800, 8
    cpu = C.WideDeepCpu(vocab, dim, hidden=(16, 8), seed=1)
    orc = R.WideDeepOracle(cpu.ww, cpu.wd, cpu.w, cpu.b, cpu.wide_b, mode="lazy")
    gen = synth.CriteoSynth(32, cards=[20] * 26, vocab_pad=vocab, seed=2)
    for _ in range(3):
        ids, wts, label = gen.next()
        l1 = cpu.step(ids, wts, label)
        l2, _ = orc.step(ids, wts, label.astype(np.float64))
        np.testing.assert_allclose(l1, l2, rtol=1e-5)
    np.testing.assert_allclose(cpu.wd, orc.wd, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(cpu.ww, orc.ww, rtol=1e-4, atol=1e-6)


def test_map_parameter_model_incremental_export_replays_onto_a_replica():
    """The dict model's incremental export (status 1 = modified, 2 = erased) reproduces the source on a replica —
    the semantics tests/test_hash_gpu.py::test_incremental_export_* demand of the CUDA table."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(seed=st.integers(0, 10_000))
    def run(seed):
        rng = np.random.default_rng(seed)
        src = R.MapParameterModel(4, default_value=0.0, evict_filter_value=3)
        rep = R.MapParameterModel(4, default_value=0.0)
        src.put(np.arange(30), rng.standard_normal((30, 4)).astype(np.float32))
        rep.import_data(src.export_data())
        for _ in range(4):
            op = rng.integers(0, 4)
            ks = rng.choice(60, size=rng.integers(1, 12), replace=False)
            if op == 0:
                src.put(ks, rng.standard_normal((ks.size, 4)).astype(np.float32))
            elif op == 1:
                src.get(ks)
            elif op == 2:
                src.erase(ks)
            else:
                src.evict()
            if rng.random() < 0.5:
                rep.import_data(src.export_data(incremental=True))
        rep.import_data(src.export_data(incremental=True))
        assert sorted(src.rows) == sorted(rep.rows)
        for k in src.rows:
            assert np.array_equal(src.rows[k], rep.rows[k])
        assert src.export_data(incremental=True)[0].size == 0
    run()
