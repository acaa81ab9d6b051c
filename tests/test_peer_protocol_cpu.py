"""The device-driven exchange protocol (peer_sharded.PeerRank / EmulatedPeerGroup) driven on the CPU with stand-in
kernels (tests/fake_peer_ops.py): phase order, buffer wiring and the pointer tables of `connect` are the real
product code; G emulated ranks must equal the oracle on the full tables (same check as the GPU emulation test)."""
import numpy as np
import pytest
import torch

from mindrec_b200 import peer_sharded
from oracle import ref_numpy as R
from tests import fake_peer_ops


@pytest.fixture()
def cpu_ops(monkeypatch):
    monkeypatch.setattr(peer_sharded, "ops", fake_peer_ops)
    return fake_peer_ops


@pytest.mark.parametrize("world,vocab", [(1, 53), (2, 200), (3, 101), (4, 64)])
def test_emulated_ranks_on_cpu_match_oracle_on_full_tables(cpu_ops, world, vocab):
    b, f, dim, sens = 12, 4, 8, 1024.0
    grp = peer_sharded.EmulatedPeerGroup(world, vocab, dim, b * f, "cpu", seed=5, sens=sens)
    bias = torch.tensor([0.25], dtype=torch.float32)
    rng = np.random.default_rng(world)
    for step in range(3):
        ids = [rng.integers(0, vocab, size=(b, f)).astype(np.int32) for _ in range(world)]
        for r in range(world):
            ids[r][:, 0] = rng.integers(0, 5, size=b)                  # keys shared by every rank
        wts = [(rng.random((b, f)) < 0.9).astype(np.float32) for _ in range(world)]
        delta = [rng.standard_normal((b, 1)).astype(np.float32) for _ in range(world)]
        gx = [rng.standard_normal((b, f * dim)).astype(np.float32) for _ in range(world)]
        wide0, deep0 = (t.numpy().astype(np.float64) for t in grp.full_tables())
        if step == 0:
            acc, lin = np.ones_like(wide0), np.zeros_like(wide0)
            m, v = np.zeros_like(deep0), np.zeros_like(deep0)
            adam, ftrl = R.AdamState(3.5e-4, eps=1e-8), R.FtrlState(5e-2, l1=1e-8, l2=1e-8)
        t = lambda lst: [torch.from_numpy(x) for x in lst]
        deep_outs = [torch.empty((b, f * dim)) for _ in range(world)]
        wide_outs = [torch.empty((b, 1)) for _ in range(world)]
        grp.forward(t(ids), t(wts), bias, deep_outs, wide_outs)
        for r in range(world):
            want = (deep0[ids[r]] * wts[r][..., None]).reshape(b, f * dim).astype(np.float32)
            assert np.array_equal(deep_outs[r].numpy(), want)
            want_w = (wide0[ids[r], 0] * wts[r]).sum(1, keepdims=True) + 0.25
            np.testing.assert_allclose(wide_outs[r].numpy(), want_w, rtol=1e-5, atol=1e-7)
        grp.backward(t(delta), t(gx))
        for rk in grp.ranks:
            assert int(rk.err) == 0                                     # no wait saw a missing signal, no overflow
        ids_cat = np.concatenate(ids).reshape(-1)
        mask_cat = np.concatenate(wts).reshape(-1).astype(np.float64)
        g_deep = np.concatenate(gx).reshape(-1, dim).astype(np.float64) * mask_cat[:, None] / (sens * world)
        g_wide = np.repeat(np.concatenate(delta).astype(np.float64), f, axis=0) * mask_cat[:, None] / (sens * world)
        uniq, inverse = np.unique(ids_cat, return_inverse=True)
        gs_deep, gs_wide = np.zeros((uniq.size, dim)), np.zeros((uniq.size, 1))
        np.add.at(gs_deep, inverse, g_deep)
        np.add.at(gs_wide, inverse, g_wide)
        adam.begin_step()
        R.lazy_adam_sparse(deep0, m, v, uniq, gs_deep, adam)
        R.ftrl_sparse(wide0, acc, lin, uniq, gs_wide, ftrl)
        wide1, deep1 = (t_.numpy() for t_ in grp.full_tables())
        np.testing.assert_allclose(deep1, deep0, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(wide1, wide0, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("world", [2, 3])
def test_key_phase_one_step_ahead_on_the_second_buffer_set(cpu_ops, world):
    """Steady state of the graphed step: while step t still has its backward to do, the key phase of batch t + 1
    (plan, publish, key push, owner dedup) runs on the OTHER buffer set; step t + 1 adopts it and starts at serve.
    Two ranks groups run the same four batches, one planning in line, one a step ahead: identical tables."""
    vocab, b, f, dim = 211, 10, 4, 8
    rng = np.random.default_rng(3)
    def batch():
        ids = [torch.from_numpy(rng.integers(0, vocab, size=(b, f)).astype(np.int32)) for _ in range(world)]
        wts = [torch.from_numpy((rng.random((b, f)) < 0.9).astype(np.float32)) for _ in range(world)]
        delta = [torch.from_numpy(rng.standard_normal((b, 1)).astype(np.float32)) for _ in range(world)]
        gx = [torch.from_numpy(rng.standard_normal((b, f * dim)).astype(np.float32)) for _ in range(world)]
        return ids, wts, delta, gx
    batches = [batch() for _ in range(4)]
    bias = torch.tensor([0.1])
    inline = peer_sharded.EmulatedPeerGroup(world, vocab, dim, b * f, "cpu", seed=5)
    ahead = peer_sharded.EmulatedPeerGroup(world, vocab, dim, b * f, "cpu", seed=5)
    outs = lambda: ([torch.empty((b, f * dim)) for _ in range(world)], [torch.empty((b, 1)) for _ in range(world)])
    split_seen = False
    for t, (ids, wts, delta, gx) in enumerate(batches):
        di, wi = outs()
        da, wa = outs()
        inline.forward(ids, wts, bias, di, wi)
        ahead.forward(ids, wts, bias, da, wa, planned=t > 0)
        for r in range(world):
            assert torch.equal(di[r], da[r]) and torch.equal(wi[r], wa[r])
        if t + 1 < len(batches):
            # split serve: poison the landing buffers (their rows were consumed by the expand above), then the rows the
            # coming update does not touch go out early (under step t's DenseLayers) and the rest after it — every row of
            # batch t + 1 must be written by exactly one of the two, or the NaNs show up in the next forward
            for rk in ahead.ranks:
                rk.buf["land_deep"].fill_(float("nan"))
                rk.buf["land_wide"].fill_(float("nan"))
            ahead.key_phase_next(batches[t + 1][0])
            early = [int(torch.isfinite(rk.buf["land_wide"]).sum()) for rk in ahead.ranks]
            n_u = [int(rk._s(nxt=True)["bounds"][world]) for rk in ahead.ranks]
            dirty = [int(sum(bin(int(w) & 0xffffffff).count("1") for w in rk.dirty.tolist())) for rk in ahead.ranks]
            assert all(d > 0 for d in dirty)                           # the update of step t names rows
            assert all(0 < e <= u for e, u in zip(early, n_u))         # part of batch t + 1 went out early ...
            split_seen = split_seen or any(e < u for e, u in zip(early, n_u))      # ... and part of it had to wait
        inline.backward(delta, gx)
        ahead.backward(delta, gx)
        for a_, b_ in zip(inline.full_tables(), ahead.full_tables()):
            assert torch.equal(a_, b_)
    assert split_seen
    for rk in ahead.ranks:
        assert int(rk.err) == 0 and rk.cur == 1                        # three adoptions: the sets alternated


def test_look_ahead_plan_equals_inline_plan_on_cpu(cpu_ops):
    """p_plan_local(nxt=True) + p_adopt leaves exactly the state of an inline p_plan_local."""
    vocab, n = 300, 40
    alloc = lambda name, shape, dtype: torch.zeros(shape, dtype=dtype)
    a = peer_sharded.PeerRank(1, 3, vocab, 8, n, "cpu", alloc)
    bk = peer_sharded.PeerRank(1, 3, vocab, 8, n, "cpu", alloc)
    ids = torch.from_numpy(np.random.default_rng(0).integers(-2, vocab + 3, size=n).astype(np.int32))
    a.p_plan_local(ids)
    bk.p_plan_local(ids, nxt=True)
    bk.p_adopt()
    u = int(a.uq.count)
    assert int(bk.uq.count) == u
    for fa, fb, k in ((a.uq.uniq, bk.uq.uniq, u), (a.uq.inverse, bk.uq.inverse, n), (a.uq.perm, bk.uq.perm, n),
                      (a.uq.seg_of, bk.uq.seg_of, n), (a.uq.seg_start, bk.uq.seg_start, u + 1)):
        assert torch.equal(fa[:k], fb[:k])
    assert torch.equal(a.bounds, bk.bounds)


@pytest.mark.parametrize("world", [1, 2, 4])
def test_emulated_hash_sharding_on_cpu_matches_one_table(cpu_ops, monkeypatch, world):
    """PeerHashRank's orchestration (owner-major int64 keys, key inbox blanking, serve by slot, update by slot with
    the device-side valid count) against ONE stand-in table fed the same keys."""
    from mindrec_b200 import hash as H
    monkeypatch.setattr(H, "MapParameter", fake_peer_ops.FakeMapParameter)
    dim, n, bits = 8, 60, 30
    rng = np.random.default_rng(world)
    # inboxes sized for the worst case here (every rank's keys on one owner); the default capacity is n per rank
    grp = peer_sharded.EmulatedPeerHashGroup(world, dim, n, "cpu", key_bits=bits, capacity=1 << 9, seed=7, learning_rate=1e-2,
                                             cap_rows=world * n)
    one = fake_peer_ops.FakeMapParameter(value_shape=dim, capacity=1 << 11, seed=7)
    m1, v1 = one.add_arena(0.0), one.add_arena(0.0)
    hyper = fake_peer_ops.adam_hyper(1e-2, device="cpu")
    pool = rng.integers(0, 1 << bits, size=300)
    for step in range(3):
        keys = [rng.choice(pool[: 100 * (step + 1)], size=n) for _ in range(world)]
        for k in keys:
            k[:6] = pool[:6]                                            # shared across ranks
        grads = [(np.abs(rng.standard_normal((n, dim))) + 0.5).astype(np.float32) for _ in range(world)]
        keys_t = [torch.from_numpy(k) for k in keys]
        outs = [torch.empty((n, dim)) for _ in range(world)]
        grp.forward(keys_t, outs)
        slots = one.lookup_slots(torch.cat(keys_t)).clone()
        want = fake_peer_ops.gather(one.values, slots).view(world, n, dim)
        for r in range(world):
            torch.testing.assert_close(outs[r], want[r], rtol=1e-5, atol=1e-8)
        grp.backward([torch.from_numpy(g) for g in grads])
        fake_peer_ops.adam_begin_step(hyper)
        uq = fake_peer_ops.unique(slots, table_like=torch.empty((one.capacity, 0)))
        c = one.capacity
        fake_peer_ops.sparse_lazy_adam(one.values[:c], m1[:c], v1[:c], hyper, torch.from_numpy(np.concatenate(grads)), None, uq)
        for rk in grp.ranks:
            assert int(rk.err) == 0 and not rk.table.overflowed
        k_sh, v_sh = grp.get_data()
        k_1, v_1 = one.get_data()
        assert torch.equal(k_sh, k_1)
        torch.testing.assert_close(v_sh, v_1, rtol=1e-5, atol=1e-8)
    assert sum(len(rk.table) for rk in grp.ranks) == len(one)


def test_inbox_overflow_is_flagged_on_cpu(cpu_ops, monkeypatch):
    """An owner that is sent more keys than its inbox holds raises error bit 2 (bench.py aborts on it) — no stray store."""
    from mindrec_b200 import hash as H
    monkeypatch.setattr(H, "MapParameter", fake_peer_ops.FakeMapParameter)
    world, dim, n = 4, 4, 16
    grp = peer_sharded.EmulatedPeerHashGroup(world, dim, n, "cpu", key_bits=20, capacity=1 << 8, seed=1)
    keys = [torch.arange(r * n, (r + 1) * n, dtype=torch.int64) * world for r in range(world)]     # all owned by rank 0
    outs = [torch.empty((n, dim)) for _ in range(world)]
    grp.forward(keys, outs)
    assert any(int(rk.err) & 2 for rk in grp.ranks)
