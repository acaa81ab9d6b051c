"""Device-driven sharded exchange (peer_sharded.py): G emulated ranks on ONE GPU vs the oracle on the full tables,
plus the static-capacity (`n_valid`) forms of the dedup and of the fused row updates it relies on."""
import numpy as np
import pytest
import torch

from mindrec_b200 import ops, peer_sharded
from oracle import ref_numpy as ref

pytestmark = pytest.mark.gpu


def _group_inputs(world, vocab, b, f, dim, cuda, seed):
    rng = np.random.default_rng(seed)
    ids = [rng.integers(0, vocab, size=(b, f)).astype(np.int32) for _ in range(world)]
    for r in range(world):                       # hot keys shared by every rank, and duplicates inside a rank
        ids[r][:, 0] = rng.integers(0, 7, size=b)
        ids[r][::3, 1] = ids[r][::3, 2]
    wts = [(rng.random((b, f)) < 0.9).astype(np.float32) for _ in range(world)]
    delta = [rng.standard_normal((b, 1)).astype(np.float32) for _ in range(world)]
    gx = [rng.standard_normal((b, f * dim)).astype(np.float32) for _ in range(world)]
    to = lambda lst: [torch.from_numpy(x).to(cuda) for x in lst]
    return ids, wts, delta, gx, to(ids), to(wts), to(delta), to(gx)


@pytest.mark.parametrize("world,vocab", [(1, 257), (2, 1000), (3, 1000), (8, 4099)])
def test_emulated_ranks_match_oracle_on_full_tables(cuda, world, vocab):
    b, f, dim, sens = 48, 5, 16, 1024.0
    grp = peer_sharded.EmulatedPeerGroup(world, vocab, dim, b * f, cuda, seed=5, sens=sens)
    bias = torch.tensor([0.25], dtype=torch.float32, device=cuda)
    state = dict(w_wide=None)
    for step in range(3):                        # three steps: the inboxes and flags are re-used across steps
        ids, wts, delta, gx, ids_t, wts_t, delta_t, gx_t = _group_inputs(world, vocab, b, f, dim, cuda, 10 + step)
        wide0, deep0 = (t.cpu().numpy().astype(np.float64) for t in grp.full_tables())
        if step == 0:
            acc, lin = np.ones_like(wide0), np.zeros_like(wide0)
            m, v = np.zeros_like(deep0), np.zeros_like(deep0)
            adam = ref.AdamState(3.5e-4, eps=1e-8)
            ftrl = ref.FtrlState(5e-2, l1=1e-8, l2=1e-8)
        deep_outs = [torch.empty((b, f * dim), dtype=torch.float32, device=cuda) for _ in range(world)]
        wide_outs = [torch.empty((b, 1), dtype=torch.float32, device=cuda) for _ in range(world)]
        grp.forward(ids_t, wts_t, bias, deep_outs, wide_outs)
        for r in range(world):
            want = (deep0[ids[r]] * wts[r][..., None]).reshape(b, f * dim)
            assert np.array_equal(deep_outs[r].cpu().numpy(), want.astype(np.float32))      # copies x {0,1}: exact
            want_w = (wide0[ids[r], 0] * wts[r]).sum(1, keepdims=True) + 0.25
            np.testing.assert_allclose(wide_outs[r].cpu().numpy(), want_w, rtol=1e-5, atol=1e-7)
        grp.backward(delta_t, gx_t)
        torch.cuda.synchronize()
        for rk in grp.ranks:
            assert int(rk.err.item()) == 0
        # oracle: one table, the concatenated global batch, gradients divided by sens * G
        ids_cat = np.concatenate(ids).reshape(-1)
        mask_cat = np.concatenate(wts).reshape(-1).astype(np.float64)
        g_deep = np.concatenate(gx).reshape(-1, dim).astype(np.float64) * mask_cat[:, None] / (sens * world)
        g_wide = np.repeat(np.concatenate(delta).astype(np.float64), f, axis=0) * mask_cat[:, None] / (sens * world)
        uniq, inverse = np.unique(ids_cat, return_inverse=True)
        gs_deep = np.zeros((uniq.size, dim))
        np.add.at(gs_deep, inverse, g_deep)
        gs_wide = np.zeros((uniq.size, 1))
        np.add.at(gs_wide, inverse, g_wide)
        adam.begin_step()
        ref.lazy_adam_sparse(deep0, m, v, uniq, gs_deep, adam)
        ref.ftrl_sparse(wide0, acc, lin, uniq, gs_wide, ftrl)
        wide1, deep1 = (t.cpu().numpy() for t in grp.full_tables())
        np.testing.assert_allclose(deep1, deep0, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(wide1, wide0, rtol=1e-5, atol=1e-7)


def test_offsets_cover_every_inbox_exactly_once(cuda):
    """inbox_off / dst_off / src_off derived on the device from the bounds matrix tile each inbox without gaps."""
    world = 4
    rng = np.random.default_rng(0)
    sizes = rng.integers(0, 9, size=(world, world))            # sizes[s][o]: rank s's bucket for owner o
    ball = np.zeros((world, world + 1), dtype=np.int32)
    ball[:, 1:] = np.cumsum(sizes, 1)
    ball_t = torch.from_numpy(ball.reshape(-1)).to(cuda)
    for me in range(world):
        ctrl = torch.tensor([me, world], dtype=torch.int32, device=cuda)
        dst, src = torch.zeros(world, dtype=torch.int32, device=cuda), torch.zeros(world + 1, dtype=torch.int32, device=cuda)
        inbox, n_r = torch.zeros(world, dtype=torch.int32, device=cuda), torch.zeros(1, dtype=torch.int32, device=cuda)
        ops.shard_offsets(ball_t, ctrl, dst, src, inbox, n_r)
        assert dst.tolist() == ball[:, me].tolist()
        assert src.tolist() == [0] + np.cumsum(sizes[:, me]).tolist()
        assert int(n_r) == int(sizes[:, me].sum())
        assert inbox.tolist() == [int(sizes[:me, o].sum()) for o in range(world)]


def test_inbox_overflow_raises_flag_not_a_fault(cuda):
    rows = torch.arange(40, dtype=torch.int32, device=cuda)
    bounds = torch.tensor([0, 40], dtype=torch.int32, device=cuda)
    inbox_off = torch.zeros(1, dtype=torch.int32, device=cuda)
    inbox = torch.full((16,), -1, dtype=torch.int32, device=cuda)
    guard = torch.full((64,), -7, dtype=torch.int32, device=cuda)
    ptrs = torch.tensor([inbox.data_ptr()], dtype=torch.int64, device=cuda)
    err = torch.zeros(1, dtype=torch.int32, device=cuda)
    ops.push_rows_to_peers(rows, bounds, inbox_off, ptrs, torch.empty((16, 0), device=cuda), torch.empty((0, 0), device=cuda), err)
    assert int(err) == 2
    assert inbox.tolist() == list(range(16))
    assert (guard == -7).all()


def test_wait_times_out_with_error_bit(cuda):
    flags = torch.zeros(2, dtype=torch.int32, device=cuda)
    flags[0] = 5
    epoch = torch.tensor([5], dtype=torch.int32, device=cuda)
    err = torch.zeros(1, dtype=torch.int32, device=cuda)
    ops.peer_wait(flags[:1], epoch, err)
    assert int(err) == 0
    ops.peer_wait(flags, epoch, err, max_cycles_log2=24)          # rank 1 never signals: ~10 ms bounded spin
    assert int(err) == 1


@pytest.mark.parametrize("n_valid", [0, 1, 37, 1000])
def test_unique_with_device_valid_count(cuda, n_valid):
    cap, vocab = 1000, 300
    rng = np.random.default_rng(n_valid)
    ids = rng.integers(0, vocab, size=cap).astype(np.int32)
    ids_t = torch.from_numpy(ids).to(cuda)
    nv = torch.tensor([n_valid], dtype=torch.int32, device=cuda)
    uq = ops.unique(ids_t, table_like=torch.empty((vocab, 0), device=cuda), n_valid=nv)
    want = np.unique(ids[:n_valid])                    # the padding past n_valid is never looked at
    assert int(uq.count) == want.size
    assert np.array_equal(uq.uniq[:want.size].cpu().numpy(), want)
    inv = uq.inverse.cpu().numpy()[:n_valid]
    assert np.array_equal(want[inv], ids[:n_valid])
    assert int(uq.seg_start[want.size]) == n_valid


@pytest.mark.parametrize("dim", [1, 16])
def test_row_updates_with_device_valid_count_ignore_the_padding(cuda, dim):
    cap, vocab, n_valid = 512, 200, 131
    rng = np.random.default_rng(dim)
    ids = rng.integers(0, vocab, size=cap).astype(np.int32)
    g = rng.standard_normal((cap, dim)).astype(np.float32)
    nv = torch.tensor([n_valid], dtype=torch.int32, device=cuda)
    like = torch.empty((vocab, 0), device=cuda)

    def run(ids_np, g_np, n_valid_t):
        torch.manual_seed(0)
        w = torch.randn((vocab, dim), device=cuda)
        m, v = torch.zeros_like(w), torch.zeros_like(w)
        acc, lin = torch.ones_like(w), torch.zeros_like(w)
        w2 = w.clone()
        ids_t, g_t = torch.from_numpy(ids_np).to(cuda), torch.from_numpy(g_np).to(cuda)
        uq = ops.unique(ids_t, table_like=like, n_valid=n_valid_t)
        ah = ops.adam_hyper(1e-2, device=cuda)
        ops.adam_begin_step(ah)
        ops.sparse_lazy_adam(w, m, v, ah, g_t, None, uq, n_valid=n_valid_t)
        ops.sparse_ftrl(w2, acc, lin, ops.ftrl_hyper(5e-2, l1=1e-3, l2=1e-3, device=cuda), g_t, None, uq, n_valid=n_valid_t)
        return [t.cpu().numpy() for t in (w, m, v, w2, acc, lin)]

    got = run(ids, g, nv)                                  # padded inbox + device-side count
    want = run(ids[:n_valid].copy(), g[:n_valid].copy(), None)   # the exact-size call
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("world", [1, 2, 4])
def test_emulated_hash_sharding_matches_one_table(cuda, world):
    """MapParameter sharded by hash(key) mod G: rows are bit-identical to ONE table fed the same keys
    (a key's initial row is keyed by the key, not by its slot or owner), and after LazyAdam the union of the G
    tables equals the single table."""
    from mindrec_b200 import hash as H
    dim, n, bits = 16, 400, 40
    rng = np.random.default_rng(world)
    grp = peer_sharded.EmulatedPeerHashGroup(world, dim, n, cuda, key_bits=bits, capacity=1 << 12, seed=7,
                                             learning_rate=1e-2)
    one = H.MapParameter(key_dtype=torch.int64, value_shape=dim, default_value="normal", capacity=1 << 14,
                         device=cuda, seed=7)
    m1, v1 = one.add_arena(0.0), one.add_arena(0.0)
    hyper = ops.adam_hyper(1e-2, device=cuda)
    pool = rng.integers(0, 1 << bits, size=3000)
    for step in range(3):
        keys = [rng.choice(pool[: 1000 * (step + 1)], size=n) for _ in range(world)]        # new keys every step
        for k in keys:
            k[:20] = pool[:20]                                                              # shared across ranks
        grads = [rng.standard_normal((n, dim)).astype(np.float32) for _ in range(world)]
        keys_t = [torch.from_numpy(k).to(cuda) for k in keys]
        outs = [torch.empty((n, dim), device=cuda) for _ in range(world)]
        grp.forward(keys_t, outs)
        all_keys = torch.cat(keys_t)
        slots = one.lookup_slots(all_keys).clone()
        want = ops.gather(one.values, slots).view(world, n, dim)
        for r in range(world):
            if step == 0:                            # fresh rows: bit-identical (initialisation keyed by the key)
                assert torch.equal(outs[r], want[r])
            else:                                    # updated rows: the gradient sums associate differently
                torch.testing.assert_close(outs[r], want[r], rtol=1e-5, atol=1e-7)
        grp.backward([torch.from_numpy(g).to(cuda) for g in grads])
        ops.adam_begin_step(hyper)
        uq = ops.unique(slots, table_like=torch.empty((one.capacity, 0), device=cuda))
        c = one.capacity
        ops.sparse_lazy_adam(one.values[:c], m1[:c], v1[:c], hyper, torch.from_numpy(np.concatenate(grads)).to(cuda), None, uq)
        torch.cuda.synchronize()
        for rk in grp.ranks:
            assert int(rk.err.item()) == 0 and not rk.table.overflowed
        k_sh, v_sh = grp.get_data()
        k_1, v_1 = one.get_data()
        order = torch.argsort(k_1)
        assert torch.equal(k_sh, k_1[order])
        torch.testing.assert_close(v_sh, v_1[order], rtol=1e-5, atol=1e-7)
    assert sum(len(rk.table) for rk in grp.ranks) == len(one)
    if world > 1:                                   # the owner function spreads the keys
        sizes = [len(rk.table) for rk in grp.ranks]
        assert min(sizes) > 0.5 * len(one) / world


def test_shard_remap_hash_and_fill_tail(cuda):
    g, bits = 4, 20
    keys = torch.tensor([0, 5, (1 << bits) - 1, 1 << bits, -1, -2, -7, 123456], dtype=torch.int64, device=cuda)
    out = ops.shard_remap_hash(keys, torch.empty((g, 0), device=cuda), torch.empty((bits, 0), device=cuda)).tolist()
    for k, o in zip(keys.tolist(), out):
        if 0 <= k < (1 << bits):
            assert o & ((1 << bits) - 1) == k and 0 <= (o >> bits) < g
        else:
            assert o == g << bits
    k32 = keys[:3].to(torch.int32)
    assert ops.shard_remap_hash(k32, torch.empty((g, 0), device=cuda), torch.empty((bits, 0), device=cuda)).tolist() == out[:3]
    buf = torch.arange(10, dtype=torch.int64, device=cuda)
    ops.fill_tail(buf, torch.tensor([4], dtype=torch.int32, device=cuda), torch.tensor([-1], dtype=torch.int64, device=cuda))
    assert buf.tolist() == [0, 1, 2, 3] + [-1] * 6


@pytest.mark.parametrize("world", [2, 4])
def test_key_phase_one_step_ahead_on_the_second_buffer_set(cuda, world):
    """The graphed step's steady state on the real kernels: batch t + 1's key phase (plan, publish, key push, owner
    dedup) runs on the other buffer set while step t still has its backward to do; step t + 1 adopts it and starts at
    serve.  Same four batches through an in-line group and a one-step-ahead group: bit-identical tables."""
    vocab, b, f, dim = 2003, 40, 6, 16
    rng = np.random.default_rng(3)

    def batch():
        t = lambda a: torch.from_numpy(a).to(cuda)
        ids = [t(rng.integers(0, vocab, size=(b, f)).astype(np.int32)) for _ in range(world)]
        wts = [t((rng.random((b, f)) < 0.9).astype(np.float32)) for _ in range(world)]
        delta = [t(rng.standard_normal((b, 1)).astype(np.float32)) for _ in range(world)]
        gx = [t(rng.standard_normal((b, f * dim)).astype(np.float32)) for _ in range(world)]
        return ids, wts, delta, gx
    batches = [batch() for _ in range(4)]
    bias = torch.tensor([0.1], device=cuda)
    inline = peer_sharded.EmulatedPeerGroup(world, vocab, dim, b * f, cuda, seed=5)
    ahead = peer_sharded.EmulatedPeerGroup(world, vocab, dim, b * f, cuda, seed=5)
    outs = lambda: ([torch.empty((b, f * dim), device=cuda) for _ in range(world)],
                    [torch.empty((b, 1), device=cuda) for _ in range(world)])
    split_seen = False
    for t, (ids, wts, delta, gx) in enumerate(batches):
        (di, wi), (da, wa) = outs(), outs()
        inline.forward(ids, wts, bias, di, wi)
        ahead.forward(ids, wts, bias, da, wa, planned=t > 0)
        for r in range(world):
            assert torch.equal(di[r], da[r]) and torch.equal(wi[r], wa[r])
        if t + 1 < len(batches):
            # split serve on the real kernels: the landing buffers are poisoned, the rows the coming update does not
            # touch go out now (mode 1), the rest after the update (mode 2): a row written by neither shows up as NaN
            for rk in ahead.ranks:
                rk.buf["land_deep"].fill_(float("nan"))
                rk.buf["land_wide"].fill_(float("nan"))
            ahead.key_phase_next(batches[t + 1][0])
            early = [int(torch.isfinite(rk.buf["land_wide"]).sum()) for rk in ahead.ranks]
            n_u = [int(rk._s(nxt=True)["bounds"][world]) for rk in ahead.ranks]
            assert all(0 < e <= u for e, u in zip(early, n_u))
            split_seen = split_seen or any(e < u for e, u in zip(early, n_u))
        inline.backward(delta, gx)
        ahead.backward(delta, gx)
        for a_, b_ in zip(inline.full_tables(), ahead.full_tables()):
            assert torch.equal(a_, b_)
    assert split_seen
    for rk in ahead.ranks:
        assert int(rk.err.item()) == 0 and rk.cur == 1


@pytest.mark.parametrize("world,n", [(1, 1000), (2, 4099), (4, 1 << 20), (8, 3_885_058), (3, 7)])
def test_peer_allreduce_sums_in_rank_order_on_every_rank(cuda, world, n):
    """mrec_peer_allreduce with G emulated ranks on one GPU (every rank's call; pointer tables as in PeerAllReduce):
    every destination holds the fp32 sum added in rank order — bit-exact, identical on all ranks — incl. an n that is
    not a multiple of 4, a slice boundary inside the buffer and the DenseLayer gradient size of the benchmark."""
    gen = torch.Generator(device=cuda)
    gen.manual_seed(n + world)
    src = [torch.randn(n, device=cuda, generator=gen) * (10.0 ** (r % 3)) for r in range(world)]
    dst = [torch.full((n,), float("nan"), device=cuda) for _ in range(world)]
    p_src = torch.tensor([t.data_ptr() for t in src], dtype=torch.int64, device=cuda)
    p_dst = torch.tensor([t.data_ptr() for t in dst], dtype=torch.int64, device=cuda)
    for r in range(world):
        ops.peer_allreduce(p_src, p_dst, torch.tensor([r, world], dtype=torch.int32, device=cuda), dst[r])
    want = torch.zeros(n, device=cuda)
    for r in range(world):
        want = want + src[r]                    # 0 + s0 + s1 + ... in fp32, the kernel's order
    for r in range(world):
        assert torch.equal(dst[r], want)
