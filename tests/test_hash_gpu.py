"""K6 parity: the GPU hash table behind MapParameter / HashEmbeddingLookup vs the dict model of the oracle.
Key membership, counts and returned rows are exact; slot indices are implementation-private."""
import numpy as np
import pytest
import torch

from mindrec_b200 import hash as H
from mindrec_b200 import ops
from oracle import ref_numpy as R

pytestmark = pytest.mark.gpu


def _keys(rng, n, space, dtype):
    return (rng.zipf(1.2, size=n) % space).astype(dtype)


@pytest.mark.parametrize("kdt", [torch.int32, torch.int64])
def test_get_put_erase_against_dict_model(cuda, kdt):
    rng = np.random.default_rng(0)
    dim = 8
    mp = H.MapParameter(key_dtype=kdt, value_shape=dim, default_value=0.5, capacity=4096, device=cuda)
    model = R.MapParameterModel(dim, default_value=0.5)
    npdt = np.int32 if kdt == torch.int32 else np.int64
    space = 3000 if kdt == torch.int32 else 2 ** 40
    for it in range(6):
        keys = _keys(rng, 700, space, npdt)
        got = mp.get(torch.from_numpy(keys).to(cuda)).cpu().numpy()
        np.testing.assert_array_equal(got, model.get(keys))
        pk = rng.choice(keys, size=50, replace=False) if it % 2 == 0 else _keys(rng, 50, space, npdt)
        pk = np.unique(pk)
        pv = rng.standard_normal((pk.size, dim)).astype(np.float32)
        mp.put(torch.from_numpy(pk).to(cuda), torch.from_numpy(pv).to(cuda))
        model.put(pk, pv)
        ek = np.unique(rng.choice(keys, size=30))
        mp.erase(torch.from_numpy(ek).to(cuda))
        model.erase(ek)
        np.testing.assert_array_equal(mp.get_keys().cpu().numpy().astype(np.int64), model.keys())
        assert len(mp) == model.keys().size
    k, v = mp.get_data()
    ref_rows = np.stack([model.rows[int(x)] for x in k.cpu().numpy()])
    np.testing.assert_array_equal(v.cpu().numpy(), ref_rows)
    assert not mp.overflowed


def test_reserved_keys_and_missing_keys_read_default_row(cuda):
    mp = H.MapParameter(key_dtype=torch.int64, value_shape=4, default_value="ones", capacity=64, device=cuda)
    keys = torch.tensor([5, -1, -2, 5, 9], dtype=torch.int64, device=cuda)
    out = mp.get(keys, insert_default_value=False)
    assert torch.equal(out, torch.ones((5, 4), device=cuda)) and len(mp) == 0
    mp.get(keys)
    assert mp.get_keys().tolist() == [5, 9]          # -1 / -2 are never stored (embedding.py:55-57)


def test_duplicates_in_one_call_resolve_to_one_slot(cuda):
    mp = H.MapParameter(key_dtype=torch.int32, value_shape=4, default_value="normal", capacity=1 << 16, device=cuda)
    keys = torch.randint(0, 50, (20000,), dtype=torch.int32, device=cuda)
    slots = mp.lookup_slots(keys).clone()
    assert len(mp) == torch.unique(keys).numel()
    # same key -> same slot, different keys -> different slots
    pairs = torch.unique(torch.stack([keys.long(), slots.long()], 1), dim=0)
    assert pairs.shape[0] == torch.unique(keys).numel() == torch.unique(slots).numel()
    rows = mp.get(keys)
    assert torch.equal(rows, mp.values[slots.long()])


def test_permit_and_evict_filters_follow_the_model(cuda):
    dim = 4
    mp = H.MapParameter(key_dtype=torch.int32, value_shape=dim, default_value=2.0, permit_filter_value=3,
                        evict_filter_value=2, capacity=1024, device=cuda)
    model = R.MapParameterModel(dim, default_value=2.0, permit_filter_value=3, evict_filter_value=2)
    rng = np.random.default_rng(1)
    for it in range(12):
        keys = rng.integers(0, 40, size=60).astype(np.int32)
        got = mp.get(torch.from_numpy(keys).to(cuda)).cpu().numpy()
        np.testing.assert_array_equal(got, model.get(keys))
        if it % 3 == 2:
            mp.evict()
            model.evict()
        np.testing.assert_array_equal(mp.get_keys().cpu().numpy().astype(np.int64), model.keys())


def test_normal_init_is_keyed_by_key_and_has_the_right_moments(cuda):
    a = H.MapParameter(key_dtype=torch.int64, value_shape=80, default_value="normal", capacity=1 << 18, device=cuda, seed=7)
    b = H.MapParameter(key_dtype=torch.int64, value_shape=80, default_value="normal", capacity=1 << 17, device=cuda, seed=7)
    keys = torch.randperm(50000, device=cuda)[:20000].to(torch.int64) * 7919
    ra = a.get(keys)
    rb = b.get(keys.flip(0)).flip(0)
    assert torch.equal(ra, rb)                       # row depends on (seed, key), not on slot or arrival order
    assert abs(float(ra.mean())) < 2e-4 and abs(float(ra.std()) - 0.01) < 2e-4
    assert torch.equal(a.get(keys), ra)              # second lookup returns the stored rows


def test_full_table_raises_overflow_flag_not_corruption(cuda):
    mp = H.MapParameter(key_dtype=torch.int32, value_shape=2, default_value=0.0, capacity=64, device=cuda)
    keys = torch.arange(200, dtype=torch.int32, device=cuda)
    out = mp.get(keys)
    assert mp.overflowed and len(mp) == 64
    assert out.shape == (200, 2)


def test_hash_embedding_lookup_matches_reference_flow(cuda):
    """embedding.py:189-195: Unique -> MapTensorGet -> Gather(inverse) equals a direct lookup."""
    emb = H.HashEmbeddingLookup(16, key_dtype=torch.int64, param_init="normal", capacity=1 << 14, device=cuda)
    ids = torch.randint(0, 2 ** 40, (64, 26), dtype=torch.int64, device=cuda)
    ids[:, 0] = 12345
    out = emb(ids)
    assert out.shape == (64, 26, 16)
    uq = ops.unique(ids)
    u = int(uq.count.item())
    weight_unique = emb.embedding_table.get(uq.uniq[:u])
    ref = weight_unique[uq.inverse.long()].view(64, 26, 16)
    assert torch.equal(out, ref)
    assert torch.equal(out[0, 0], out[63, 0])


def test_hash_embedding_lookup_argument_validation():
    with pytest.raises(TypeError, match="sparse"):
        H.HashEmbeddingLookup(8, sparse="yes", device="cpu")
    with pytest.raises(ValueError, match="vocab_cache_size"):
        H.HashEmbeddingLookup(8, vocab_cache_size=-1, device="cpu")
    with pytest.raises(RuntimeError, match="parameter server"):
        H.HashEmbeddingLookup(8, vocab_cache_size=10, device="cpu")
    with pytest.raises(ValueError, match="embedding_size"):
        H.HashEmbeddingLookup(0, device="cpu")


def test_lazy_adam_on_map_parameter_by_slot(cuda):
    """Optimizer on a MapParameter: rows addressed by slot, state in sibling arenas."""
    dim = 16
    mp = H.MapParameter(key_dtype=torch.int64, value_shape=dim, default_value="normal", capacity=1 << 12, device=cuda)
    m, v = mp.add_arena(0.0), mp.add_arena(0.0)
    gen = torch.Generator(device=cuda)
    gen.manual_seed(17)
    keys = torch.randint(0, 300, (1000,), dtype=torch.int64, device=cuda, generator=gen) * 1000003
    slots = mp.lookup_slots(keys).clone()
    w0 = mp.values.clone()
    # positive gradients: a per-key sum that cancels to ~0 has no 1e-5 relative meaning and Adam's m / sqrt(v) is
    # discontinuous there (an unseeded draw once landed on such a key)
    g = torch.randn((1000, dim), device=cuda, generator=gen).abs() + 0.5
    hyper = ops.adam_hyper(1e-2, device=cuda)
    ops.adam_begin_step(hyper)
    uq = ops.unique(slots, table_like=mp.values)
    ops.sparse_lazy_adam(mp.values, m, v, hyper, g, None, uq)
    # reference on the host, keyed by key
    st = R.AdamState(1e-2); st.begin_step()
    w_ref, m_ref, v_ref = w0.cpu().numpy(), np.zeros_like(w0.cpu().numpy()), np.zeros_like(w0.cpu().numpy())
    uniq, inverse, _, _ = R.unique_sorted(slots.cpu().numpy())
    R.lazy_adam_sparse(w_ref, m_ref, v_ref, uniq, R.segment_sum(g.cpu().numpy(), inverse, uniq.size), st)
    np.testing.assert_allclose(mp.values.cpu().numpy(), w_ref, rtol=1e-5, atol=1e-7)
    assert torch.equal(mp.values[mp.capacity], w0[mp.capacity])   # default row untouched


def test_optimizer_never_moves_the_default_row(cuda):
    """with permit_filter_value > 1 (or a full table) lookups return slot C, the shared default row.  An
    optimizer driven through as_parameter() must leave row C of the values and of every moment arena alone."""
    dim = 8
    mp = H.MapParameter(key_dtype=torch.int64, value_shape=dim, default_value="ones", permit_filter_value=2,
                        capacity=64, device=cuda)
    m, v = mp.add_arena(0.0), mp.add_arena(0.0)
    keys = torch.arange(300, dtype=torch.int64, device=cuda) * 7 + 1
    hyper = ops.adam_hyper(1e-2, device=cuda)
    p = mp.as_parameter()
    assert p.data.shape[0] == mp.capacity
    for it in range(3):                                # first pass: nothing admitted; later: table overflows
        slots = mp.lookup_slots(keys).clone()
        assert bool((slots == mp.capacity).any())
        g = torch.ones((keys.numel(), dim), device=cuda)
        uq = ops.unique(slots, table_like=p.data)
        ops.adam_begin_step(hyper)
        ops.sparse_lazy_adam(p.data, mp.arena_rows(m), mp.arena_rows(v), hyper, g, None, uq)
    assert mp.overflowed
    c = mp.capacity
    assert torch.equal(mp.values[c], torch.ones(dim, device=cuda))
    assert torch.equal(m[c], torch.zeros(dim, device=cuda)) and torch.equal(v[c], torch.zeros(dim, device=cuda))
    touched = torch.unique(slots[slots < c]).long()
    assert touched.numel() == c and bool((mp.values[touched] != 1.0).all())


def _sorted_data(mp):
    k, v = mp.get_data()
    order = torch.argsort(k)
    return k[order].cpu().numpy(), v[order].cpu().numpy()


@pytest.mark.parametrize("kdt", [torch.int32, torch.int64])
def test_incremental_export_replays_onto_a_replica(cuda, kdt):
    """Full export, then two rounds of put / touch / erase / evict, each shipped as an incremental export:
    the replica ends bit-identical and the statuses name exactly the changed and the erased keys."""
    rng = np.random.default_rng(3)
    src = H.MapParameter(key_dtype=kdt, value_shape=8, capacity=1 << 12, default_value="zeros", device=cuda,
                         evict_filter_value=3)
    rep = H.MapParameter(key_dtype=kdt, value_shape=8, capacity=1 << 12, default_value="zeros", device=cuda)
    t = lambda a, dt=kdt: torch.as_tensor(a, dtype=dt, device=cuda)
    keys0 = rng.choice(100000, size=600, replace=False)
    src.put(t(keys0), torch.from_numpy(rng.standard_normal((600, 8)).astype(np.float32)).to(cuda))
    k, v, st = src.export_data()
    assert int(st.sum()) == 0 and k.numel() == 600
    rep.import_data((k, v, st))
    for a, b in zip(_sorted_data(src), _sorted_data(rep)):
        assert np.array_equal(a, b)

    # round 1: overwrite 50 old keys, add 40 new ones, erase 30, erase-and-reinsert 5
    over, new = keys0[:50], np.arange(200000, 200040)
    gone, back = keys0[100:130], keys0[130:135]
    src.put(t(over), torch.full((50, 8), 2.5, device=cuda))
    src.put(t(new), torch.full((40, 8), -1.0, device=cuda))
    src.erase(t(gone))
    src.erase(t(back))
    src.put(t(back), torch.full((5, 8), 9.0, device=cuda))
    k, v, st = src.export_data(incremental=True)
    got_mod = set(k[st == H.MapParameter.STATUS_MODIFIED].tolist())
    got_gone = set(k[st == H.MapParameter.STATUS_ERASED].tolist())
    assert got_mod == set(over.tolist()) | set(new.tolist()) | set(back.tolist())
    assert got_gone == set(gone.tolist())
    rep.import_data((k, v, st))
    for a, b in zip(_sorted_data(src), _sorted_data(rep)):
        assert np.array_equal(a, b)

    # round 2: only 20 keys are looked up for four steps; everything else falls to the eviction filter
    alive = keys0[200:220]
    for _ in range(5):
        src.get(t(alive))
    src.evict()
    assert len(src) == 20
    k, v, st = src.export_data(incremental=True)
    assert set(k[st == H.MapParameter.STATUS_MODIFIED].tolist()) == set(alive.tolist())
    assert int((st == H.MapParameter.STATUS_ERASED).sum()) == len(rep) - 20
    rep.import_data((k, v, st))
    for a, b in zip(_sorted_data(src), _sorted_data(rep)):
        assert np.array_equal(a, b)
    # nothing happened since: the next incremental export is empty
    k, v, st = src.export_data(incremental=True)
    assert k.numel() == 0


def test_incremental_export_falls_back_to_full_when_the_erase_log_overflows(cuda):
    mp = H.MapParameter(key_dtype=torch.int64, value_shape=4, capacity=1 << 12, default_value="zeros", device=cuda)
    mp._erase_log = torch.empty(8, dtype=torch.int64, device=cuda)           # tiny log
    keys = torch.arange(100, dtype=torch.int64, device=cuda)
    mp.put(keys, torch.ones((100, 4), device=cuda))
    mp.export_data()
    mp.erase(keys[:50].contiguous())
    k, v, st = mp.export_data(incremental=True)
    assert int(st.sum()) == 0 and k.numel() == 50                            # a full export of what is left
    k, v, st = mp.export_data(incremental=True)
    assert k.numel() == 0
