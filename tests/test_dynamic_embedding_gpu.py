"""Wide&Deep built on two HashEmbeddingLookups (dynamic_embedding=True,
models/wide_deep/src/wide_and_deep.py:268-274) with LazyAdam + FTRL on the MapParameters (:415-430), and the growth of
the hash table underneath.

The oracle is the dense WideDeepOracle(mode="lazy") keyed through the id: with a constant default row and admission on
first sight a MapParameter is indistinguishable from a [V, D] table filled with that constant of which only the looked
up rows ever move (LazyAdam / sparse FTRL).  Resident rows and their optimizer state are compared key by key; keys
never looked up must not be resident."""
import numpy as np
import pytest
import torch

from mindrec_b200 import cells, hash as H, ops, synth
from oracle import ref_numpy as R

pytestmark = pytest.mark.gpu
DEEP0, WIDE0 = 0.03, 0.01


def _build(cuda, capacity, vocab=3000, batch=257, emb=16, hidden=(64, 32), auto_grow=True):
    cfg = cells.WideDeepConfig(batch_size=batch, vocab_size=vocab, emb_dim=emb, deep_layer_dim=hidden,
                               use_mixed_precision=False, sparse=True, dynamic_embedding=True, seed=7,
                               emb_init=DEEP0, hash_capacity=capacity, hash_auto_grow=auto_grow)
    model = cells.WideDeepModel(cfg, device=cuda)
    # the wide table's default row is its own constant (one emb_init drives both lookups in the reference)
    wt = model.wide_embeddinglookup.embedding_table
    wt.values[wt.capacity].fill_(WIDE0)
    net = cells.NetWithLossClass(model, cfg)
    step = cells.TrainStepWrap(net, sens=1024.0, sparse=True, dynamic_embedding=True)
    assert step.lazy_adam                                   # wide_and_deep.py:415-419
    oracle = R.WideDeepOracle(np.full((vocab, 1), WIDE0, np.float32), np.full((vocab, emb), DEEP0, np.float32),
                              [w.cpu().numpy() for w in model.dense.weights],
                              [b.cpu().numpy() for b in model.dense.biases],
                              model.wide_b.data.cpu().numpy(), sens=1024.0, mode="lazy", l2_coef=cfg.l2_coef)
    return cfg, model, step, oracle


def _rows_by_key(mp, arena):
    k, s = mp._export()
    return k.cpu().numpy(), ops.gather(arena, s.contiguous()).cpu().numpy()


def _close(got, ref):
    np.testing.assert_allclose(got, ref, rtol=2e-4, atol=2e-5 * max(np.abs(ref).max(), 1e-12))


@pytest.mark.parametrize("capacity", [1 << 14, 64], ids=["presized", "grows"])
def test_dynamic_embedding_train_step_matches_oracle(cuda, capacity):
    cfg, model, step, oracle = _build(cuda, capacity)
    gen = synth.CriteoSynth(cfg.batch_size, cards=[50] * 26, vocab_pad=cfg.vocab_size, seed=3)
    seen = set()
    for it in range(3):
        ids, wts, label = gen.next()
        seen |= set(ids.reshape(-1).tolist())
        lw, ld = step(torch.from_numpy(ids).to(cuda), torch.from_numpy(wts).to(cuda), torch.from_numpy(label).to(cuda))
        rw, rd = oracle.step(ids, wts, label.astype(np.float64))
        np.testing.assert_allclose(float(lw), rw, rtol=1e-5)
        np.testing.assert_allclose(float(ld), rd, rtol=1e-5)
    td, tw = model.deep_embeddinglookup.embedding_table, model.wide_embeddinglookup.embedding_table
    assert not td.overflowed and not tw.overflowed
    if capacity == 64:
        assert td.capacity > 64 and td.grown >= 1 and tw.grown >= 1
    want = np.asarray(sorted(seen), dtype=np.int64)
    md, vd = (td.arena(i) for i in step.optimizer_d._map_state[0])
    acc, lin = (tw.arena(i) for i in step.optimizer_w._map_state[0])
    for mp, arena, ref in ((td, td.values, oracle.wd), (td, md, oracle.md), (td, vd, oracle.vd),
                           (tw, tw.values, oracle.ww), (tw, acc, oracle.acc), (tw, lin, oracle.lin)):
        keys, rows = _rows_by_key(mp, arena)
        np.testing.assert_array_equal(keys, want)           # exactly the looked-up keys are resident
        _close(rows, ref[keys])
    # rows of keys that were never looked up stayed at their initial value in the oracle (LazyAdam / sparse FTRL)
    rest = np.setdiff1d(np.arange(cfg.vocab_size), want)
    assert np.all(oracle.wd[rest] == np.float32(DEEP0)) and np.all(oracle.ww[rest] == np.float32(WIDE0))
    # the default rows did not move
    assert torch.equal(td.values[td.capacity], torch.full((cfg.emb_dim,), DEEP0, device=cuda))
    assert torch.equal(tw.values[tw.capacity], torch.full((1,), WIDE0, device=cuda))
    for w, r in zip(model.dense.weights, oracle.mlp_w):
        _close(w.cpu().numpy(), r)


def test_dynamic_embedding_argument_checks(cuda):
    cfg = cells.WideDeepConfig(batch_size=8, vocab_size=100, emb_dim=8, deep_layer_dim=(8,), use_mixed_precision=False,
                               sparse=False, dynamic_embedding=True, hash_capacity=64)
    model = cells.WideDeepModel(cfg, device=cuda)
    with pytest.raises(ValueError, match="sparse=True"):
        cells.TrainStepWrap(cells.NetWithLossClass(model, cfg), dynamic_embedding=True)
    cfg.sparse = True
    with pytest.raises(ValueError, match="dynamic_embedding=True"):
        cells.TrainStepWrap(cells.NetWithLossClass(model, cfg), sparse=True)
    step = cells.TrainStepWrap(cells.NetWithLossClass(model, cfg), sparse=True, dynamic_embedding=True)
    ids = torch.zeros((8, 39), dtype=torch.int32, device=cuda)
    with pytest.raises(RuntimeError, match="hash_auto_grow"):
        step.capture(ids, torch.ones((8, 39), device=cuda), torch.zeros((8, 1), device=cuda))


@pytest.mark.parametrize("kdt", [torch.int32, torch.int64])
def test_grow_keeps_keys_rows_state_and_filters(cuda, kdt):
    """grow(): resident keys, candidates (sightings < permit), their admission / eviction words and the rows of every
    arena move; tombstones vanish; lookups after the growth behave like the dict model that never noticed."""
    rng = np.random.default_rng(4)
    dim = 8
    mp = H.MapParameter(key_dtype=kdt, value_shape=dim, default_value=0.5, permit_filter_value=2, evict_filter_value=4,
                        capacity=256, device=cuda)
    side = mp.add_arena(7.0)
    model = R.MapParameterModel(dim, default_value=0.5, permit_filter_value=2, evict_filter_value=4)
    npdt = np.int32 if kdt == torch.int32 else np.int64
    space = 400 if kdt == torch.int32 else 2 ** 40
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    for it in range(8):
        keys = (rng.integers(0, 400, size=120) * (1 if kdt == torch.int32 else 2 ** 31 + 11) % space).astype(npdt)
        got = mp.get(t(keys)).cpu().numpy()
        np.testing.assert_array_equal(got, model.get(keys))
        if it == 2:
            ek = np.unique(keys[:20])
            mp.erase(t(ek))
            model.erase(ek)
        if it in (1, 4):
            before_k, before_rows = _rows_by_key(mp, mp.values)
            _, before_side = _rows_by_key(mp, mp.arena(0))
            occupied = int(mp.state[4].item())
            mp.grow()
            assert mp.capacity == 256 << (1 if it == 1 else 2)
            st = mp.state.tolist()
            assert st[2] == 0 and st[4] == occupied and st[0] == before_k.size      # no tombstones, same occupancy
            k2, rows2 = _rows_by_key(mp, mp.values)
            np.testing.assert_array_equal(k2, before_k)
            np.testing.assert_array_equal(rows2, before_rows)
            np.testing.assert_array_equal(_rows_by_key(mp, mp.arena(0))[1], before_side)
            assert torch.equal(mp.values[mp.capacity], torch.full((dim,), 0.5, device=cuda))
            assert torch.equal(mp.arena(0)[mp.capacity], torch.full((dim,), 7.0, device=cuda))
        if it % 3 == 2:
            mp.evict()
            model.evict()
        np.testing.assert_array_equal(mp.get_keys().cpu().numpy().astype(np.int64), model.keys())
    assert not mp.overflowed


def test_maybe_grow_prevents_the_overflow_flag(cuda):
    emb = H.HashEmbeddingLookup(4, key_dtype=torch.int64, param_init="zeros", capacity=64, device=cuda)
    for it in range(5):
        keys = torch.arange(it * 500, (it + 1) * 500, dtype=torch.int64, device=cuda) * 7919
        out = emb(keys)
        assert out.shape == (500, 4)
    mp = emb.embedding_table
    assert not mp.overflowed and len(mp) == 2500 and mp.capacity >= 4096
    assert int(mp.state[4].item()) <= mp.MAX_LOAD * mp.capacity
