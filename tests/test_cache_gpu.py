"""host <-> device embedding cache (mindrec_b200.cache).  A Wide&Deep model whose tables live in
pinned host memory behind a device cache far smaller than the working set of the run trains to the same tables as the
model with resident tables (same kernels, same row math; only the segment-sum tile boundaries differ, hence 1e-6)."""
import numpy as np
import pytest
import torch

from mindrec_b200 import cache, cells, ops, synth

pytestmark = pytest.mark.gpu


def test_cached_lookup_serves_misses_from_the_host_and_writes_back_on_flush(cuda):
    vocab, dim = 5000, 8
    emb = cache.CachedEmbeddingLookup(vocab, dim, vocab_cache_size=600, param_init="normal", device=cuda)
    t = emb.embedding_table
    ref = t.host_values.clone()
    rng = np.random.default_rng(0)
    for it in range(8):
        ids = torch.from_numpy(rng.integers(0, vocab, size=(50, 8)).astype(np.int32)).to(cuda)
        out = emb(ids)
        assert torch.equal(out.cpu(), ref[ids.cpu().long()])                     # cold rows come from the host tier
        # "train": add 1 to the looked-up rows by slot (what an optimizer does through as_parameter())
        slots = torch.unique(emb.last_slots.reshape(-1).long())
        t.values[slots] += 1.0
        ref[torch.unique(ids.cpu().long())] += 1.0
    assert t.flushes >= 2 and t.misses > 600                                        # the working set did not fit
    assert torch.equal(t.full_table(), ref)                                         # nothing lost across flushes
    with pytest.raises(ValueError, match="does not fit"):
        emb(torch.zeros((2000,), dtype=torch.int32, device=cuda))
    with pytest.raises(IndexError):
        emb(torch.tensor([vocab + 3], dtype=torch.int32, device=cuda))


def test_wide_deep_with_embedding_cache_matches_resident_tables(cuda):
    kw = dict(batch_size=257, vocab_size=30000, emb_dim=16, deep_layer_dim=(64, 32), use_mixed_precision=False,
              sparse=True, parameter_server=True, seed=7)
    cfg_r = cells.WideDeepConfig(**kw)
    cfg_c = cells.WideDeepConfig(vocab_cache_size=12000, **kw)
    resident = cells.WideDeepModel(cfg_r, device=cuda)
    cached = cells.WideDeepModel(cfg_c, device=cuda)
    assert cached.cached and cached.dynamic
    # same initial state: copy the resident tables into the host tier
    cached.wide_embeddinglookup.embedding_table.host_values.copy_(resident.wide_embeddinglookup.embedding_table.data)
    cached.deep_embeddinglookup.embedding_table.host_values.copy_(resident.deep_embeddinglookup.embedding_table.data)
    cached.dense.flat.copy_(resident.dense.flat)
    step_r = cells.TrainStepWrap(cells.NetWithLossClass(resident, cfg_r), sparse=True, parameter_server=True)
    step_c = cells.TrainStepWrap(cells.NetWithLossClass(cached, cfg_c), sparse=True, parameter_server=True,
                                 cache_enable=True)
    assert step_r.lazy_adam and step_c.lazy_adam
    gen = synth.CriteoSynth(257, cards=[1000] * 26, alpha=0.0, vocab_pad=30000, seed=3)  # uniform ids: many distinct rows
    for it in range(5):
        b = tuple(torch.from_numpy(x).to(cuda) for x in gen.next())
        lr, lc = float(step_r(*b)[0]), float(step_c(*b)[0])
        np.testing.assert_allclose(lc, lr, rtol=1e-6)
    td, tw = cached.deep_embeddinglookup.embedding_table, cached.wide_embeddinglookup.embedding_table
    assert td.flushes >= 1 and tw.flushes >= 1
    pairs = [(td.full_table(), resident.embedding_table.data.cpu()),
             (tw.full_table(), resident.wide_embeddinglookup.embedding_table.data.cpu()),
             (td.host_arenas[0], step_r.optimizer_d.moment1[0].cpu()), (td.host_arenas[1], step_r.optimizer_d.moment2[0].cpu()),
             (tw.host_arenas[0], step_r.optimizer_w.accum[0].cpu()), (tw.host_arenas[1], step_r.optimizer_w.linear[0].cpu()),
             (cached.dense.flat.cpu(), resident.dense.flat.cpu())]
    for got, want in pairs:
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6 * float(want.abs().max()))
