"""Row-sharded path on the GPU (single rank: degenerate splits, real kernels) vs the unsharded cell."""
import numpy as np
import pytest
import torch

from mindrec_b200 import cells, sharded, synth

pytestmark = pytest.mark.gpu


def test_one_rank_sharded_step_equals_unsharded_cell(cuda):
    vocab, dim, b, hidden = 5000, 16, 300, (64, 32)
    step = sharded.ShardedWideDeepStep(b, vocab, dim, hidden, cuda, seed=3, use_mixed_precision=False)
    cfg = cells.WideDeepConfig(batch_size=b, vocab_size=vocab, emb_dim=dim, deep_layer_dim=hidden,
                               use_mixed_precision=False, sparse=True, seed=9)
    model = cells.WideDeepModel(cfg, device=cuda)
    model.wide_embeddinglookup.embedding_table.data.copy_(step.tables.wide[:vocab])
    model.deep_embeddinglookup.embedding_table.data.copy_(step.tables.deep[:vocab])
    model.dense.flat.copy_(step.dense.flat)
    ref = cells.TrainStepWrap(cells.NetWithLossClass(model, cfg), sparse=True, lazy_adam=True)
    gen = synth.CriteoSynth(b, cards=[80] * 26, vocab_pad=vocab, seed=4)
    for _ in range(3):
        ids, wts, label = (torch.from_numpy(x).to(cuda) for x in gen.next())
        l1, _ = step(ids, wts, label)
        l2, _ = ref(ids, wts, label)
        np.testing.assert_allclose(float(l1), float(l2), rtol=1e-6)
    wide, deep = step.tables.gather_full()
    torch.testing.assert_close(deep, model.embedding_table.data, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(wide, model.wide_embeddinglookup.embedding_table.data, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(step.dense.flat, model.dense.flat, rtol=1e-5, atol=1e-8)


def test_shard_bounds_kernel(cuda):
    from mindrec_b200 import ops
    uniq = torch.tensor([1, 5, 5 + 100, 250, 399, 0, 0, 0], dtype=torch.int32, device=cuda)
    count = torch.tensor([5], dtype=torch.int32, device=cuda)
    edges = torch.tensor([0, 100, 200, 300, 400], dtype=torch.int32, device=cuda)
    assert ops.shard_bounds(uniq, count, edges).tolist() == [0, 2, 3, 4, 5]


def test_shard_remap_kernel_matches_plan(cuda):
    from mindrec_b200 import ops
    plan = sharded.ShardPlan(1000, 8)
    ids = torch.arange(-3, 1005, dtype=torch.int32, device=cuda)
    got = ops.shard_remap(ids, torch.empty((1000, 0), device=cuda), torch.empty((8, plan.rows_per_rank, 0), device=cuda))
    assert torch.equal(got, plan.remap(ids))
