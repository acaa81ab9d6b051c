"""Wide&Deep multitable cell vs the float64 oracle: forward logit and every parameter after two
training steps, plus the dense-gradient scatter-add kernel it is built on."""
import numpy as np
import pytest
import torch

from mindrec_b200 import multitable as MT
from mindrec_b200 import ops
from oracle import ref_numpy as R

pytestmark = pytest.mark.gpu


def _inputs(cfg, rng, cuda):
    b = cfg.batch_size
    cont = rng.random((b, cfg.continue_field_size)).astype(np.float32)
    ind = rng.integers(0, cfg.indicator_size, size=(b, cfg.n_indicator)).astype(np.int32)
    e128 = rng.integers(0, cfg.emb_128_size, size=(b, cfg.n_emb128)).astype(np.int32)
    e64 = rng.integers(0, cfg.emb64_single_size, size=(b, cfg.n_emb64_single)).astype(np.int32)
    multi = []
    for s in cfg.multi_slots:
        ids = rng.integers(0, min(cfg.emb64_multi_size, 40), size=(b, s)).astype(np.int32)   # heavy key re-use
        mask = (rng.random((b, s)) < 0.7).astype(np.float32)
        multi.append((ids, mask))
    label = (rng.random(b) < 0.25).astype(np.float32)
    t = lambda a: torch.from_numpy(a).to(cuda)
    dev = (t(label), t(cont), t(ind), t(e128), t(e64), [t(i) for i, _ in multi], [t(m) for _, m in multi])
    return (label, cont, ind, e128, e64, multi), dev


def test_multitable_step_matches_oracle(cuda):
    cfg = MT.MultitableConfig(batch_size=96, n_indicator=2, n_emb128=3, n_emb64_single=2, multi_slots=(3, 5, 2, 4, 1, 6),
                              continue_field_size=8, emb_128_size=500, emb64_single_size=300, emb64_multi_size=200,
                              indicator_size=16, deep_dim_list=(64, 32, 32), use_mixed_precision=False, seed=4)
    model = MT.MultitableWideDeepModel(cfg, device=cuda)
    for w in model.dense.weights:
        w.mul_(10.0)                                    # logits of order 1
    cpu = lambda t: t.detach().cpu().numpy()
    oracle = R.MultitableOracle({k: cpu(v) for k, v in model.deep_tables().items()},
                                {k: cpu(v) for k, v in model.wide_tables().items()},
                                [cpu(w) for w in model.dense.weights], [cpu(b) for b in model.dense.biases])
    step = MT.TrainStepWrap(MT.NetWithLossClass(model, cfg), cfg, sens=1000.0)
    rng = np.random.default_rng(0)
    for it in range(2):
        host, dev = _inputs(cfg, rng, cuda)
        if it == 0:
            logit = cpu(model(*dev[1:]))
            want, _ = oracle.forward(*host[1:])
            np.testing.assert_allclose(logit, want, rtol=1e-5, atol=1e-6)
        loss, _ = step(*dev)
        want_loss = oracle.step(*host)
        np.testing.assert_allclose(float(loss), float(want_loss), rtol=1e-5)
    # gradients reach the tables through sums of ~100 terms: compare the accumulated update, scaled to the weights
    for k, v in list(model.deep_tables().items()) + list(model.wide_tables().items()):
        ref = oracle.deep.get(k, oracle.wide.get(k))
        np.testing.assert_allclose(cpu(v), ref.reshape(v.shape), rtol=2e-4, atol=2e-6, err_msg=k)
    for i, w in enumerate(model.dense.weights):
        np.testing.assert_allclose(cpu(w), oracle.mlp_w[i], rtol=2e-4, atol=2e-6)


def test_multitable_mixed_precision_runs_and_learns(cuda):
    cfg = MT.MultitableConfig(batch_size=256, n_indicator=2, n_emb128=2, n_emb64_single=2, multi_slots=(2, 2, 2, 2, 2, 2),
                              emb_128_size=2000, emb64_single_size=300, emb64_multi_size=200, deep_dim_list=(64, 64),
                              use_mixed_precision=True, seed=5)
    model = MT.MultitableWideDeepModel(cfg, device=cuda)
    step = MT.TrainStepWrap(MT.NetWithLossClass(model, cfg), cfg)
    rng = np.random.default_rng(1)
    _, dev = _inputs(cfg, rng, cuda)
    losses = [float(step(*dev)[0]) for _ in range(30)]
    assert np.isfinite(losses).all() and losses[-1] < 0.8 * losses[0]


@pytest.mark.parametrize("dim,div", [(64, 1), (64, 5), (1, 7), (6, 1)])
def test_segment_sum_scatter_add_accumulates_a_dense_gradient(cuda, dim, div):
    rng = np.random.default_rng(dim + div)
    vocab, b = 300, 64
    table = torch.zeros((vocab, dim), device=cuda)
    want = np.zeros((vocab, dim))
    for call in range(3):                               # several inputs looking up one table
        ids = rng.integers(0, 50, size=(b, div)).astype(np.int32)
        vals = rng.standard_normal((b, dim)).astype(np.float32)
        mask = rng.random(b * div).astype(np.float32)
        uq = ops.unique(torch.from_numpy(ids.reshape(-1)).to(cuda), table_like=table)
        ops.segment_sum_scatter_add(table, torch.from_numpy(vals).to(cuda), torch.from_numpy(mask).to(cuda), uq)
        want += R._scatter_rows((vocab, dim), ids, vals, mask, div=div)
    np.testing.assert_allclose(table.cpu().numpy(), want, rtol=1e-5, atol=1e-5)
