"""DeepFM and Deep&Cross training steps (BASELINE configs 3/4, scaled down) vs the numpy oracle."""
import numpy as np
import pytest
import torch

from mindrec_b200 import interaction as I
from mindrec_b200 import synth
from oracle import ref_numpy as R

pytestmark = pytest.mark.gpu


def _close(got, ref, rtol=2e-4, atol_rel=2e-5):
    np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=rtol, atol=atol_rel * max(np.abs(ref).max(), 1e-12))


def test_deepfm_train_step_matches_oracle(cuda):
    cfg = I.DeepFMConfig(batch_size=200, data_vocab_size=1500, data_emb_dim=16,
                         deep_layer_args=((64, 32), "relu"), convert_dtype=False, seed=5)
    model = I.DeepFMModel(cfg, device=cuda)
    step = I.DeepFMTrainStep(I.DeepFMNetWithLoss(model, cfg.l2_coef), lr=cfg.learning_rate, eps=cfg.epsilon,
                             loss_scale=cfg.loss_scale)
    orc = R.DeepFMOracle(model.fm_w.data.cpu().numpy(), model.embedding_table.data.cpu().numpy(),
                         [w.cpu().numpy() for w in model.dense.weights],
                         [b.cpu().numpy() for b in model.dense.biases], lr=cfg.learning_rate, eps=cfg.epsilon,
                         l2_coef=cfg.l2_coef, sens=cfg.loss_scale)
    gen = synth.CriteoSynth(cfg.batch_size, cards=[30] * 26, vocab_pad=cfg.data_vocab_size, seed=1)
    for _ in range(3):
        ids, wts, label = gen.next()
        loss = step(*(torch.from_numpy(x).to(cuda) for x in (ids, wts, label)))
        ref = orc.step(ids, wts, label.astype(np.float64))
        np.testing.assert_allclose(float(loss), ref, rtol=1e-5)
    _close(model.fm_w.data, orc.w)
    _close(model.embedding_table.data, orc.v)
    for w, r in zip(model.dense.weights, orc.mlp_w):
        _close(w, r)


def test_deepfm_mixed_precision_runs(cuda):
    cfg = I.DeepFMConfig(batch_size=256, data_vocab_size=3000, data_emb_dim=16, convert_dtype=True, seed=6)
    model = I.DeepFMModel(cfg, device=cuda)
    step = I.DeepFMTrainStep(I.DeepFMNetWithLoss(model, cfg.l2_coef), loss_scale=1024.0)
    gen = synth.CriteoSynth(cfg.batch_size, cards=[60] * 26, vocab_pad=cfg.data_vocab_size, seed=2)
    ids, wts, label = (torch.from_numpy(x).to(cuda) for x in gen.next())
    losses = [float(step(ids, wts, label)) for _ in range(20)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0]


def test_deep_cross_train_step_matches_oracle(cuda):
    cfg = I.DeepCrossConfig(batch_size=150, vocab_size=1200, emb_dim=8, deep_layer_dim=(48, 32),
                            cross_layer_num=6, seed=7)
    model = I.DeepCrossModel(cfg, device=cuda)
    step = I.DeepCrossTrainStep(I.DeepCrossNetWithLoss(model), lr=cfg.learning_rate, eps=cfg.epsilon,
                                loss_scale=cfg.loss_scale)
    n = lambda t: t.cpu().numpy()
    orc = R.DeepCrossOracle(n(model.embedding_table.data), [n(w) for w in model.tower.weights],
                            [n(b) for b in model.tower.biases], [n(w) for w in model.head.weights],
                            [n(b) for b in model.head.biases], n(model.cross_weight), n(model.cross_bias),
                            lr=cfg.learning_rate, eps=cfg.epsilon, sens=cfg.loss_scale)
    gen = synth.CriteoSynth(cfg.batch_size, cards=[25] * 26, vocab_pad=cfg.vocab_size, seed=3)
    for _ in range(3):
        ids, wts, label = gen.next()
        loss = step(*(torch.from_numpy(x).to(cuda) for x in (ids, wts, label)))
        ref = orc.step(ids, wts, label.astype(np.float64))
        np.testing.assert_allclose(float(loss), ref, rtol=1e-5)
    _close(model.embedding_table.data, orc.table)
    _close(model.cross_weight, orc.cw)
    _close(model.cross_bias, orc.cb)
    for w, r in zip(model.tower.weights, orc.tw):
        _close(w, r)
    _close(model.head.weights[0], orc.hw[0])


def test_cross_layer_cell_is_the_one_layer_stack(cuda):
    layer = I.CrossLayer(4, 40, device=cuda)
    x = torch.randn((9, 40), device=cuda)
    y = layer(x, x)
    ref = x * (x @ layer.cross_weight.data.view(-1, 1)) + layer.cross_bias.data + x
    torch.testing.assert_close(y, ref, rtol=1e-5, atol=1e-5)
