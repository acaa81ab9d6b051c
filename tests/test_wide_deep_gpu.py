"""Whole Wide&Deep training step (BASELINE config 1 shape, scaled down) vs the numpy oracle."""
import numpy as np
import pytest
import torch

from mindrec_b200 import cells, synth
from oracle import ref_numpy as R

pytestmark = pytest.mark.gpu


def _build(cuda, mode, vocab=3000, batch=257, emb=16, hidden=(64, 32)):
    cfg = cells.WideDeepConfig(batch_size=batch, vocab_size=vocab, emb_dim=emb, deep_layer_dim=hidden,
                               use_mixed_precision=False, sparse=(mode != "dense"), seed=7)
    model = cells.WideDeepModel(cfg, device=cuda)
    net = cells.NetWithLossClass(model, cfg)
    step = cells.TrainStepWrap(net, sens=1024.0, sparse=cfg.sparse, lazy_adam=(mode == "lazy"))
    oracle = R.WideDeepOracle(model.wide_embeddinglookup.embedding_table.data.cpu().numpy(),
                              model.deep_embeddinglookup.embedding_table.data.cpu().numpy(),
                              [w.cpu().numpy() for w in model.dense.weights],
                              [b.cpu().numpy() for b in model.dense.biases],
                              model.wide_b.data.cpu().numpy(), sens=1024.0, mode=mode, l2_coef=cfg.l2_coef)
    return cfg, model, step, oracle


@pytest.mark.parametrize("mode", ["lazy", "adam", "dense"])
def test_train_step_matches_oracle(cuda, mode):
    cfg, model, step, oracle = _build(cuda, mode)
    gen = synth.CriteoSynth(cfg.batch_size, cards=[50] * 26, vocab_pad=cfg.vocab_size, seed=3)
    for it in range(3):
        ids, wts, label = gen.next()
        lw, ld = step(torch.from_numpy(ids).to(cuda), torch.from_numpy(wts).to(cuda),
                      torch.from_numpy(label).to(cuda))
        rw, rd = oracle.step(ids, wts, label.astype(np.float64))
        np.testing.assert_allclose(float(lw), rw, rtol=1e-5)
        np.testing.assert_allclose(float(ld), rd, rtol=1e-5)
    pairs = [(model.wide_embeddinglookup.embedding_table.data, oracle.ww),
             (model.deep_embeddinglookup.embedding_table.data, oracle.wd),
             (step.optimizer_w.accum[0], oracle.acc), (step.optimizer_w.linear[0], oracle.lin),
             (step.optimizer_d.moment1[0], oracle.md), (step.optimizer_d.moment2[0], oracle.vd),
             (model.wide_b.data, oracle.wide_b)]
    pairs += [(w, r) for w, r in zip(model.dense.weights, oracle.mlp_w)]
    pairs += [(b, r) for b, r in zip(model.dense.biases, oracle.mlp_b)]
    for got, ref in pairs:
        # fp32 GEMMs (TF32 off) + fp32 segment sums vs float64: a few 1e-5 of the tensor's scale
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=2e-4, atol=2e-5 * max(np.abs(ref).max(), 1e-12))


def test_untouched_rows_do_not_move_under_lazy_adam(cuda):
    cfg, model, step, _ = _build(cuda, "lazy")
    before = model.embedding_table.data.clone()
    gen = synth.CriteoSynth(cfg.batch_size, cards=[50] * 26, vocab_pad=cfg.vocab_size, seed=5)
    ids, wts, label = gen.next()
    step(torch.from_numpy(ids).to(cuda), torch.from_numpy(wts).to(cuda), torch.from_numpy(label).to(cuda))
    touched = torch.zeros(cfg.vocab_size, dtype=torch.bool, device=cuda)
    touched[torch.from_numpy(ids).to(cuda).long().reshape(-1)] = True
    after = model.embedding_table.data
    assert torch.equal(after[~touched], before[~touched])
    assert not torch.equal(after[touched], before[touched])


def test_graph_replay_equals_eager(cuda):
    cfg, model, step, _ = _build(cuda, "lazy")
    cfg2, model2, step2, _ = _build(cuda, "lazy")
    gen = synth.CriteoSynth(cfg.batch_size, cards=[50] * 26, vocab_pad=cfg.vocab_size, seed=9)
    batches = [gen.next() for _ in range(4)]
    dev = [tuple(torch.from_numpy(x).to(cuda) for x in b) for b in batches]
    # capture() runs warm-up steps that mutate state: run the same steps eagerly on the twin
    step.capture(*dev[0], warmup=2)
    for _ in range(2):  # the capture itself records, it does not execute
        step2(*dev[0])
    for b in dev[1:]:
        step.replay(*b)
        step2(*b)
    torch.cuda.synchronize()
    assert torch.equal(model.embedding_table.data, model2.embedding_table.data)
    assert torch.equal(model.wide_embeddinglookup.embedding_table.data,
                       model2.wide_embeddinglookup.embedding_table.data)


def test_mixed_precision_step_runs_and_decreases_loss(cuda):
    cfg = cells.WideDeepConfig(batch_size=512, vocab_size=5000, emb_dim=80, use_mixed_precision=True,
                               sparse=True, seed=11)
    model = cells.WideDeepModel(cfg, device=cuda)
    step = cells.TrainStepWrap(cells.NetWithLossClass(model, cfg), sparse=True, lazy_adam=True)
    gen = synth.CriteoSynth(cfg.batch_size, cards=[100] * 26, vocab_pad=cfg.vocab_size, seed=13)
    ids, wts, label = (torch.from_numpy(x).to(cuda) for x in gen.next())
    losses = [float(step(ids, wts, label)[0]) for _ in range(30)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0]


def test_fused_sigmoid_xent_matches_oracle(cuda):
    from mindrec_b200 import ops
    rng = np.random.default_rng(3)
    b = 16000
    a = (rng.standard_normal((b, 1)) * 4).astype(np.float32)
    d = (rng.standard_normal((b, 1)) * 4).astype(np.float32)
    a[0], d[0] = 60.0, 60.0          # saturating logits must not overflow
    a[1], d[1] = -70.0, -50.0
    y = (rng.random((b, 1)) < 0.25).astype(np.float32)
    sens = torch.tensor([1024.0], device=cuda)
    logit, loss, delta, delta16, dsum = ops.sigmoid_xent(torch.from_numpy(a).to(cuda), torch.from_numpy(d).to(cuda),
                                                          torch.from_numpy(y).to(cuda), sens, half=True)
    x = a.astype(np.float64) + d.astype(np.float64)
    np.testing.assert_allclose(logit.cpu().numpy(), (a + d), rtol=0, atol=0)
    np.testing.assert_allclose(float(loss), R.sigmoid_xent(x, y).mean(), rtol=1e-5)
    ref_delta = 1024.0 * (R.sigmoid(x) - y) / b
    # sigmoid(x) - 1 cancels in fp32 for large x: compare at 1e-6 of the gradient scale (sens / B)
    np.testing.assert_allclose(delta.cpu().numpy(), ref_delta, rtol=1e-5, atol=1e-6 * 1024.0 / b)
    np.testing.assert_allclose(float(dsum), ref_delta.sum(), rtol=1e-4, atol=1e-6)
    np.testing.assert_array_equal(delta16.cpu().numpy(), delta.cpu().numpy().astype(np.float16))


def test_online_train_and_checkpoint_round_trip(cuda):
    """RecModel.online_train drives the W&D step from host batches; export -> import reproduces the state."""
    from mindrec_b200 import train
    cfg, model, step, _ = _build(cuda, "lazy")
    gen = synth.CriteoSynth(cfg.batch_size, cards=[50] * 26, vocab_pad=cfg.vocab_size, seed=21)

    def stream():
        while True:
            yield gen.next()
    params = train.RecModel(step, device=cuda).online_train(stream(), max_steps=4)
    assert params.cur_step_num == 4 and np.isfinite(float(params.net_outputs[0]))
    state = train.export_tables(step)
    cfg2, model2, step2, _ = _build(cuda, "lazy")
    train.import_tables(step2, state)
    batch = tuple(torch.from_numpy(x).to(cuda) for x in gen.next())
    step(*batch)
    step2(*batch)
    assert torch.equal(model.embedding_table.data, model2.embedding_table.data)
    assert torch.equal(model.dense.flat, model2.dense.flat)


def test_replay_with_staged_next_batch_equals_plain_replay(cuda):
    cfg, model, step, _ = _build(cuda, "lazy")
    cfg2, model2, step2, _ = _build(cuda, "lazy")
    gen = synth.CriteoSynth(cfg.batch_size, cards=[50] * 26, vocab_pad=cfg.vocab_size, seed=31)
    host = [tuple(torch.from_numpy(x).pin_memory() for x in gen.next()) for _ in range(5)]
    dev0 = tuple(t.to(cuda) for t in host[0])
    step.capture(*dev0, warmup=1)
    step2.capture(*dev0, warmup=1)
    for i in range(5):
        step.replay(*host[i], next_batch=host[(i + 1) % 5])
        step2.replay(*host[i])
    torch.cuda.synchronize()
    assert torch.equal(model.embedding_table.data, model2.embedding_table.data)
    assert torch.equal(model.dense.flat, model2.dense.flat)


def test_interleaved_adam_state_step_is_bit_identical(cuda):
    """WideDeepConfig(interleave_adam_state=True) changes where the deep table's w, m, v live, not what a step does."""
    outs = []
    for inter in (False, True):
        cfg = cells.WideDeepConfig(batch_size=257, vocab_size=3000, emb_dim=16, deep_layer_dim=(64, 32),
                                   use_mixed_precision=False, sparse=True, seed=7, interleave_adam_state=inter)
        model = cells.WideDeepModel(cfg, device=cuda)
        step = cells.TrainStepWrap(cells.NetWithLossClass(model, cfg), sens=1024.0, sparse=True, lazy_adam=True)
        gen = synth.CriteoSynth(cfg.batch_size, cards=[50] * 26, vocab_pad=cfg.vocab_size, seed=3)
        losses = []
        for it in range(3):
            ids, wts, label = gen.next()
            losses.append(float(step(torch.from_numpy(ids).to(cuda), torch.from_numpy(wts).to(cuda),
                                     torch.from_numpy(label).to(cuda))[0]))
        outs.append((losses, model.embedding_table.data.clone(), step.optimizer_d.moment1[0].clone(),
                     step.optimizer_d.moment2[0].clone(), model.dense.flat.clone()))
    assert outs[0][0] == outs[1][0]
    for a, b in zip(outs[0][1:], outs[1][1:]):
        assert torch.equal(a, b)
    with pytest.raises(ValueError, match="LazyAdam"):
        cfg = cells.WideDeepConfig(batch_size=8, vocab_size=100, emb_dim=8, deep_layer_dim=(8,), sparse=True,
                                   interleave_adam_state=True)
        cells.TrainStepWrap(cells.NetWithLossClass(cells.WideDeepModel(cfg, device=cuda), cfg), sparse=True)
