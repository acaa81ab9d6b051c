"""CUDA path (through the aot C-ABI) against the committed golden fixtures."""
import os

import numpy as np
import pytest
import torch

from mindrec_b200 import cells, ops

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _d(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


def test_lookup_dedup_and_updates_against_fixture(cuda):
    z = np.load(os.path.join(G, "lookup_update.npz"))
    ids, wts, table, wide = (_d(z[k], cuda) for k in ("ids", "wts", "table", "wide"))
    np.testing.assert_array_equal(ops.gather_masked(table, ids, wts).cpu().numpy(), z["gather_masked"])
    bias = torch.tensor([0.125], device=cuda)
    np.testing.assert_allclose(ops.gather_reduce(wide, ids, wts, bias).cpu().numpy(), z["gather_reduce"], rtol=1e-5, atol=1e-7)
    uq = ops.unique(ids, table_like=table)
    u = int(uq.count.item())
    np.testing.assert_array_equal(uq.uniq[:u].cpu().numpy(), z["uniq"])
    np.testing.assert_array_equal(uq.inverse.cpu().numpy(), z["inverse"])
    np.testing.assert_array_equal(uq.perm.cpu().numpy(), z["perm"])
    np.testing.assert_array_equal(uq.seg_start[:u + 1].cpu().numpy(), z["seg_start"])
    uf, inv_f, cnt = ops.unique_first(ids)
    np.testing.assert_array_equal(uf[:int(cnt.item())].cpu().numpy(), z["uniq_first"])
    np.testing.assert_array_equal(inv_f.cpu().numpy(), z["inverse_first"])
    gs = ops.segment_sum(_d(z["g"], cuda), wts.reshape(-1), uq)
    np.testing.assert_allclose(gs[:u].cpu().numpy(), z["gsum"], rtol=1e-5, atol=1e-5 * np.abs(z["gsum"]).max())
    w, m, v = table.clone(), torch.zeros_like(table), torch.zeros_like(table)
    hyper = ops.adam_hyper(3.5e-4, eps=1e-8, loss_scale=1024.0, device=cuda)
    ops.adam_begin_step(hyper)
    ops.sparse_lazy_adam(w, m, v, hyper, _d(z["g"], cuda), wts.reshape(-1), uq)
    np.testing.assert_allclose(w.cpu().numpy(), z["adam_w"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(m.cpu().numpy(), z["adam_m"], rtol=1e-4, atol=1e-5 * np.abs(z["adam_m"]).max())
    ww, acc, lin = wide.clone(), torch.ones_like(wide), torch.zeros_like(wide)
    ops.sparse_ftrl(ww, acc, lin, ops.ftrl_hyper(5e-2, 1e-8, 1e-8, loss_scale=1024.0, device=cuda),
                    _d(z["gw"], cuda), wts.reshape(-1), uq)
    np.testing.assert_allclose(ww.cpu().numpy(), z["ftrl_w"], rtol=1e-5, atol=1e-6 * np.abs(z["ftrl_w"]).max())
    np.testing.assert_allclose(acc.cpu().numpy(), z["ftrl_acc"], rtol=1e-5)


def test_interaction_against_fixture(cuda):
    z = np.load(os.path.join(G, "interaction.npz"))
    vx = _d(z["vx"], cuda)
    np.testing.assert_allclose(ops.fm_fwd(vx).cpu().numpy(), z["fm"], rtol=1e-5, atol=1e-5 * 0.5 * np.square(z["vx"]).sum((1, 2)).max())
    np.testing.assert_allclose(ops.fm_bwd(vx, _d(z["gout"], cuda)).cpu().numpy(), z["dvx"], rtol=1e-5, atol=1e-6)
    x0, cw, cb = (_d(z[k], cuda) for k in ("x0", "cw", "cb"))
    y, p = ops.cross_fwd(x0, cw, cb)
    np.testing.assert_allclose(y.cpu().numpy(), z["y"], rtol=1e-5, atol=1e-6)
    dx, dw, db = ops.cross_bwd(x0, _d(z["gy"], cuda), cw, cb, p)
    np.testing.assert_allclose(dx.cpu().numpy(), z["dx"], rtol=1e-4, atol=1e-5 * np.abs(z["dx"]).max())
    np.testing.assert_allclose(dw.cpu().numpy(), z["dw"], rtol=1e-4, atol=1e-5 * np.abs(z["dw"]).max())
    np.testing.assert_allclose(db.cpu().numpy(), z["db"], rtol=1e-4, atol=1e-5 * np.abs(z["db"]).max())


@pytest.mark.parametrize("mode", ["lazy", "adam", "dense"])
def test_wide_deep_steps_against_fixture(cuda, mode):
    z = np.load(os.path.join(G, "wide_deep_steps.npz"))
    vocab, dim = z["table"].shape
    cfg = cells.WideDeepConfig(batch_size=z["ids0"].shape[0], vocab_size=vocab, emb_dim=dim, deep_layer_dim=(16, 8),
                               use_mixed_precision=False, sparse=(mode != "dense"))
    model = cells.WideDeepModel(cfg, device=cuda)
    model.wide_embeddinglookup.embedding_table.data.copy_(_d(z["wide"], cuda))
    model.deep_embeddinglookup.embedding_table.data.copy_(_d(z["table"], cuda))
    for i in range(3):
        model.dense.weights[i].copy_(_d(z["mlp_w%d" % i], cuda))
        model.dense.biases[i].copy_(_d(z["mlp_b%d" % i], cuda))
    model.wide_b.data.copy_(_d(z["wide_b"], cuda))
    step = cells.TrainStepWrap(cells.NetWithLossClass(model, cfg), sparse=cfg.sparse, lazy_adam=(mode == "lazy"))
    for i in range(3):
        lw, ld = step(_d(z["ids%d" % i], cuda), _d(z["wts%d" % i], cuda), _d(z["label%d" % i], cuda))
        np.testing.assert_allclose(float(lw), z["%s_loss_w" % mode][i], rtol=1e-5)
        np.testing.assert_allclose(float(ld), z["%s_loss_d" % mode][i], rtol=1e-5)
    for got, key in ((model.embedding_table.data, "%s_deep" % mode),
                     (model.wide_embeddinglookup.embedding_table.data, "%s_wide" % mode),
                     (model.dense.weights[0], "%s_mlp_w0" % mode)):
        ref = z[key]
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=2e-4, atol=2e-5 * np.abs(ref).max())


def test_multitable_steps_against_fixture(cuda):
    """The CUDA multitable cell against bytes that travel with the repository (tests/golden/multitable_steps.npz)."""
    import importlib.util
    from mindrec_b200 import multitable as MT
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(G, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    c = mod.MT_SHAPES
    z = np.load(os.path.join(G, "multitable_steps.npz"))
    cfg = MT.MultitableConfig(batch_size=c["batch"], n_indicator=c["n_indicator"], n_emb128=c["n_emb128"],
                              n_emb64_single=c["n_emb64_single"], multi_slots=c["multi_slots"],
                              continue_field_size=c["continue_fields"], emb_128_size=c["emb_128_size"],
                              emb64_single_size=c["emb64_single_size"], emb64_multi_size=c["emb64_multi_size"],
                              indicator_size=c["indicator_size"], deep_dim_list=c["hidden"], use_mixed_precision=False)
    model = MT.MultitableWideDeepModel(cfg, device=cuda)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    for k, v in list(model.deep_tables().items()) + list(model.wide_tables().items()):
        v.copy_(t(z["init_" + k]).view(v.shape))
    for i, (w, b) in enumerate(zip(model.dense.weights, model.dense.biases)):
        w.copy_(t(z["mlp_w%d" % i]))
        b.copy_(t(z["mlp_b%d" % i]))
    step = MT.TrainStepWrap(MT.NetWithLossClass(model, cfg), cfg, sens=1000.0)
    for s in range(2):
        args = (t(z["cont%d" % s]), t(z["ind%d" % s]), t(z["e128_%d" % s]), t(z["e64_%d" % s]),
                [t(z["multi_ids%d_%d" % (s, k)]) for k in range(6)], [t(z["multi_mask%d_%d" % (s, k)]) for k in range(6)])
        if s == 0:
            np.testing.assert_allclose(model(*args).cpu().numpy(), z["logit0"], rtol=1e-5, atol=1e-6)
        loss, _ = step(t(z["label%d" % s]), *args)
        np.testing.assert_allclose(float(loss), z["loss"][s], rtol=1e-5)
    for k, v in list(model.deep_tables().items()) + list(model.wide_tables().items()):
        np.testing.assert_allclose(v.cpu().numpy().reshape(z["final_" + k].shape), z["final_" + k], rtol=2e-4, atol=2e-6,
                                   err_msg=k)
    np.testing.assert_allclose(model.dense.weights[0].cpu().numpy(), z["final_mlp_w0"], rtol=2e-4, atol=2e-6)
