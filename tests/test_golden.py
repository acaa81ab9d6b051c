"""The oracle against the frozen vectors in tests/golden/ (CPU only).  See tests/golden/make_golden.py for
what the fixtures are (and are not): the reference ships no golden vectors for this path."""
import os

import numpy as np

from oracle import ref_c as C
from oracle import ref_numpy as R

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_reproduces_lookup_update_fixture():
    z = np.load(os.path.join(G, "lookup_update.npz"))
    ids, wts, table, wide = z["ids"], z["wts"], z["table"], z["wide"]
    vocab = table.shape[0]
    np.testing.assert_array_equal(R.gather_masked(table, ids, wts), z["gather_masked"])
    np.testing.assert_array_equal(R.gather_reduce(wide, ids, wts, np.array([0.125], np.float32)), z["gather_reduce"])
    uniq, inverse, perm, seg_start = R.unique_sorted(ids, bound=vocab)
    for a, b in ((uniq, "uniq"), (inverse, "inverse"), (perm, "perm"), (seg_start, "seg_start")):
        np.testing.assert_array_equal(a, z[b])
    assert uniq[-1] == vocab          # the out-of-range id collapsed onto V
    uf, inv_f = R.unique_first(ids)
    np.testing.assert_array_equal(uf, z["uniq_first"])
    np.testing.assert_array_equal(inv_f, z["inverse_first"])
    np.testing.assert_allclose(R.segment_sum(z["g"], inverse, uniq.size, wts.reshape(-1)), z["gsum"], rtol=1e-6)
    w, m, v = table.copy(), np.zeros_like(table), np.zeros_like(table)
    st = R.AdamState(3.5e-4, eps=1e-8, loss_scale=1024.0)
    st.begin_step()
    R.lazy_adam_sparse(w, m, v, uniq, z["gsum"], st)
    np.testing.assert_allclose(w, z["adam_w"], rtol=1e-6)
    np.testing.assert_allclose(v, z["adam_v"], rtol=1e-6)


def test_c_port_reproduces_lookup_update_fixture():
    z = np.load(os.path.join(G, "lookup_update.npz"))
    ids, wts, table = z["ids"], z["wts"], z["table"]
    np.testing.assert_array_equal(C.gather_masked(table, ids, wts), z["gather_masked"])
    uniq, inverse, perm, seg_start = C.unique(ids, table.shape[0])
    for a, b in ((uniq, "uniq"), (inverse, "inverse"), (perm, "perm"), (seg_start, "seg_start")):
        np.testing.assert_array_equal(a, z[b])
    gs = C.segment_sum(z["g"], table.shape[1], 1, np.ascontiguousarray(wts.reshape(-1)), perm, seg_start)
    np.testing.assert_allclose(gs, z["gsum"], rtol=1e-5, atol=1e-3)


def test_oracle_reproduces_interaction_fixture():
    z = np.load(os.path.join(G, "interaction.npz"))
    np.testing.assert_allclose(R.fm_forward(z["vx"]), z["fm"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(R.fm_backward(z["vx"], z["gout"]), z["dvx"], rtol=1e-6, atol=1e-6)
    y, _, s = R.cross_forward(z["x0"], z["cw"], z["cb"])
    np.testing.assert_allclose(y, z["y"], rtol=1e-6, atol=1e-6)
    dx, dw, db = R.cross_backward(z["x0"], z["cw"], z["cb"], z["gy"])
    np.testing.assert_allclose(dx, z["dx"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(dw, z["dw"], rtol=1e-6, atol=1e-6)


def test_oracle_reproduces_wide_deep_steps_fixture():
    z = np.load(os.path.join(G, "wide_deep_steps.npz"))
    mlp_w = [z["mlp_w%d" % i] for i in range(3)]
    mlp_b = [z["mlp_b%d" % i] for i in range(3)]
    for mode in ("lazy", "adam", "dense"):
        orc = R.WideDeepOracle(z["wide"], z["table"], mlp_w, mlp_b, z["wide_b"], mode=mode)
        for i in range(3):
            lw, ld = orc.step(z["ids%d" % i], z["wts%d" % i], z["label%d" % i].astype(np.float64))
            np.testing.assert_allclose(lw, z["%s_loss_w" % mode][i], rtol=1e-6)
            np.testing.assert_allclose(ld, z["%s_loss_d" % mode][i], rtol=1e-6)
        np.testing.assert_allclose(orc.wd, z["%s_deep" % mode], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(orc.ww, z["%s_wide" % mode], rtol=1e-6, atol=1e-9)


def _multitable_fixture():
    sys_path = os.path.join(G, "make_golden.py")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", sys_path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return np.load(os.path.join(G, "multitable_steps.npz")), mod.MT_SHAPES


DEEP_NAMES = ("emb128_embedding", "emb64_single", "emb64_multi", "emb64_indicator")
WIDE_NAMES = ("wide_continue_w", "wide_emb128_w", "wide_emb64_single_w", "wide_emb64_multi_w", "wide_indicator_w", "wide_bias")


def test_oracle_reproduces_multitable_fixture():
    z, c = _multitable_fixture()
    n_layers = len(c["hidden"]) + 1
    orc = R.MultitableOracle({k: z["init_" + k] for k in DEEP_NAMES}, {k: z["init_" + k] for k in WIDE_NAMES},
                             [z["mlp_w%d" % i] for i in range(n_layers)], [z["mlp_b%d" % i] for i in range(n_layers)])
    for step in range(2):
        multi = [(z["multi_ids%d_%d" % (step, k)], z["multi_mask%d_%d" % (step, k)]) for k in range(6)]
        args = (z["cont%d" % step], z["ind%d" % step], z["e128_%d" % step], z["e64_%d" % step], multi)
        if step == 0:
            np.testing.assert_allclose(orc.forward(*args)[0], z["logit0"], rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(orc.step(z["label%d" % step], *args), z["loss"][step], rtol=1e-6)
    for k in DEEP_NAMES:
        np.testing.assert_allclose(orc.deep[k], z["final_" + k], rtol=1e-6, atol=1e-9)
    for k in WIDE_NAMES:
        np.testing.assert_allclose(orc.wide[k], z["final_" + k], rtol=1e-6, atol=1e-9)
