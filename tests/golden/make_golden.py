"""Generates tests/golden/*.npz — frozen, seeded input/output vectors of the oracle (oracle/ref_numpy.py).

The reference (mindspore-lab/mindrec) holds no golden vectors or known-answer tests for this path and MindSpore
cannot be imported here, so these fixtures do NOT come from running the reference: they freeze the oracle's
restatement so that (a) a later edit of the oracle cannot silently change its meaning (tests/test_golden.py
re-runs the oracle against them on CPU) and (b) the CUDA kernels are compared with bytes that travel with the
repository (tests/test_golden_gpu.py).  Re-generate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from mindrec_b200 import synth  # noqa: E402
from oracle import ref_numpy as R  # noqa: E402


def main():
    rng = np.random.default_rng(20260101)
    # --- lookup + dedup + sparse optimizers (Wide&Deep shapes, scaled down) ---
    vocab, dim, b = 997, 8, 48
    gen = synth.CriteoSynth(b, cards=[37] * 26, vocab_pad=vocab, seed=20260101)
    ids, wts, label = gen.next()
    ids[0, 5] = vocab + 3          # one out-of-range id
    table = (rng.standard_normal((vocab, dim)) * 0.01).astype(np.float32)
    wide = (rng.standard_normal((vocab, 1)) * 0.01).astype(np.float32)
    g = (rng.standard_normal((b * 39, dim)) * 1024).astype(np.float32)
    gw = (rng.standard_normal((b, 1)) * 1024).astype(np.float32)
    uniq, inverse, perm, seg_start = R.unique_sorted(ids, bound=vocab)
    ufirst, ifirst = R.unique_first(ids)
    gsum = R.segment_sum(g, inverse, uniq.size, wts.reshape(-1))
    gsum_w = R.segment_sum(gw, inverse, uniq.size, wts.reshape(-1), div=39)
    w, m, v = table.copy(), np.zeros_like(table), np.zeros_like(table)
    st = R.AdamState(3.5e-4, eps=1e-8, loss_scale=1024.0)
    st.begin_step()
    R.lazy_adam_sparse(w, m, v, uniq, gsum, st)
    ww, acc, lin = wide.copy(), np.ones_like(wide), np.zeros_like(wide)
    R.ftrl_sparse(ww, acc, lin, uniq, gsum_w, R.FtrlState(5e-2, 1e-8, 1e-8, loss_scale=1024.0))
    np.savez_compressed(os.path.join(HERE, "lookup_update.npz"), ids=ids, wts=wts, table=table, wide=wide, g=g, gw=gw,
                        gather_masked=R.gather_masked(table, ids, wts),
                        gather_reduce=R.gather_reduce(wide, ids, wts, np.array([0.125], np.float32)),
                        uniq=uniq, inverse=inverse, perm=perm, seg_start=seg_start, uniq_first=ufirst,
                        inverse_first=ifirst, gsum=gsum.astype(np.float32), gsum_w=gsum_w.astype(np.float32),
                        adam_w=w, adam_m=m, adam_v=v, ftrl_w=ww, ftrl_acc=acc, ftrl_lin=lin)
    # --- FM and cross stack ---
    vx = (rng.standard_normal((9, 39, 16)) * 0.3).astype(np.float32)
    gout = rng.standard_normal((9, 1)).astype(np.float32)
    x0 = (rng.standard_normal((7, 120)) * 0.1).astype(np.float32)
    cw = (rng.standard_normal((6, 120)) * 0.05).astype(np.float32)
    cb = (rng.standard_normal((6, 120)) * 0.05).astype(np.float32)
    gy = rng.standard_normal((7, 120)).astype(np.float32)
    y, _, s = R.cross_forward(x0, cw, cb)
    dx, dw, db = R.cross_backward(x0, cw, cb, gy)
    np.savez_compressed(os.path.join(HERE, "interaction.npz"), vx=vx, gout=gout, fm=R.fm_forward(vx).astype(np.float32),
                        dvx=R.fm_backward(vx, gout).astype(np.float32), x0=x0, cw=cw, cb=cb, gy=gy,
                        y=y.astype(np.float32), s=s.astype(np.float32), dx=dx.astype(np.float32),
                        dw=dw.astype(np.float32), db=db.astype(np.float32))
    # --- full Wide&Deep steps (3 steps, all three optimizer modes) ---
    hidden = (16, 8)
    dims = [39 * dim] + list(hidden) + [1]
    mlp_w = [(rng.standard_normal((dims[i], dims[i + 1])) * 0.05).astype(np.float32) for i in range(len(dims) - 1)]
    mlp_b = [(rng.standard_normal(dims[i + 1]) * 0.05).astype(np.float32) for i in range(len(dims) - 1)]
    out = dict(table=table, wide=wide, wide_b=np.array([0.01], np.float32))
    for i, (a, c) in enumerate(zip(mlp_w, mlp_b)):
        out["mlp_w%d" % i], out["mlp_b%d" % i] = a, c
    gen = synth.CriteoSynth(b, cards=[37] * 26, vocab_pad=vocab, seed=7)
    batches = [gen.next() for _ in range(3)]
    for i, (bi, bw, bl) in enumerate(batches):
        out["ids%d" % i], out["wts%d" % i], out["label%d" % i] = bi, bw, bl
    for mode in ("lazy", "adam", "dense"):
        orc = R.WideDeepOracle(wide, table, mlp_w, mlp_b, out["wide_b"], mode=mode)
        losses = [orc.step(bi, bw, bl.astype(np.float64)) for bi, bw, bl in batches]
        out["%s_loss_w" % mode] = np.array([l[0] for l in losses], np.float32)
        out["%s_loss_d" % mode] = np.array([l[1] for l in losses], np.float32)
        out["%s_deep" % mode], out["%s_wide" % mode] = orc.wd, orc.ww
        out["%s_mlp_w0" % mode] = orc.mlp_w[0]
    np.savez_compressed(os.path.join(HERE, "wide_deep_steps.npz"), **out)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


MT_SHAPES = dict(batch=40, continue_fields=8, n_indicator=2, n_emb128=3, n_emb64_single=2, multi_slots=(3, 5, 2, 4, 1, 6),
                 emb_128_size=120, emb64_single_size=90, emb64_multi_size=60, indicator_size=16, hidden=(24, 16))


def multitable():
    """Wide&Deep multitable cell: two training steps of the oracle (python make_golden.py multitable)."""
    c = MT_SHAPES
    rng = np.random.default_rng(20260102)
    n = lambda *shape: (rng.standard_normal(shape) * 0.01).astype(np.float32)
    deep = {"emb128_embedding": n(c["emb_128_size"], 128), "emb64_single": n(c["emb64_single_size"], 64),
            "emb64_multi": n(c["emb64_multi_size"], 64), "emb64_indicator": n(c["indicator_size"], 64)}
    wide = {"wide_continue_w": n(c["continue_fields"]), "wide_emb128_w": n(c["emb_128_size"], 1),
            "wide_emb64_single_w": n(c["emb64_single_size"], 1), "wide_emb64_multi_w": n(c["emb64_multi_size"], 1),
            "wide_indicator_w": n(c["indicator_size"], 1), "wide_bias": n(1)}
    din = c["continue_fields"] + c["n_indicator"] * 64 + c["n_emb128"] * 128 + c["n_emb64_single"] * 64 + 6 * 64
    dims = [din] + list(c["hidden"]) + [1]
    mlp_w = [(rng.standard_normal((dims[i], dims[i + 1])) * 0.1).astype(np.float32) for i in range(len(dims) - 1)]
    mlp_b = [(rng.standard_normal(dims[i + 1]) * 0.01).astype(np.float32) for i in range(len(dims) - 1)]
    out = {}
    for k, v in list(deep.items()) + list(wide.items()):
        out["init_" + k] = v
    for i, (a, b_) in enumerate(zip(mlp_w, mlp_b)):
        out["mlp_w%d" % i], out["mlp_b%d" % i] = a, b_
    orc = R.MultitableOracle(deep, wide, mlp_w, mlp_b)
    b = c["batch"]
    losses = []
    for step in range(2):
        cont = rng.random((b, c["continue_fields"])).astype(np.float32)
        ind = rng.integers(0, c["indicator_size"], size=(b, c["n_indicator"])).astype(np.int32)
        e128 = rng.integers(0, c["emb_128_size"], size=(b, c["n_emb128"])).astype(np.int32)
        e64 = rng.integers(0, c["emb64_single_size"], size=(b, c["n_emb64_single"])).astype(np.int32)
        multi = [(rng.integers(0, 25, size=(b, s)).astype(np.int32), (rng.random((b, s)) < 0.7).astype(np.float32))
                 for s in c["multi_slots"]]
        label = (rng.random(b) < 0.25).astype(np.float32)
        if step == 0:
            out["logit0"] = orc.forward(cont, ind, e128, e64, multi)[0].astype(np.float32)
        losses.append(orc.step(label, cont, ind, e128, e64, multi))
        out.update({"label%d" % step: label, "cont%d" % step: cont, "ind%d" % step: ind, "e128_%d" % step: e128,
                    "e64_%d" % step: e64})
        for k, (ids, m) in enumerate(multi):
            out["multi_ids%d_%d" % (step, k)], out["multi_mask%d_%d" % (step, k)] = ids, m
    out["loss"] = np.array(losses, np.float32)
    for k, v in list(orc.deep.items()) + list(orc.wide.items()):
        out["final_" + k] = v
    out["final_mlp_w0"] = orc.mlp_w[0]
    np.savez_compressed(os.path.join(HERE, "multitable_steps.npz"), **out)
    print("wrote multitable_steps.npz")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "multitable":
        multitable()
    else:
        main()
        multitable()
