"""BASELINE north_star "no PyTorch": the embedding path driven through mindrec_b200.runtime (ctypes over the library's
own mrec_rt_* entry points) in a fresh interpreter that must never import torch.  Parity against the numpy oracle as in
the torch-backed tests: bit-exact gathers / dedup, 1e-5 optimizer rows; plus CUDA-graph replay and a pinned H2D copy."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent('''
    import sys
    import traceback


    class NoTorch:                                   # any attempt to import torch fails loudly, with the culprit's stack
        def find_spec(self, name, path=None, target=None):
            if name == "torch" or name.startswith("torch."):
                traceback.print_stack()
                raise ImportError("the torch-free path tried to import " + name)
            return None


    sys.meta_path.insert(0, NoTorch())
    import numpy as np
    from mindrec_b200 import _lib, ops, runtime, synth
    from oracle import ref_numpy as R

    dev = runtime.Device(0)
    rng = np.random.default_rng(0)
    vocab, dim, b, f = 20011, 80, 512, 39
    gen = synth.CriteoSynth(b, cards=[700] * 26, vocab_pad=vocab, seed=1)
    ids, wts, label = gen.next()
    w = (rng.standard_normal((vocab, dim)) * 0.01).astype(np.float32)
    ww = (rng.standard_normal((vocab, 1)) * 0.01).astype(np.float32)
    d_w, d_ww, d_ids, d_wts = dev.from_numpy(w), dev.from_numpy(ww), dev.from_numpy(ids), dev.from_numpy(wts)
    before = _lib.launch_count()

    # lookups: bit-exact
    np.testing.assert_array_equal(ops.gather(d_w, d_ids).numpy(), R.gather(w, ids))
    np.testing.assert_array_equal(ops.gather_masked(d_w, d_ids, d_wts).numpy(), R.gather_masked(w, ids, wts))
    got16 = ops.gather_masked(d_w, d_ids, d_wts, out_dtype="float16").numpy()
    np.testing.assert_array_equal(got16, R.gather_masked(w, ids, wts).astype(np.float16))
    bias = dev.tensor([0.25])
    np.testing.assert_allclose(ops.gather_reduce(d_ww, d_ids, d_wts, bias).numpy(), R.gather_reduce(ww, ids, wts, 0.25), rtol=1e-5, atol=1e-7)

    # dedup: bit-exact indices
    uq = ops.unique(d_ids, table_like=d_w)
    uniq, inverse, perm, seg_start = R.unique_sorted(ids, bound=vocab)
    u = uq.count.item()
    assert u == uniq.size
    np.testing.assert_array_equal(uq.uniq.numpy()[:u], uniq)
    np.testing.assert_array_equal(uq.inverse.numpy(), inverse)
    np.testing.assert_array_equal(uq.perm.numpy(), perm)

    # fused sparse LazyAdam + FTRL, two steps, then the same two steps replayed from a captured CUDA graph
    g = (np.abs(rng.standard_normal((b * f, dim))) + 0.5).astype(np.float32)
    delta = ((np.abs(rng.standard_normal((b, 1))) + 0.5) / b).astype(np.float32)
    d_g, d_delta = dev.from_numpy(g), dev.from_numpy(delta)
    mask = d_wts.reshape(-1)

    def run(table, m, v, wide, acc, lin, ha, hf, steps):
        for _ in range(steps):
            ops.adam_begin_step(ha)
            ops.sparse_lazy_adam(table, m, v, ha, d_g, mask, uq)
            ops.sparse_ftrl(wide, acc, lin, hf, d_delta, mask, uq)

    def state():
        return (dev.from_numpy(w), dev.zeros((vocab, dim)), dev.zeros((vocab, dim)), dev.from_numpy(ww),
                dev.full((vocab, 1), 1.0), dev.zeros((vocab, 1)),
                ops.adam_hyper(3.5e-4, eps=1e-8, loss_scale=4.0, device=dev), ops.ftrl_hyper(5e-2, 1e-8, 1e-8, loss_scale=4.0, device=dev))

    eager = state()
    run(*eager, steps=2)
    graphed = state()
    run(*graphed, steps=0)
    with dev.capture() as graph:
        run(*graphed, steps=1)
    graph.launch()
    graph.launch()
    dev.synchronize()
    for a, c in zip(eager[:6], graphed[:6]):
        np.testing.assert_array_equal(a.numpy(), c.numpy())
    rw, rm, rv = w.copy(), np.zeros_like(w), np.zeros_like(w)
    rww, racc, rlin = ww.copy(), np.ones_like(ww), np.zeros_like(ww)
    ast, fst = R.AdamState(3.5e-4, eps=1e-8, loss_scale=4.0), R.FtrlState(5e-2, l1=1e-8, l2=1e-8, loss_scale=4.0)
    for _ in range(2):
        ast.begin_step()
        R.lazy_adam_sparse(rw, rm, rv, uniq, R.segment_sum(g, inverse, uniq.size, wts.reshape(-1)), ast)
        R.ftrl_sparse(rww, racc, rlin, uniq, R.segment_sum(delta, inverse, uniq.size, wts.reshape(-1), div=f), fst)
    for got, ref in zip(eager[:6], (rw, rm, rv, rww, racc, rlin)):
        np.testing.assert_allclose(got.numpy(), ref, rtol=1e-5, atol=1e-6 * np.abs(ref).max())

    # FM + cross stack through the same buffers
    vx = (rng.standard_normal((64, 39, 16)) * 0.1).astype(np.float32)
    ref_fm = R.fm_forward(vx)
    np.testing.assert_allclose(ops.fm_fwd(dev.from_numpy(vx)).numpy().reshape(-1), np.asarray(ref_fm).reshape(-1),
                               rtol=1e-5, atol=1e-5 * float(0.5 * (vx.astype(np.float64) ** 2).sum(axis=(1, 2)).max()))

    # pinned staging buffer -> device on a second stream, ordered by an event
    pin = dev.pinned((b, f), "int32")
    pin[...] = ids[::-1]
    copy_stream = runtime.Stream()
    staged = dev.empty((b, f), "int32")
    dev.async_host_copies = True
    with dev.use_stream(copy_stream):
        staged.copy_(pin)
    ev = runtime.Event()
    ev.record(copy_stream)
    dev.stream.wait_event(ev)
    np.testing.assert_array_equal(ops.gather(d_w, staged).numpy(), R.gather(w, ids[::-1]))

    # mrec_rt_gemm (cuBLASLt bound with dlopen): every transpose / storage / epilogue combination vs float64 numpy
    rng2 = np.random.default_rng(5)
    for (m_, n_, k_) in ((64, 48, 40), (257, 24, 136), (8, 512, 64)):
        for ta in (False, True):
            for tb in (False, True):
                for ab, cd in (("float32", "float32"), ("float16", "float16"), ("float16", "float32")):
                    a_h = rng2.standard_normal((k_, m_) if ta else (m_, k_)).astype(ab)
                    b_h = rng2.standard_normal((n_, k_) if tb else (k_, n_)).astype(ab)
                    want = (a_h.T if ta else a_h).astype(np.float64) @ (b_h.T if tb else b_h).astype(np.float64)
                    tol = dict(rtol=2e-3, atol=2e-2) if "float16" in (ab, cd) else dict(rtol=1e-5, atol=1e-4)
                    out = dev.empty((m_, n_), cd)
                    runtime.gemm(dev.from_numpy(a_h), dev.from_numpy(b_h), out, trans_a=ta, trans_b=tb)
                    np.testing.assert_allclose(out.numpy().astype(np.float64), want, **tol)
                    bias_h = rng2.standard_normal(n_).astype(cd)
                    runtime.gemm(dev.from_numpy(a_h), dev.from_numpy(b_h), out, bias=dev.from_numpy(bias_h), relu=True,
                                 trans_a=ta, trans_b=tb)
                    np.testing.assert_allclose(out.numpy().astype(np.float64), np.maximum(want + bias_h.astype(np.float64), 0), **tol)
    acc = dev.from_numpy(np.ones((64, 48), np.float32))          # beta = 1 accumulates into out
    a_h, b_h = rng2.standard_normal((64, 40)).astype(np.float32), rng2.standard_normal((40, 48)).astype(np.float32)
    runtime.gemm(dev.from_numpy(a_h), dev.from_numpy(b_h), acc, alpha=0.5, beta=1.0)
    np.testing.assert_allclose(acc.numpy(), 1.0 + 0.5 * (a_h.astype(np.float64) @ b_h), rtol=1e-5, atol=1e-4)
    for bad in (lambda: runtime.gemm(dev.empty((4, 5)), dev.empty((6, 7)), dev.empty((4, 7))),
                lambda: runtime.gemm(dev.empty((4, 5)), dev.empty((5, 7), "float16"), dev.empty((4, 7))),
                lambda: runtime.gemm(dev.empty((4, 5)), dev.empty((5, 7)), dev.empty((4, 7)), relu=True)):
        try:
            bad()
            raise SystemExit("gemm accepted inconsistent arguments")
        except (ValueError, TypeError):
            pass
    src = rng2.standard_normal(1003).astype(np.float32)
    np.testing.assert_array_equal(ops.cast_f32_f16(dev.from_numpy(src)).numpy(), src.astype(np.float16))

    assert _lib.launch_count() - before > 20
    assert "torch" not in sys.modules, "the runtime path imported torch"
    print("RUNTIME_OK", _lib.launch_count() - before)
''')


def test_embedding_path_runs_without_torch(cuda):
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "RUNTIME_OK" in r.stdout


def test_package_import_does_not_pull_torch():
    code = "import sys; import mindrec_b200, mindrec_b200.ops, mindrec_b200.runtime, mindrec_b200.data; assert 'torch' not in sys.modules"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
