"""The torch-free Wide&Deep training step (mindrec_b200.rt_wide_deep.WideDeepRT: runtime buffers, aot kernels, cuBLASLt
through mrec_rt_gemm) against the numpy oracle, in a process where importing torch raises.  fp32 DenseLayers at the
tolerances of tests/test_wide_deep_gpu.py; CUDA-graph replay bit-identical to eager; mixed precision (fp16 DenseLayers,
fp32 accumulation) against the fp64 oracle at fp16 tolerances."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent('''
    import sys


    class NoTorch:
        def find_spec(self, name, path=None, target=None):
            if name == "torch" or name.startswith("torch."):
                raise ImportError("the torch-free path tried to import " + name)
            return None


    sys.meta_path.insert(0, NoTorch())
    import numpy as np
    from mindrec_b200 import _lib, runtime, synth
    from mindrec_b200.rt_wide_deep import WideDeepRT
    from oracle import ref_numpy as R

    dev = runtime.Device(0)
    vocab, batch, emb, fields, hidden = 3000, 257, 16, 39, (64, 32)
    rng = np.random.default_rng(11)


    def build(mixed):
        step = WideDeepRT(dev, batch, fields, vocab, emb, hidden, mixed=mixed, seed=5)
        wide = rng.normal(0, 0.01, (vocab, 1)).astype(np.float32)
        deep = rng.normal(0, 0.01, (vocab, emb)).astype(np.float32)
        step.load_state(wide=wide, deep=deep)
        flat = step.flat.numpy()
        ws, bs, o = [], [], 0
        dims = step.dims
        for i in range(len(dims) - 1):
            k, m = dims[i], dims[i + 1]
            ws.append(flat[o:o + k * m].reshape(k, m).copy()); o += k * m
            bs.append(flat[o:o + m].copy()); o += m
        orc = R.WideDeepOracle(wide, deep, ws, bs, flat[o:o + 1].copy(), sens=1024.0, mode="lazy")
        return step, orc


    def batches(seed, count):
        gen = synth.CriteoSynth(batch, cards=[50] * 26, vocab_pad=vocab, seed=seed)
        return [gen.next() for _ in range(count)]


    def unflatten(step):
        flat, out, o = step.flat.numpy(), [], 0
        for i in range(len(step.dims) - 1):
            k, m = step.dims[i], step.dims[i + 1]
            out.append(flat[o:o + k * m].reshape(k, m)); o += k * m
            out.append(flat[o:o + m]); o += m
        return out, flat[o:o + 1]

    # ---- fp32 DenseLayers, eager: vs the oracle -------------------------------------------------------------------
    step, orc = build(False)
    n0 = _lib.launch_count()
    for ids, wts, label in batches(3, 3):
        loss = step.train_step(ids, wts, label).item()
        want, _ = orc.step(ids, wts, label.astype(np.float64))
        np.testing.assert_allclose(loss, want, rtol=1e-5)
    params, wide_b = unflatten(step)
    ref_params = [x for pair in zip(orc.mlp_w, orc.mlp_b) for x in pair]
    pairs = [(step.wide.numpy(), orc.ww), (step.deep.numpy(), orc.wd), (step.acc.numpy(), orc.acc), (step.lin.numpy(), orc.lin),
             (step.m.numpy(), orc.md), (step.vv.numpy(), orc.vd), (wide_b, orc.wide_b)] + list(zip(params, ref_params))
    for got, ref in pairs:
        np.testing.assert_allclose(got, np.asarray(ref).reshape(got.shape), rtol=2e-4, atol=2e-5 * max(np.abs(ref).max(), 1e-12))
    assert _lib.launch_count() - n0 >= 3 * 15

    # ---- graph replay == eager, bit for bit (the capture's warm-up step trains: run it on the twin too) ------------
    rng = np.random.default_rng(11)
    a, _ = build(False)
    rng = np.random.default_rng(11)
    e, _ = build(False)
    bt = batches(9, 4)
    a.set_inputs(*bt[0])
    a.capture(warmup=1)
    assert a.launches_per_step >= 15
    e.train_step(*bt[0])
    dbt = [tuple(dev.from_numpy(x) for x in b_) for b_ in bt]          # device-resident copies: staged D2D on the copy stream
    for i in range(1, 4):
        nxt = (dbt[i + 1] if i % 2 else bt[i + 1]) if i + 1 < 4 else None   # announced from the device / from the host / not
        cur = dbt[i] if (i - 1) % 2 and i > 1 else bt[i]
        la = a.train_step(*cur, next_batch=nxt).item()
        le = e.train_step(*bt[i]).item()
        assert la == le
    for x, y in ((a.deep, e.deep), (a.wide, e.wide), (a.flat, e.flat), (a.m, e.m), (a.acc, e.acc)):
        np.testing.assert_array_equal(x.numpy(), y.numpy())

    # ---- mixed precision: fp16 DenseLayers with fp32 accumulation vs the fp64 oracle ---------------------------------
    rng = np.random.default_rng(11)
    step, orc = build(True)
    for ids, wts, label in batches(3, 3):
        loss = step.train_step(ids, wts, label).item()
        want, _ = orc.step(ids, wts, label.astype(np.float64))
        np.testing.assert_allclose(loss, want, rtol=2e-3)
    d = step.deep.numpy() - orc.wd
    assert np.linalg.norm(d) / np.linalg.norm(orc.wd) < 2e-2
    assert "torch" not in sys.modules, "the torch-free step imported torch"
    print("RT_WIDE_DEEP_OK", a.launches_per_step)
''')


def test_torch_free_wide_deep_step_matches_oracle(cuda):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and "RT_WIDE_DEEP_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
