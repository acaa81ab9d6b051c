"""The graphed device-driven Wide&Deep step (peer_sharded.PeerShardedWideDeepStep: ONE CUDA graph per step — exchange
kernels, DenseLayers, the all-reduce over peer memory, dense Adam, the next batch's key phase on a forked branch behind
an external event) on a ONE-rank process group against the unsharded cell.  With one rank every wait finds its flag
set, so this covers the capture / replay / double-buffering logic and every kernel of the step on a single GPU; the
G > 1 runs are bench.py's `parity_check`.  Runs in a subprocess: it initialises torch.distributed."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    from mindrec_b200 import _lib, cells, peer_sharded, synth

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[1], RANK="0", WORLD_SIZE="1")
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    mixed = sys.argv[2] == "fp16"
    vocab, dim, b, fields, hidden = 50021, 16, 512, 39, (64, 32)
    step = peer_sharded.PeerShardedWideDeepStep(b, vocab, dim, hidden, dev, seed=3, use_mixed_precision=mixed, fields=fields)
    cfg = cells.WideDeepConfig(batch_size=b, field_size=fields, vocab_size=vocab, emb_dim=dim, deep_layer_dim=hidden,
                               use_mixed_precision=mixed, sparse=True, seed=9)
    model = cells.WideDeepModel(cfg, device=dev)
    wide0, deep0 = step.tables.gather_full()
    model.wide_embeddinglookup.embedding_table.data.copy_(wide0)
    model.deep_embeddinglookup.embedding_table.data.copy_(deep0)
    model.dense.flat.copy_(step.dense.flat)
    ref = cells.TrainStepWrap(cells.NetWithLossClass(model, cfg), sparse=True, lazy_adam=True)
    scale = vocab / synth.vocab_size(synth.CARD_KAGGLE)
    cards = [max(3, int(c * scale * 0.9)) for c in synth.CARD_KAGGLE]      # every id in range
    assert synth.vocab_size(cards) <= vocab
    gen = synth.CriteoSynth(b, cards=cards, vocab_pad=vocab, seed=4)
    raw = [gen.next() for _ in range(7)]
    for j in (2, 3, 5):                                      # ids outside [0, V): the zero row forward, no update
        raw[j][0][::37, 3] = vocab + 11
        raw[j][0][5::41, 30] = -7
    host = [tuple(torch.from_numpy(x).pin_memory() for x in hb) for hb in raw]
    batches = [tuple(x.to(dev) for x in hb) for hb in host]
    n0 = _lib.launch_count()
    step.capture(*batches[0], warmup=2)                      # trains 2 steps on batch 0
    assert step.launches_per_step > 20
    for _ in range(2):
        ref(*batches[0])
    losses = []
    # look-ahead from device tensors, from pinned host tensors, a step without look-ahead, then look-ahead again
    plan = [(batches[1], batches[2]), (batches[2], host[3]), (host[3], None), (batches[4], batches[5]), (batches[5], host[6]),
            (host[6], None)]
    for cur, nxt in plan:
        loss = step.replay(*cur, next_batch=nxt)[0]
        want = ref(*[x.to(dev) for x in cur])[0]
        losses.append((float(loss), float(want)))
    torch.cuda.synchronize()
    assert step.tables.error_flags() == 0
    tol = dict(rtol=1e-3, atol=1e-6) if mixed else dict(rtol=1e-5, atol=1e-8)
    print("losses", losses)
    wide, deep = step.tables.gather_full()
    for name, a, b_ in (("deep", deep, model.embedding_table.data), ("wide", wide, model.wide_embeddinglookup.embedding_table.data),
                        ("dense", step.dense.flat, model.dense.flat)):
        d = (a - b_).abs()
        print(name, "max", float(d.max()), "n>1e-6", int((d > 1e-6).sum()), "of", d.numel())
    for got, want in losses:
        np.testing.assert_allclose(got, want, rtol=1e-3 if mixed else 1e-6)
    torch.testing.assert_close(deep, model.embedding_table.data, **tol)
    torch.testing.assert_close(wide, model.wide_embeddinglookup.embedding_table.data, **tol)
    torch.testing.assert_close(step.dense.flat, model.dense.flat, **tol)
    # sharded checkpoint: this rank's slice -> merged state -> a fresh unsharded cell continues identically
    from mindrec_b200 import train
    merged = train.merge_sliced_tables([train.export_sharded_tables(step)])
    model2 = cells.WideDeepModel(cfg, device=dev)
    ref2 = cells.TrainStepWrap(cells.NetWithLossClass(model2, cfg), sparse=True, lazy_adam=True)
    train.import_tables(ref2, merged)
    l_sh = float(step.replay(*batches[1])[0])
    l_un = float(ref2(*batches[1])[0])
    np.testing.assert_allclose(l_sh, l_un, rtol=1e-3 if mixed else 1e-6)
    wide, deep = step.tables.gather_full()
    torch.testing.assert_close(deep, model2.embedding_table.data, **tol)
    torch.testing.assert_close(step.dense.flat, model2.dense.flat, **tol)
    step.close()
    dist.destroy_process_group()
    print("PEER_STEP_OK", step.launches_per_step)
''')


@pytest.mark.parametrize("mlp", ["fp32", "fp16"])
def test_graphed_peer_step_on_one_rank_matches_the_unsharded_cell(cuda, mlp):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-c", SCRIPT, str(port), mlp], capture_output=True, text=True, timeout=600,
                       cwd=ROOT, env=env)
    assert r.returncode == 0 and "PEER_STEP_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
