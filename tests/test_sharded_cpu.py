"""Host logic of the row-sharded path, world_size 2 over gloo on CPU (no GPU needed).

The kernels are replaced by the numpy oracle (tests/oracle_kernels.py); what is under test is the exchange:
owner-major remap, bucket bounds, split sizes, the key / row / gradient all-to-alls, owner-side dedup and the
mean all-reduce of the DenseLayer gradients.  The sharded result must equal ONE process training on the
concatenated global batch (gradients_mean semantics)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mindrec_b200 import nn, sharded, synth
from oracle import ref_numpy as R
from tests import fake_cuda, oracle_kernels as OK

VOCAB, DIM, B, HIDDEN = 701, 8, 24, (16, 8)


def _patch():
    """The product modules have ONE path (CUDA).  The exchange logic is run on the CPU by replacing, from the test
    side, the kernel layer with the oracle and the stream objects with synchronous stand-ins."""
    sharded.ops = OK
    sharded._cu = fake_cuda
    sharded._pinned = lambda t: t
    nn.ops = OK


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batches(world, steps):
    out = []
    for r in range(world):
        gen = synth.CriteoSynth(B, cards=[20] * 26, vocab_pad=VOCAB, seed=100, rank=r)
        out.append([gen.next() for _ in range(steps)])
    return out


def _worker(rank, world, port, steps, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        _patch()
        step = sharded.ShardedWideDeepStep(B, VOCAB, DIM, HIDDEN, "cpu", seed=3, use_mixed_precision=False,
                                           graph_dense=False)
        wide0, deep0 = step.tables.gather_full()
        init = dict(wide=wide0.numpy().copy(), deep=deep0.numpy().copy(),
                    w=[w.numpy().copy() for w in step.dense.weights],
                    b=[b.numpy().copy() for b in step.dense.biases], wide_b=step.wide_b.numpy().copy())
        losses = []
        for ids, wts, label in _batches(world, steps)[rank]:
            loss, _ = step(torch.from_numpy(ids), torch.from_numpy(wts), torch.from_numpy(label))
            losses.append(float(loss))
        wide, deep = step.tables.gather_full()
        if rank == 0:
            ret.update(init=init, wide=wide.numpy(), deep=deep.numpy(), losses=losses,
                       w0=step.dense.weights[0].numpy().copy())
        ret["loss%d" % rank] = losses
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_training_equals_single_process_on_the_global_batch():
    world, steps = 2, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, steps, ret), nprocs=world, join=True)
    init = ret["init"]
    orc = R.WideDeepOracle(init["wide"], init["deep"], init["w"], init["b"], init["wide_b"], mode="lazy")
    per_rank = _batches(world, steps)
    for s in range(steps):
        ids = np.concatenate([per_rank[r][s][0] for r in range(world)])
        wts = np.concatenate([per_rank[r][s][1] for r in range(world)])
        label = np.concatenate([per_rank[r][s][2] for r in range(world)]).astype(np.float64)
        ref_loss, _ = orc.step(ids, wts, label)
        mean_loss = np.mean([ret["loss%d" % r][s] for r in range(world)])
        np.testing.assert_allclose(mean_loss, ref_loss, rtol=1e-5)
    np.testing.assert_allclose(ret["deep"], orc.wd, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(ret["wide"], orc.ww, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(ret["w0"], orc.mlp_w[0], rtol=1e-4, atol=1e-7)


def test_shard_plan_remap_round_trip():
    plan = sharded.ShardPlan(1000, 8)
    ids = torch.arange(-3, 1005)
    km = plan.remap(ids)
    ok = (ids >= 0) & (ids < 1000)
    assert (km[~ok] == 8 * plan.rows_per_rank).all()
    owner, local = km[ok] // plan.rows_per_rank, km[ok] % plan.rows_per_rank
    assert torch.equal(owner, ids[ok] % 8) and torch.equal(local, ids[ok] // 8)
    assert torch.equal(local * 8 + owner, ids[ok])
    assert km[ok].unique().numel() == 1000
    with pytest.raises(ValueError):
        sharded.ShardPlan(2 ** 31, 2)


def test_single_rank_sharded_step_equals_oracle(monkeypatch):
    """world_size 1 path (no process group): same exchange code, degenerate splits."""
    monkeypatch.setattr(sharded, "ops", OK)
    monkeypatch.setattr(sharded, "_cu", fake_cuda)
    monkeypatch.setattr(sharded, "_pinned", lambda t: t)
    monkeypatch.setattr(nn, "ops", OK)
    step = sharded.ShardedWideDeepStep(B, VOCAB, DIM, HIDDEN, "cpu", seed=5, use_mixed_precision=False,
                                       graph_dense=False)
    wide0, deep0 = step.tables.gather_full()
    orc = R.WideDeepOracle(wide0.numpy(), deep0.numpy(), [w.numpy() for w in step.dense.weights],
                           [b.numpy() for b in step.dense.biases], step.wide_b.numpy(), mode="lazy")
    gen = synth.CriteoSynth(B, cards=[20] * 26, vocab_pad=VOCAB, seed=7)
    for _ in range(2):
        ids, wts, label = gen.next()
        loss, _ = step(torch.from_numpy(ids), torch.from_numpy(wts), torch.from_numpy(label))
        ref, _ = orc.step(ids, wts, label.astype(np.float64))
        np.testing.assert_allclose(float(loss), ref, rtol=1e-5)
    wide, deep = step.tables.gather_full()
    np.testing.assert_allclose(deep.numpy(), orc.wd, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(wide.numpy(), orc.ww, rtol=1e-4, atol=1e-7)
