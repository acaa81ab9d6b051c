"""BASELINE config 5 step (multitable_sharded.ShardedMultitableStep) on a ONE-rank process group: the whole protocol
runs (plan, publish, key / row / gradient exchange through the IPC inboxes, owner dedup, fused row updates, MapParameter
admission + eviction), only the peers are missing.  The lookups and the row updates are checked against the numpy oracle
fed the step's own DenseLayer gradient; graph replay is checked against eager execution.  The G > 1 form is checked by
bench.py's parity_check (G ranks vs this one-rank form) on the multi-GPU box."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist

from mindrec_b200 import multitable_sharded as M
from oracle import ref_numpy as R
from tools import sharded_parity

pytestmark = pytest.mark.gpu
D = 128


@pytest.fixture(scope="module")
def one_rank_group(cuda):
    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29741")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=cuda)
        created = True
    yield None
    if created:
        dist.destroy_process_group()


def _build(cuda, b, rows, ft, fh, **kw):
    return M.ShardedMultitableStep(b, rows, cuda, n_table_fields=ft, n_hash_fields=fh, deep_dim_list=(64, 32),
                                   hash_capacity=1 << 14, use_mixed_precision=False, seed=3, **kw)


def test_step_matches_oracle_on_one_rank(cuda, one_rank_group):
    b, rows, ft, fh = 96, 5003, 5, 4
    step = _build(cuda, b, rows, ft, fh, permit_filter_value=2, evict_filter_value=2, evict_every=0)
    deep = step.tables.rk.deep.cpu().numpy().copy()
    wide = step.tables.rk.wide.cpu().numpy().copy()
    m, v = np.zeros_like(deep), np.zeros_like(deep)
    acc, lin = np.full_like(wide, 0.1), np.zeros_like(wide)
    adam = R.AdamState(3e-3, eps=1e-6, loss_scale=1000.0)
    ftrl = R.FtrlState(0.1, l1=5e-4, l2=5e-4, loss_scale=1000.0)
    hmodel = R.MapParameterModel(D, default_value=0.0, permit_filter_value=2)
    hm, hv = {}, {}
    hadam = R.AdamState(3e-3, eps=1e-6, loss_scale=1000.0)
    batches = sharded_parity.c5_batches(b, ft, fh, rows, 40, 11, 0, 4)
    for it, (ids, keys, label) in enumerate(batches):
        keys = keys % 300 * 7919                                     # few distinct keys: admissions happen
        bias_before = float(step.wide_bias)
        deep_before, wide_before = step.tables.rk.deep.cpu().numpy(), step.tables.rk.wide.cpu().numpy()
        loss, _ = step(torch.from_numpy(ids).to(cuda), torch.from_numpy(keys).to(cuda), torch.from_numpy(label).to(cuda))
        assert np.isfinite(float(loss)) and step.error_flags() == 0
        io = step._io
        # ---- forward: the expanded rows are the shard's rows as they stood before the step (bit-exact) ----
        np.testing.assert_array_equal(io["x_table"].cpu().numpy(), deep_before[ids].reshape(b, ft * D))
        want_wide = wide_before[ids, 0].astype(np.float64).sum(1, keepdims=True) + bias_before
        np.testing.assert_allclose(io["wide_out"].cpu().numpy(), want_wide, rtol=1e-5, atol=1e-7)
        hrows = hmodel.get(keys.reshape(-1))                          # admission counted here, once per step
        got_h = io["x_hash"].cpu().numpy().reshape(-1, D)
        resident = np.asarray([k in hmodel.rows for k in keys.reshape(-1).tolist()])
        new = np.asarray([k in hmodel.rows and k not in hm for k in keys.reshape(-1).tolist()])
        assert np.all(got_h[~resident] == 0.0)                         # not admitted yet: the default row
        for k, row in zip(keys.reshape(-1)[new].tolist(), got_h[new]):  # Philox init keyed by the key: adopt it
            hmodel.rows[k] = row.copy()
            hm[k], hv[k] = np.zeros(D, np.float32), np.zeros(D, np.float32)
        np.testing.assert_allclose(got_h, np.stack([hmodel.rows.get(k, hmodel.default) for k in keys.reshape(-1).tolist()]),
                                   rtol=1e-5, atol=1e-7)
        _, first, kinv0 = np.unique(keys.reshape(-1), return_index=True, return_inverse=True)
        np.testing.assert_array_equal(got_h, got_h[first[kinv0]])       # every copy of a key reads the same row
        # ---- backward: LazyAdam / FTRL rows from the step's own DenseLayer gradient ----
        g_t = step._last["g_table"].cpu().numpy().reshape(-1, D)
        delta = step._last["delta"].cpu().numpy()
        uniq, inverse, _, _ = R.unique_sorted(ids, bound=rows)
        adam.begin_step()
        R.lazy_adam_sparse(deep, m, v, uniq, R.segment_sum(g_t, inverse, uniq.size), adam)
        R.ftrl_sparse(wide, acc, lin, uniq, R.segment_sum(delta, inverse, uniq.size, div=ft), ftrl)
        g_h = step._last["g_hash"].cpu().numpy().reshape(-1, D)
        hadam.begin_step()
        ku, kinv = np.unique(keys.reshape(-1), return_inverse=True)
        gs = R.segment_sum(g_h, kinv, ku.size)
        for k, g in zip(ku.tolist(), gs):
            if k in hm:                                               # rows of unadmitted keys receive no update
                w_, m_, v_ = R.adam_rows(hmodel.rows[k][None], hm[k][None], hv[k][None], g[None], hadam)
                hmodel.rows[k], hm[k], hv[k] = w_[0], m_[0], v_[0]
    tol = dict(rtol=1e-5)
    np.testing.assert_allclose(step.tables.rk.deep.cpu().numpy(), deep, atol=1e-6 * np.abs(deep).max(), **tol)
    np.testing.assert_allclose(step.tables.rk.wide.cpu().numpy(), wide, atol=1e-6 * np.abs(wide).max(), **tol)
    k, vals = step.hash.rk.table.get_data()
    assert k.cpu().tolist() == sorted(hmodel.rows)
    ref_rows = np.stack([hmodel.rows[x] for x in k.cpu().tolist()])
    np.testing.assert_allclose(vals.cpu().numpy(), ref_rows, atol=1e-6 * np.abs(ref_rows).max(), **tol)
    step.close()


def test_graph_replay_equals_eager_and_evicts(cuda, one_rank_group):
    b, rows, ft, fh = 64, 3001, 4, 3
    a = _build(cuda, b, rows, ft, fh, permit_filter_value=1, evict_filter_value=1, evict_every=2)
    e = _build(cuda, b, rows, ft, fh, permit_filter_value=1, evict_filter_value=1, evict_every=2)
    batches = [tuple(torch.from_numpy(x).to(cuda) for x in hb) for hb in sharded_parity.c5_batches(b, ft, fh, rows, 40, 5, 0, 5)]
    a.capture(*batches[0], warmup=2)
    assert a.launches_per_step and a.launches_per_step > 20
    for _ in range(2):
        e(*batches[0])
    for bt in batches[1:]:
        la = float(a.replay(*bt)[0])
        le = float(e(*bt)[0])
        assert la == le
    torch.cuda.synchronize()
    assert torch.equal(a.tables.rk.deep, e.tables.rk.deep) and torch.equal(a.tables.rk.wide, e.tables.rk.wide)
    assert torch.equal(a.dense.flat, e.dense.flat) and torch.equal(a.wide_bias, e.wide_bias)
    (ka, va), (ke, ve) = a.hash.rk.table.get_data(), e.hash.rk.table.get_data()
    assert torch.equal(ka, ke) and torch.equal(va, ve)
    # eviction ran: keys of the first batches that never came back are gone
    # 6 steps, sweep every 2nd step, keys unseen for more than 1 step go: what is left is what steps 5 and 6 looked up
    seen_late = set(torch.cat([bt[1].reshape(-1) for bt in batches[3:]]).tolist())
    seen_all = set(torch.cat([bt[1].reshape(-1) for bt in batches]).tolist())
    assert set(ka.tolist()) == seen_late and len(seen_late) < len(seen_all)
    assert a.error_flags() == 0 and e.error_flags() == 0
    a.close()
    e.close()
