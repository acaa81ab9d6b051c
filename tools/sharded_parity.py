"""Multi-GPU parity of the row-sharded steps, run by bench.py at N > 1 (and usable on its own under torchrun):

    wide_deep(...)    G ranks of peer_sharded.PeerShardedWideDeepStep (device-driven exchange, CUDA graphs, look-ahead
                      plan — the benchmarked path) on rank-specific batches vs ONE unsharded cells.TrainStepWrap on
                      rank 0 trained on the concatenated global batch from the same initial state
                      (gradients_mean semantics, models/wide_deep/src/wide_and_deep.py:455-470).
    multitable(...)   G ranks of multitable_sharded.ShardedMultitableStep vs the same class on a one-rank group on
                      rank 0 (owner = key mod 1: the unsharded layout) fed the concatenated batch.

Shapes are the benchmarked ones (batch, fields, dims, DenseLayer sizes) with the table scaled down so that rank 0 can
also hold the unsharded copy.

What is compared: every table row, the DenseLayer weights and the per-step loss.  Adam's first update of an element is
lr * sign(g) whatever |g| is, so an element whose gradient sum is at rounding-noise level can legitimately flip between
two summation orders; the verdict is therefore `outlier fraction <= 1e-6` at 2e-5 of the tensor's scale plus a relative
L2 bound, and the raw max |diff| is reported beside it.
"""
import numpy as np
import torch
import torch.distributed as dist

TOL_SCALE = 2e-5


def _cmp(name, got, want, out, l2_tol):
    """fp32 DenseLayers (l2_tol 1e-5) are the parity gate: at most 1e-6 of the elements beyond 2e-5 of the scale, and the
    relative L2 error.  fp16 DenseLayers (l2_tol 1e-3): the two runs call GEMMs of different M (B vs G * B rows), which
    cuBLAS tiles and rounds differently, and Adam turns a last-bit fp16 difference of a gradient into up to lr per step
    — there the relative L2 error (<= fp16 epsilon) carries the verdict and the outlier count is reported only."""
    d = (got.double() - want.double())
    scale = float(want.abs().max())
    n_out = int((d.abs() > TOL_SCALE * scale).sum())
    rel_l2 = float(d.norm() / max(float(want.double().norm()), 1e-30))
    out[name] = {"max_abs_diff": float(d.abs().max()), "scale": scale, "outliers": n_out, "numel": got.numel(),
                 "rel_l2": rel_l2}
    if l2_tol > 1e-5:
        return rel_l2 <= l2_tol
    return n_out <= 1e-6 * got.numel() + 3 and rel_l2 <= l2_tol


def wide_deep(world, rank, dev, batch, fields, emb, hidden, rows_per_rank=2_000_003, steps=3, mixed=False, alpha=1.05,
              graph=True, ahead=True):
    """graph / ahead = False run the same step eagerly / without the one-step-ahead key phase (bisection aids)."""
    from mindrec_b200 import cells, peer_sharded, synth
    vocab = rows_per_rank * world
    step = peer_sharded.PeerShardedWideDeepStep(batch, vocab, emb, hidden, dev, seed=3, use_mixed_precision=mixed,
                                                fields=fields, graph=graph)
    wide0, deep0 = step.tables.gather_full()
    flat0 = step.dense.flat.clone()
    scale = vocab / synth.vocab_size(synth.CARD_KAGGLE)
    cards = [max(3, int(c * scale * 0.98)) for c in synth.CARD_KAGGLE]
    gens = [synth.CriteoSynth(batch, cards=cards, alpha=alpha, vocab_pad=vocab, seed=777, rank=r) for r in range(world)]
    host = [[g.next() for _ in range(steps + 1)] for g in gens]          # batch 0 is the capture's warm-up batch
    mine = [tuple(torch.from_numpy(x).to(dev) for x in b) for b in host[rank]]
    step.capture(*mine[0], warmup=2)                                     # trains 2 steps on batch 0
    losses = []
    for s in range(1, steps + 1):
        nxt = mine[s + 1] if (ahead and s + 1 <= steps) else None
        losses.append(step.replay(*mine[s], next_batch=nxt)[0].reshape(1).clone())
    torch.cuda.synchronize()
    flags = step.tables.error_flags()
    wide, deep = step.tables.gather_full()
    lt = torch.cat(losses)
    all_l = [torch.empty_like(lt) for _ in range(world)]
    dist.all_gather(all_l, lt)
    res = {"steps": steps + 2, "mlp": "fp16" if mixed else "fp32", "rows": vocab, "batch_per_gpu": batch, "dim": emb,
           "exchange_flags": flags}
    ok = flags == 0
    if rank == 0:
        cfg = cells.WideDeepConfig(batch_size=batch * world, field_size=fields, vocab_size=vocab, emb_dim=emb,
                                   deep_layer_dim=hidden, use_mixed_precision=mixed, sparse=True, seed=9)
        model = cells.WideDeepModel(cfg, device=dev)
        model.wide_embeddinglookup.embedding_table.data.copy_(wide0)
        model.deep_embeddinglookup.embedding_table.data.copy_(deep0)
        model.dense.flat.copy_(flat0)
        del wide0, deep0
        ref = cells.TrainStepWrap(cells.NetWithLossClass(model, cfg), sparse=True, lazy_adam=True)
        ref_losses = []
        for s in [0, 0] + list(range(1, steps + 1)):
            cat = [torch.from_numpy(np.concatenate([host[r][s][i] for r in range(world)])).to(dev) for i in range(3)]
            ref_losses.append(float(ref(*cat)[0]))
        sh = torch.stack(all_l).mean(0).tolist()
        l_tol = 2e-3 if mixed else 1e-5
        res["loss_rel_diff"] = max(abs(a - b) / abs(b) for a, b in zip(sh, ref_losses[2:]))
        ok &= res["loss_rel_diff"] <= l_tol and all(np.isfinite(sh))
        l2 = 1e-3 if mixed else 1e-5
        ok &= _cmp("deep", deep, model.embedding_table.data, res, l2)
        ok &= _cmp("wide", wide, model.wide_embeddinglookup.embedding_table.data, res, l2)
        ok &= _cmp("dense", step.dense.flat, model.dense.flat, res, l2)
        del ref, model
    del wide, deep
    step.close()
    del step
    torch.cuda.empty_cache()
    okt = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(okt, 0)
    res["ok"] = bool(okt.item())
    return res


def _hash_rows(table, dev, world, group=None):
    """(keys, rows) of every rank's MapParameter, concatenated on every rank and sorted by key."""
    k, v = table.get_data()
    n = torch.tensor([k.numel()], device=dev)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    n_max = int(max(int(x.item()) for x in ns))
    kp = torch.full((n_max,), -1, dtype=torch.int64, device=dev)
    vp = torch.zeros((n_max, v.shape[1]), dtype=torch.float32, device=dev)
    kp[:k.numel()] = k
    vp[:k.numel()] = v
    ks = [torch.empty_like(kp) for _ in range(world)]
    vs = [torch.empty_like(vp) for _ in range(world)]
    dist.all_gather(ks, kp, group=group)
    dist.all_gather(vs, vp, group=group)
    k, v = torch.cat(ks), torch.cat(vs)
    keep = k >= 0
    k, v = k[keep], v[keep]
    order = torch.argsort(k)
    return k[order], v[order]


def c5_batches(batch, n_table_fields, n_hash_fields, rows_total, key_bits, seed, rank, count, alpha=1.05):
    """Config-5 inputs: Zipf ids over the sharded table (one slice of rows per field), int64 Zipf keys over
    2^key_bits for the MapParameter, Bernoulli(0.25) labels."""
    rng = np.random.default_rng(seed + rank)
    per = rows_total // n_table_fields
    out = []
    for _ in range(count):
        z = rng.zipf(alpha, size=(batch, n_table_fields)) - 1
        ids = (np.arange(n_table_fields, dtype=np.int64)[None, :] * per + z % per).astype(np.int32)
        zk = rng.zipf(alpha, size=(batch, n_hash_fields)) - 1
        keys = ((zk * n_hash_fields + np.arange(n_hash_fields, dtype=np.int64)[None, :]) % (1 << key_bits)).astype(np.int64)
        label = (rng.random((batch, 1)) < 0.25).astype(np.float32)
        out.append((ids, keys, label))
    return out


def multitable(world, rank, dev, batch, rows_per_rank=1_000_003, steps=3, mixed=False, group_one=None, **kw):
    from mindrec_b200 import multitable_sharded as M
    rows = rows_per_rank * world
    step = M.ShardedMultitableStep(batch, rows, dev, use_mixed_precision=mixed, seed=5, **kw)
    wide0, deep0 = step.tables.gather_full()
    flat0, bias0 = step.dense.flat.clone(), step.wide_bias.clone()
    host = [c5_batches(batch, step.ft, step.fh, rows, step.hash.rk.bits, 4242, r, steps + 1) for r in range(world)]
    mine = [tuple(torch.from_numpy(x).to(dev) for x in b) for b in host[rank]]
    step.capture(*mine[0], warmup=2)
    losses = []
    for s in range(1, steps + 1):
        losses.append(step.replay(*mine[s])[0].reshape(1).clone())
    torch.cuda.synchronize()
    flags = step.error_flags()
    wide, deep = step.tables.gather_full()
    hk, hv = _hash_rows(step.hash.rk.table, dev, world)
    lt = torch.cat(losses)
    all_l = [torch.empty_like(lt) for _ in range(world)]
    dist.all_gather(all_l, lt)
    res = {"steps": steps + 2, "mlp": "fp16" if mixed else "fp32", "rows": rows, "batch_per_gpu": batch,
           "exchange_flags": flags, "hash_resident_keys": int(hk.numel())}
    ok = flags == 0
    if rank == 0:
        # the one-rank reference holds the keys of ALL ranks in one MapParameter: G times the slots (a table that fills
        # up makes every find-or-insert probe its whole length)
        kw_ref = dict(kw, hash_capacity=kw.get("hash_capacity", 1 << 22) * world)
        ref = M.ShardedMultitableStep(batch * world, rows, dev, group=group_one, use_mixed_precision=mixed, seed=5, **kw_ref)
        ref.tables.rk.wide.copy_(wide0[:ref.tables.rk.wide.shape[0]])
        ref.tables.rk.deep.copy_(deep0[:ref.tables.rk.deep.shape[0]])
        ref.dense.flat.copy_(flat0)
        ref.wide_bias.copy_(bias0)
        del wide0, deep0
        ref_losses = []
        for s in [0, 0] + list(range(1, steps + 1)):
            cat = [torch.from_numpy(np.concatenate([host[r][s][i] for r in range(world)])).to(dev) for i in range(3)]
            ref_losses.append(float(ref(*cat)[0]))
        torch.cuda.synchronize()
        sh = torch.stack(all_l).mean(0).tolist()
        res["loss_rel_diff"] = max(abs(a - b) / abs(b) for a, b in zip(sh, ref_losses[2:]))
        ok &= res["loss_rel_diff"] <= (2e-3 if mixed else 1e-5) and all(np.isfinite(sh))
        l2 = 1e-3 if mixed else 1e-5
        v = rows
        ok &= _cmp("emb128", deep, ref.tables.rk.deep[:v], res, l2)
        ok &= _cmp("wide_emb128", wide, ref.tables.rk.wide[:v], res, l2)
        ok &= _cmp("dense", step.dense.flat, ref.dense.flat, res, l2)
        ok &= _cmp("wide_bias", step.wide_bias, ref.wide_bias, res, l2)
        rk, rv = ref.hash.rk.table.get_data()
        same_keys = rk.numel() == hk.numel() and bool(torch.equal(rk, hk))
        res["hash_keys_equal"] = same_keys
        ok &= same_keys and ref.error_flags() == 0
        if same_keys:
            ok &= _cmp("hash_rows", hv, rv, res, l2)
        ref.close()
        del ref
    del wide, deep
    step.close()
    del step
    torch.cuda.empty_cache()
    okt = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(okt, 0)
    res["ok"] = bool(okt.item())
    return res
