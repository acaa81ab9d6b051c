"""Multi-GPU parity check of the row-sharded path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded_multigpu.py

G ranks train 3 steps on rank-specific batches with sharded tables (NCCL key all-to-all, fused gather + NVLink
peer stores for the rows unless MREC_SHARDED_PEER=0, NCCL gradient all-to-all, mean all-reduce of the
DenseLayer gradients); rank 0 then trains ONE unsharded cell (cells.TrainStepWrap) on the concatenated global
batch from the same initial state and the two must agree (gradients_mean semantics).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindrec_b200 import cells, peer_sharded, sharded, synth  # noqa: E402

# MREC_CHECK_EXCHANGE = nccl (default) | device (eager device-driven exchange) | device-graph (whole step as one graph)
MODE = os.environ.get("MREC_CHECK_EXCHANGE", "nccl")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    vocab, dim, b, hidden, steps = 50021, 16, 512, (64, 32), 5
    if MODE == "nccl":
        step = sharded.ShardedWideDeepStep(b, vocab, dim, hidden, dev, seed=3, use_mixed_precision=False,
                                           graph_dense=False)
    else:
        step = peer_sharded.PeerShardedWideDeepStep(b, vocab, dim, hidden, dev, seed=3, use_mixed_precision=False,
                                                    graph=(MODE == "device-graph"))
    wide0, deep0 = step.tables.gather_full()
    flat0 = step.dense.flat.clone()
    gens = [synth.CriteoSynth(b, cards=[1500] * 26, vocab_pad=vocab, seed=100, rank=r) for r in range(world)]
    batches = [[g.next() for _ in range(steps)] for g in gens]
    if MODE == "device-graph":                      # capture trains 2 warm-up steps on batch 0, then replays
        batches = [[bs[0], bs[0]] + bs[1:] for bs in batches]
        steps += 1
    losses = []
    dev_batches = [tuple(torch.from_numpy(x).to(dev) for x in batches[rank][s]) for s in range(steps)]
    for s in range(steps):
        ids, wts, label = dev_batches[s]
        if MODE == "device-graph":
            if s == 0:
                step.capture(ids, wts, label, warmup=2)
                continue
            if s == 1:
                losses += [float("nan")] * 2
                continue
            # the next batch is handed over too: staged and planned (dedup + bounds) underneath this step
            nxt = dev_batches[s + 1] if s + 1 < steps else None
            losses.append(float(step.replay(ids, wts, label, next_batch=nxt)[0]))
        else:
            losses.append(float(step(ids, wts, label)[0]))
    if MODE != "nccl":
        flags = step.tables.error_flags()
        if flags:
            print("rank %d: exchange error flags %d" % (rank, flags), flush=True)
    wide, deep = step.tables.gather_full()
    all_losses = [None] * world
    dist.all_gather_object(all_losses, losses)
    ok = True
    if rank == 0:
        cfg = cells.WideDeepConfig(batch_size=b * world, vocab_size=vocab, emb_dim=dim, deep_layer_dim=hidden,
                                   use_mixed_precision=False, sparse=True, seed=9)
        model = cells.WideDeepModel(cfg, device=dev)
        model.wide_embeddinglookup.embedding_table.data.copy_(wide0)
        model.deep_embeddinglookup.embedding_table.data.copy_(deep0)
        model.dense.flat.copy_(flat0)
        ref = cells.TrainStepWrap(cells.NetWithLossClass(model, cfg), sparse=True, lazy_adam=True)
        for s in range(steps):
            cat = [np.concatenate([batches[r][s][i] for r in range(world)]) for i in range(3)]
            l_ref = float(ref(*(torch.from_numpy(x).to(dev) for x in cat))[0])
            l_sh = float(np.mean([all_losses[r][s] for r in range(world)]))
            ok &= bool(np.isnan(l_sh)) or abs(l_ref - l_sh) <= 1e-5 * abs(l_ref)
        for name, got, want in (("deep", deep, model.embedding_table.data),
                                ("wide", wide, model.wide_embeddinglookup.embedding_table.data),
                                ("dense", step.dense.flat, model.dense.flat)):
            err = float((got - want).abs().max())
            scale = float(want.abs().max())
            print("%-6s max|diff| %.3e  (scale %.3e)" % (name, err, scale))
            ok &= err <= 2e-5 * scale
        print("mode:", MODE, "| peer path:", getattr(step.tables, "peer", True) is not None, "| SHARDED PARITY", "OK" if ok else "FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
