"""Multi-GPU parity check of the hash-sharded MapParameter (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded_hash_multigpu.py

G ranks look up and LazyAdam-update rank-specific int64 keys through peer_sharded.PeerShardedHashEmbedding
(owner = hash(key) mod G, device-driven exchange over CUDA-IPC peer memory) for 3 steps; rank 0 replays all ranks'
keys and gradients on ONE MapParameter and the union of the G tables must equal it (1e-5 relative; the gradients
are kept positive so that per-key sums do not cancel — a cancelling fp32 sum has no 1e-5 relative meaning).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindrec_b200 import hash as H, ops, peer_sharded  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    dim, n, bits, steps = 32, 20000, 40, 3
    emb = peer_sharded.PeerShardedHashEmbedding(dim, n, dev, key_bits=bits, capacity=1 << 17, seed=11, learning_rate=1e-2)
    rng = np.random.default_rng(5)                       # same stream on every rank: everybody knows all batches
    pool = rng.integers(0, 1 << bits, size=60000)
    batches = [[(rng.choice(pool[: 20000 * (s + 1)], size=n), (np.abs(rng.standard_normal((n, dim))) + 0.5).astype(np.float32))
                for _ in range(world)] for s in range(steps)]
    outs = []
    for s in range(steps):
        keys, g = batches[s][rank]
        out = emb.lookup(torch.from_numpy(keys).to(dev))
        outs.append(out.clone())
        emb.update(torch.from_numpy(g).to(dev))
    torch.cuda.synchronize()
    flags = emb.error_flags()
    k_loc, v_loc = emb.rk.table.get_data()
    gathered = [None] * world
    dist.all_gather_object(gathered, (k_loc.cpu().numpy(), v_loc.cpu().numpy(), [o.cpu().numpy() for o in outs], flags))
    ok = True
    if rank == 0:
        one = H.MapParameter(key_dtype=torch.int64, value_shape=dim, default_value="normal", capacity=1 << 18, device=dev, seed=11)
        m1, v1 = one.add_arena(0.0), one.add_arena(0.0)
        hyper = ops.adam_hyper(1e-2, device=dev)
        c = one.capacity
        worst = 0.0
        for s in range(steps):
            all_keys = torch.from_numpy(np.concatenate([batches[s][r][0] for r in range(world)])).to(dev)
            slots = one.lookup_slots(all_keys).clone()
            want = ops.gather(one.values, slots).view(world, n, dim).cpu().numpy()
            for r in range(world):
                got = gathered[r][2][s]
                if s == 0:
                    ok &= bool(np.array_equal(got, want[r]))
                worst = max(worst, float(np.abs(got - want[r]).max()))
            g_all = torch.from_numpy(np.concatenate([batches[s][r][1] for r in range(world)])).to(dev)
            ops.adam_begin_step(hyper)
            uq = ops.unique(slots, table_like=torch.empty((c, 0), device=dev))
            ops.sparse_lazy_adam(one.values[:c], m1[:c], v1[:c], hyper, g_all, None, uq)
        k1, v1d = one.get_data()
        order = torch.argsort(k1)
        k1, v1d = k1[order].cpu().numpy(), v1d[order].cpu().numpy()
        ks = np.concatenate([gt[0] for gt in gathered])
        vs = np.concatenate([gt[1] for gt in gathered])
        o = np.argsort(ks)
        ok &= bool(np.array_equal(ks[o], k1))
        err = float(np.abs(vs[o] - v1d).max()) if ks.size == k1.size else float("inf")
        ok &= err <= 1e-5 * float(np.abs(v1d).max()) and worst <= 1e-5 * float(np.abs(v1d).max()) + 1e-7
        ok &= all(gt[3] == 0 for gt in gathered)
        print("keys %d (per rank %s) | rows max|diff| %.3e | lookups max|diff| %.3e | flags %s | SHARDED HASH PARITY %s"
              % (k1.size, [int(gt[0].size) for gt in gathered], err, worst, [gt[3] for gt in gathered],
                 "OK" if ok else "FAILED"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
