"""NVLink peer-store bandwidth as the exchange kernels see it (torchrun, >= 2 ranks): every rank pushes rows into the
NEXT rank's IPC buffer, all ranks at once (so every link carries traffic in both directions, like a step does).
  memcpy   cudaMemcpyAsync on the peer-mapped pointer (copy engine)
  push     mrec_push_rows_to_peers, 16-byte stores, contiguous source rows
  gather   mrec_gather_to_peers, random source rows of a 2 M-row table
Prints GB/s per direction per GPU."""
import datetime
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindrec_b200 import ops, peer_sharded  # noqa: E402
from mindrec_b200.sharded import _RawCuda  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    dim, max_rows = 80, 1 << 20
    arena = peer_sharded._IpcArena(None)
    inbox = arena.alloc(dev)("inbox", (max_rows, dim), torch.float32)
    inbox.zero_()
    torch.cuda.synchronize()
    dist.barrier()
    base = arena.exchange()
    nxt = (rank + 1) % world
    peer = torch.as_tensor(_RawCuda(base["inbox"][nxt], (max_rows, dim)), device=dev)
    ptrs = torch.tensor(base["inbox"], dtype=torch.int64, device=dev)
    table = torch.randn((2_000_003, dim), device=dev)
    src = torch.randn((max_rows, dim), device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    cap_like, mod_none = torch.empty((max_rows, 0), device=dev), torch.empty((0, 0), device=dev)
    zeros_g = torch.zeros(world, dtype=torch.int32, device=dev)

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for rows in (34_000, 68_000, 137_000, 274_000, 1 << 20):
        nbytes = rows * dim * 4
        # all of my rows go to rank `nxt`: bounds = 0 for owners <= nxt, rows beyond
        bounds = torch.tensor([0 if o <= nxt else rows for o in range(world + 1)], dtype=torch.int32, device=dev)
        rows_idx = torch.randint(0, table.shape[0], (rows,), dtype=torch.int32, device=dev)
        # gather_to_peers: rows of source s are rows[src_off[s]:src_off[s+1]] -> everything belongs to source `nxt`
        src_off = torch.tensor([0 if s <= nxt else rows for s in range(world + 1)], dtype=torch.int32, device=dev)
        res = {
            "memcpy": timed(lambda: peer[:rows].copy_(src[:rows], non_blocking=True)),
            "push": timed(lambda: ops.push_rows_to_peers(src[:rows], bounds, zeros_g, ptrs, cap_like, mod_none, err)),
            "gather": timed(lambda: ops.gather_to_peers(table, rows_idx, ptrs, zeros_g, src_off)),
            "local_gather": timed(lambda: ops.gather(table, rows_idx, out=src[:rows])),
        }
        if rank == 0:
            print("rows %8d  %6.1f MB  " % (rows, nbytes / 1e6) +
                  "  ".join("%s %7.1f us %6.0f GB/s" % (k, v * 1e3, nbytes / v / 1e6) for k, v in res.items()), flush=True)
    assert int(err.item()) == 0
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
