"""Kernel timeline of the graphed sharded Wide&Deep step (torchrun, one rank per GPU): torch.profiler (CUPTI activity
records — they cover graph-launched kernels) over a few steady-state steps; rank 0 prints one step as text
(start offset, duration, stream, kernel) and writes the Chrome trace next to it.

    torchrun --nproc-per-node 2 tools/timeline_sharded.py [--c5] > gpurun_out/timeline.txt
"""
import datetime
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindrec_b200 import peer_sharded, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    c5 = "--c5" in sys.argv
    if c5:
        from mindrec_b200 import multitable_sharded as M
        from tools import sharded_parity
        b, ft, fh = 16384, 26, 26
        rows_total = 20_000_000 * world
        step = M.ShardedMultitableStep(b, rows_total, dev, n_table_fields=ft, n_hash_fields=fh, hash_capacity=1 << 23, seed=1)
        host = sharded_parity.c5_batches(b, ft, fh, rows_total, 40, 20260105, rank, 4, alpha=1.05)
        ring = [tuple(torch.from_numpy(x).to(dev) for x in hb) for hb in host]
        step.capture(*ring[0], warmup=2)
        run = lambda i: step.replay(*ring[i % 4])
    else:
        cards = [max(3, int(c * world)) for c in synth.CARD_KAGGLE]
        step = peer_sharded.PeerShardedWideDeepStep(16000, synth.vocab_size(cards), 80, (1024, 512, 256, 128), dev, seed=1)
        gen = synth.CriteoSynth(16000, cards=cards, seed=20260101, rank=rank)
        ring = [tuple(torch.from_numpy(x).to(dev) for x in gen.next()) for _ in range(8)]
        step.capture(*ring[0], warmup=3)
        run = lambda i: step.replay(*ring[i % 8], next_batch=ring[(i + 1) % 8])
    for i in range(10):
        run(i)
    torch.cuda.synchronize()
    dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    n = 6
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(10, 10 + n):
            run(i)
        torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        path = "gpurun_out/timeline_%s_n%d.json" % ("c5" if c5 else "wd", world)
        prof.export_chrome_trace(path)
        ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
        ev.sort(key=lambda e: e["ts"])
        # cut into steps at the first kernel of graph a1 (gather_to_peers) / of the step
        # a step ends with the LazyAdam row update of the (last) table: cut behind it
        ends = [i for i, e in enumerate(ev) if "LazyAdamSink" in e["name"]]
        if c5:
            ends = ends[1::2]                       # two LazyAdam updates per step (table, MapParameter)
        starts = [i + 1 for i in ends]
        if len(starts) >= 4:
            lo, hi = starts[2], starts[3]
        else:
            lo, hi = 0, len(ev)
        t0 = ev[lo]["ts"]
        print("step of %.1f us, %d device activities" % (ev[hi]["ts"] - t0 if hi < len(ev) else -1, hi - lo))
        print("%9s %8s %6s  %s" % ("start_us", "dur_us", "stream", "name"))
        for e in ev[lo:hi]:
            print("%9.1f %8.1f %6s  %s" % (e["ts"] - t0, e["dur"], e.get("args", {}).get("stream", "?"), e["name"][:110]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
