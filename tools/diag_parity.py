"""Bisection of the multi-GPU parity check (torchrun, one rank per GPU): the Wide&Deep sharded step with / without CUDA
graphs and with / without the one-step-ahead key phase, against the unsharded cell."""
import datetime
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import sharded_parity  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    hidden = (1024, 512, 256, 128)
    for name, kw in (("eager", dict(graph=False, ahead=False)), ("graph_inline", dict(graph=True, ahead=False)),
                     ("graph_ahead", dict(graph=True, ahead=True)),
                     ("small_eager", dict(graph=False, ahead=False, batch=2000, rows_per_rank=50_021)),
                     ("small_ahead", dict(graph=True, ahead=True, batch=2000, rows_per_rank=50_021))):
        b = kw.pop("batch", 16000)
        res = sharded_parity.wide_deep(world, rank, dev, b, 39, 80, hidden, mixed=False, **kw)
        if rank == 0:
            short = {k: (v if not isinstance(v, dict) else {"max": v["max_abs_diff"], "out": v["outliers"], "l2": v["rel_l2"]})
                     for k, v in res.items()}
            print(name, json.dumps(short), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
