"""The kernels whose DRAM traffic bench.py reports, launched once each (after one warm-up launch) at the bench shapes,
for one `ncu --set full` capture:

    ncu --set full --clock-control none --import-source on -k regex:'gather_rows|segsum_stage|rows_update|cross_fwd_k|cross_bwd_k|fm_kernel' \
        -o gpurun_out/r2_prof python tools/prof_targets.py
    python tools/ncu_traffic.py gpurun_out/r2_prof.ncu-rep gpurun_out/r2_prof_order.json > profiles/r2_traffic.json

The launch order (key per profiled launch) is written to gpurun_out/r2_prof_order.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mindrec_b200 import ops, synth  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    order = []
    b, f, d = 16000, 39, 80
    vocab = synth.vocab_size(synth.CARD_KAGGLE)
    gen = synth.CriteoSynth(b, cards=synth.CARD_KAGGLE, alpha=1.05, seed=20260101)
    ids, wts, _ = (torch.from_numpy(x).to(dev) for x in gen.next())
    wmv = torch.zeros((vocab, 3, d), device=dev)
    wmv[:, 0, :].normal_(0, 0.01)
    out = torch.empty((b * f, d), device=dev)
    uni = torch.randint(0, vocab, ids.shape, device=dev, dtype=torch.int32)
    for key, i in (("gather_zipf", ids), ("gather_uniform", uni)):
        for rep in range(2):
            ops.gather(wmv, i, out=out)
            order.append([key if rep else "warmup", "gather_rows"])
    # the dominant op: fp16 gradient rows, interleaved records, and the split layout for comparison
    g16 = torch.randn((b * f, d), device=dev, dtype=torch.float16)
    uq = ops.unique(ids, table_like=wmv)
    hyper = ops.adam_hyper(3.5e-4, eps=1e-8, loss_scale=1024.0, device=dev)
    ops.adam_begin_step(hyper)
    for rep in range(2):
        ops.sparse_lazy_adam(wmv, None, None, hyper, g16, wts.reshape(-1), uq)
        order += [[("sparse_lazy_adam" if rep else "warmup"), "segsum_stage"], [("sparse_lazy_adam" if rep else "warmup"), "rows_update"]]
    del wmv
    torch.cuda.empty_cache()
    w = torch.randn((vocab, d), device=dev) * 0.01
    m, v = torch.zeros_like(w), torch.zeros_like(w)
    for rep in range(2):
        ops.sparse_lazy_adam(w, m, v, hyper, g16, wts.reshape(-1), uq)
        order += [[("sparse_lazy_adam_split" if rep else "warmup"), "segsum_stage"],
                  [("sparse_lazy_adam_split" if rep else "warmup"), "rows_update"]]
    del w, m, v
    torch.cuda.empty_cache()
    # config 3: cross stack
    bb, dp, layers = 16384, 39 * 80, 6
    x0, dy = torch.randn((bb, dp), device=dev) * 0.1, torch.randn((bb, dp), device=dev) * 0.1
    cw, cb = torch.randn((layers, dp), device=dev) * 0.01, torch.randn((layers, dp), device=dev) * 0.01
    y, p = torch.empty_like(x0), torch.empty((bb, layers), device=dev)
    dw, db = torch.empty_like(cw), torch.empty_like(cw)
    for rep in range(2):
        ops.cross_fwd(x0, cw, cb, y=y, p=p)
        order.append(["cross_fwd" if rep else "warmup", "cross_fwd_k"])
    for rep in range(2):
        ops.cross_bwd(x0, dy, cw, cb, p, dx=y, dw=dw, db=db)
        order.append(["cross_bwd" if rep else "warmup", "cross_bwd_k"])
    # config 4: FM
    vx = torch.randn((bb, 39, 16), device=dev) * 0.1
    fo, gout, dvx = torch.empty((bb, 1), device=dev), torch.randn((bb, 1), device=dev), torch.empty_like(vx)
    for rep in range(2):
        ops.fm_fwd(vx, out=fo)
        order.append(["fm_fwd" if rep else "warmup", "fm_kernel"])
    for rep in range(2):
        ops.fm_bwd(vx, gout, out=dvx)
        order.append(["fm_bwd" if rep else "warmup", "fm_kernel"])
    torch.cuda.synchronize()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(order, open(os.path.join(ROOT, "gpurun_out", "r2_prof_order.json"), "w"))
    print("launched", len(order), "profiled kernels")


if __name__ == "__main__":
    main()
