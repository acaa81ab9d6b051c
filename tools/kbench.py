"""Per-kernel micro-benchmarks at BASELINE config sizes (CUDA events, L2 flushed between iterations).

    python tools/kbench.py [gather|unique|adam|ftrl|all] [--vocab V]
Prints one line per kernel: name, median ms, algorithmic MB, achieved GB/s, fraction of measured HBM peak.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindrec_b200 import ops  # noqa: E402


def hbm_peak():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        return json.load(open(p))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


_flush = None


def flush_l2():
    """Evict the L2 by READING 512 MB (clean lines).  A write flush leaves ~126 MB of dirty lines whose write-back
    lands inside the timed kernel: measured +~20 us on every kernel here (gather D=40 vs D=80 intercept)."""
    global _flush
    if _flush is None:
        _flush = torch.zeros(128 << 20, dtype=torch.int32, device="cuda")
    if os.environ.get("MREC_KBENCH_WRITE_FLUSH"):
        _flush.zero_()
    else:
        _flush.max()


GRAPH = True


def timeit(fn, iters=20, warmup=3):
    """Median device time of fn(); with GRAPH the call is captured once and replayed, so host launch
    latency of multi-kernel ops does not leak into the number."""
    for _ in range(warmup):
        fn()
    run = fn
    if GRAPH:
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        run = g.replay
    ts = []
    for _ in range(iters):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def report(name, ms, nbytes):
    peak, how = hbm_peak()
    gbs = nbytes / ms / 1e6
    print("%-28s %8.3f ms  %9.1f MB  %8.1f GB/s  %5.1f%% of %s HBM peak" %
          (name, ms, nbytes / 1e6, gbs, 100 * gbs / peak, how), flush=True)


def zipf_ids(b, f, vocab, seed=0):
    rng = np.random.default_rng(seed)
    ids = (rng.zipf(1.05, size=(b, f)) % vocab).astype(np.int32)
    ids[:, :13] = np.arange(13)
    return torch.from_numpy(ids).cuda()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="all")
    ap.add_argument("--vocab", type=int, default=33762616)
    ap.add_argument("--batch", type=int, default=16000)
    ap.add_argument("--dim", type=int, default=80)
    ap.add_argument("--eager", action="store_true")
    ap.add_argument("--seg-modes", default="", help="adam16: comma list of MREC_SEG_MODE[:MREC_SEG_P] to compare")
    a = ap.parse_args()
    global GRAPH
    GRAPH = not a.eager
    if a.what == "dense":
        return dense_bench(a.batch)
    b, f, d, v = a.batch, 39, a.dim, a.vocab
    n = b * f
    ids = zipf_ids(b, f, v)
    mask = torch.rand((b, f), device="cuda")
    table = torch.empty((v, d), device="cuda").normal_(0, 0.01)
    if a.what in ("gather", "all"):
        out = torch.empty((b, f * d), device="cuda")
        report("gather D=%d" % d, timeit(lambda: ops.gather(table, ids, out=out)), n * (8 * d + 4))
        report("gather_masked D=%d" % d, timeit(lambda: ops.gather_masked(table, ids, mask, out=out)),
               n * (8 * d + 8))
        uni = torch.randint(0, v, (b, f), device="cuda", dtype=torch.int32)
        report("gather uniform ids", timeit(lambda: ops.gather(table, uni, out=out)), n * (8 * d + 4))
        wt = torch.empty((v, 1), device="cuda").normal_(0, 0.01)
        bias = torch.zeros(1, device="cuda")
        wo = torch.empty((b, 1), device="cuda")
        report("gather_reduce D=1", timeit(lambda: ops.gather_reduce(wt, ids, mask, bias, out=wo)),
               n * 12 + b * 4)
    if a.what in ("unique", "all"):
        res = ops.UniqueResult(n, torch.int32, ids.device)
        ms = timeit(lambda: ops.unique(ids, table_like=table, result=res))
        u = int(res.count.item())
        report("unique_bounded N=%d U=%d" % (n, u), ms, n * 8 + u * 4)
        ms = timeit(lambda: ops.unique(ids, result=res))
        report("unique 32-bit", ms, n * 8 + u * 4)
    if a.what in ("adam", "all"):
        res = ops.unique(ids, table_like=table)
        u = int(res.count.item())
        m_, v_ = torch.zeros_like(table), torch.zeros_like(table)
        g = torch.randn((n, d), device="cuda")
        hyper = ops.adam_hyper(3.5e-4, loss_scale=1024.0)
        ops.adam_begin_step(hyper)
        ms = timeit(lambda: ops.sparse_lazy_adam(table, m_, v_, hyper, g, mask, res))
        report("segsum+lazy_adam U=%d" % u, ms, n * d * 4 + u * 7 * d * 4)
        gs = torch.empty((n, d), device="cuda")
        ms = timeit(lambda: ops.segment_sum(g, mask, res, out=gs))
        report("segment_sum", ms, n * d * 4 + u * d * 4)
        del m_, v_, g, gs
    if a.what in ("adam16", "all"):
        # the in-step form: fp16 gradient rows from the mixed-precision DenseLayer backward
        res = ops.unique(ids, table_like=table)
        u = int(res.count.item())
        m_, v_ = torch.zeros_like(table), torch.zeros_like(table)
        g = torch.randn((n, d), device="cuda").half()
        hyper = ops.adam_hyper(3.5e-4, loss_scale=1024.0)
        ops.adam_begin_step(hyper)
        gs = torch.empty((n, d), device="cuda")
        for mode in a.seg_modes.split(","):
            if mode:
                sm, sp, ps = (mode.split(":") + ["", ""])[:3]          # MREC_SEG_MODE[:MREC_ROWS_R[:MREC_ROWS_PER_SM]]
                os.environ["MREC_SEG_MODE"] = sm
                os.environ["MREC_ROWS_R"] = sp or "2"
                os.environ["MREC_ROWS_PER_SM"] = ps or "0"
            ms = timeit(lambda: ops.sparse_lazy_adam(table, m_, v_, hyper, g, mask, res))
            report("segsum+lazy_adam fp16 g U=%d [mode %s]" % (u, mode or "default"), ms, n * d * 2 + n * 12 + u * 6 * d * 4)
            ms = timeit(lambda: ops.segment_sum(g, mask, res, out=gs))
            report("segment_sum fp16 g [mode %s]" % (mode or "default"), ms, n * d * 2 + n * 12 + u * d * 4)
        os.environ.pop("MREC_SEG_MODE", None)
        os.environ.pop("MREC_ROWS_R", None)
        del m_, v_, g, gs
    if a.what in ("ftrl", "all"):
        res = ops.unique(ids, table_like=table)
        u = int(res.count.item())
        wt = torch.empty((v, 1), device="cuda").normal_(0, 0.01)
        acc, lin = torch.ones_like(wt), torch.zeros_like(wt)
        g = torch.randn((b, 1), device="cuda")
        hyper = ops.ftrl_hyper(5e-2, 1e-8, 1e-8, loss_scale=1024.0)
        ms = timeit(lambda: ops.sparse_ftrl(wt, acc, lin, hyper, g, mask, res))
        report("segsum+ftrl D=1 U=%d" % u, ms, u * 28 + n * 8)


def dense_bench(b=16000):
    """DenseLayer glue kernels at config-2 shapes (fp16)."""
    for n in (1024, 512, 256, 128):
        g = torch.randn((b, n), device="cuda").half()
        y = torch.relu(torch.randn((b, n), device="cuda")).half()
        gb = torch.empty(n, device="cuda")
        report("relu_bwd_bias fp16 N=%d" % n, timeit(lambda: ops.relu_bwd_bias(g, y, gb)), b * n * 6)
    k = 128
    h = torch.relu(torch.randn((b, k), device="cuda")).half()
    w = torch.randn(k, device="cuda").half()
    bias = torch.zeros(1, device="cuda").half()
    d16 = (torch.randn((b, 1), device="cuda") * 0.01).half()
    out = torch.empty((b, 1), device="cuda")
    report("dense_head_fwd K=128", timeit(lambda: ops.dense_head_fwd(h, w, bias, out=out)), b * k * 2 + b * 4)
    gw, gbh, gbp, gh = torch.empty(k, device="cuda"), torch.empty(1, device="cuda"), torch.empty(k, device="cuda"), torch.empty_like(h)
    report("dense_head_bwd K=128", timeit(lambda: ops.dense_head_bwd(d16, h, w, True, gw, gbh, gbp, out=gh)), b * k * 4 + b * 2)
    a = torch.randn((b, 1), device="cuda")
    lab = (torch.rand((b, 1), device="cuda") < 0.25).float()
    sens = torch.tensor([1024.0], device="cuda")
    o = ops.sigmoid_xent(a, a, lab, sens, half=True)
    report("sigmoid_xent B=%d" % b, timeit(lambda: ops.sigmoid_xent(a, a, lab, sens, out=o)), b * 26)


def interaction():
    b = 16384
    vx = torch.randn((b, 39, 16), device="cuda")
    g = torch.randn((b, 1), device="cuda")
    out = torch.empty((b, 1), device="cuda")
    dvx = torch.empty_like(vx)
    report("fm_fwd 16384x39x16", timeit(lambda: ops.fm_fwd(vx, out=out)), vx.numel() * 4 + b * 4)
    report("fm_bwd 16384x39x16", timeit(lambda: ops.fm_bwd(vx, g, out=dvx)), 2 * vx.numel() * 4 + b * 4)
    dp, layers = 3120, 6
    x0 = torch.randn((b, dp), device="cuda") * 0.1
    w = torch.randn((layers, dp), device="cuda") * 0.02
    bb = torch.randn((layers, dp), device="cuda") * 0.02
    gy = torch.randn((b, dp), device="cuda")
    y, p = ops.cross_fwd(x0, w, bb)
    dx, dw, db = ops.cross_bwd(x0, gy, w, bb, p)
    report("cross_fwd 16384x3120 L6", timeit(lambda: ops.cross_fwd(x0, w, bb, y=y, p=p)), 2 * x0.numel() * 4)
    report("cross_bwd 16384x3120 L6", timeit(lambda: ops.cross_bwd(x0, gy, w, bb, p, dx=dx, dw=dw, db=db)),
           3 * x0.numel() * 4)


def hash_bench():
    """BASELINE config 5 lookup shape: 16384 x 26 int64 Zipf keys over 2^40, dim 128."""
    from mindrec_b200 import hash as H
    n, d = 16384 * 26, 128
    rng = np.random.default_rng(0)
    keys = torch.from_numpy((rng.zipf(1.05, size=n) % (1 << 40)).astype(np.int64)).cuda()
    mp = H.MapParameter(key_dtype=torch.int64, value_shape=d, default_value="normal", capacity=1 << 21, device="cuda")
    mp.lookup_slots(keys)                                # insert pass
    resident = len(mp)
    slots = mp.lookup_slots(keys)
    out = torch.empty((n, d), device="cuda")
    report("hash find_or_insert (resident) n=%d U=%d" % (n, resident), timeit(lambda: mp.lookup_slots(keys)),
           n * (8 + 4))
    report("hash get = probe + gather D=128", timeit(lambda: ops.gather(mp.values, mp.lookup_slots(keys), out=out)),
           n * (8 + 2 * d * 4))
    fresh = torch.from_numpy(rng.integers(1 << 41, 1 << 42, size=n).astype(np.int64)).cuda()

    def insert_new():
        mp2 = insert_new.mp
        mp2.lookup_slots(fresh)
    insert_new.mp = H.MapParameter(key_dtype=torch.int64, value_shape=d, default_value="normal", capacity=1 << 21,
                                   device="cuda")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    insert_new()
    b.record()
    torch.cuda.synchronize()
    report("hash insert %d new keys + init rows" % n, a.elapsed_time(b), n * (8 + d * 4))
    print("load factor %.3f, overflow %s" % (len(insert_new.mp) / insert_new.mp.capacity, insert_new.mp.overflowed))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "interaction":
        interaction()
    elif len(sys.argv) > 1 and sys.argv[1] == "hash":
        GRAPH = False
        hash_bench()
    else:
        main()
