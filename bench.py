#!/usr/bin/env python
"""Headline benchmark: Wide&Deep training samples/s on synthetic Criteo-shaped data (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

N = 1 measures BASELINE config 2: Wide&Deep, one B200, Criteo-Kaggle vocabulary (33 762 616 rows, dim 80,
fp32 tables), batch 16000, LazyAdam on the deep table + FTRL on the wide table, DenseLayers in fp16 like the
reference's `use_mixed_precision`.  A step = lookup (gather + mask + wide reduce) -> DenseLayer fwd/bwd
(cuBLAS) -> loss -> sparse-gradient dedup -> fused segment-sum + LazyAdam / FTRL row updates -> dense Adam.
N > 1 row-shards the tables over the ranks with all-to-all exchange (mindrec_b200.sharded), weak scaling.

`--impl reference` times the reference's CPU path as restated in oracle/ (MindSpore itself cannot be
installed here) on the host cores.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

_OUT = sys.stdout

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "wide_deep_train_samples_per_s"
UNIT = "samples/s"
BATCH = 16000
FIELDS = 39
EMB = 80
HIDDEN = (1024, 512, 256, 128)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--vocab-scale", type=float, default=1.0, help="shrink the Kaggle cardinalities (debug)")
    ap.add_argument("--alpha", type=float, default=1.05)
    ap.add_argument("--layout", default=os.environ.get("MREC_BENCH_LAYOUT", "interleaved"), choices=["interleaved", "split"],
                    help="deep table of the single-GPU step: w|m|v records [V,3,D] (one DRAM burst per row update) or "
                         "three [V,D] arrays")
    ap.add_argument("--c5-rows-per-gpu", type=int, default=88_000_000,
                    help="config 5: rows of the dim-128 table per GPU (w + LazyAdam m, v in fp32 = 1.5 KB per row)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the config-3 / config-4 / config-5 blocks")
    ap.add_argument("--exchange", default=os.environ.get("MREC_BENCH_EXCHANGE", "device"), choices=["nccl", "device"],
                    help="N>1: nccl = all-to-all with host-side split sizes; device = device-driven peer stores, "
                         "whole step in one CUDA graph")
    ap.add_argument("--cpu-sample-batch", type=int, default=2000)
    ap.add_argument("--cpu-vocab", type=int, default=4000000)
    return ap.parse_args()


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def scaled_cards(scale):
    from mindrec_b200 import synth
    return [max(3, int(c * scale)) for c in synth.CARD_KAGGLE]


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference path restated in oracle/ (kind = "port"), bounded sample of the same workload
# --------------------------------------------------------------------------------------------------
def cpu_config1_run(steps=20, warmup=3):
    """BASELINE configs[0], faithfully: Wide&Deep, batch 1000, 39 fields, vocab 200 000 (the reference's own
    cardinality list, models/wide_deep/src/datasets.py:354-379), dim 80, fp32, single process on the host cores."""
    import numpy as np
    from mindrec_b200 import synth
    from oracle import ref_c
    cores = ref_c.use_all_cores()
    b, vocab = 1000, 200000
    model = ref_c.WideDeepCpu(vocab, EMB, hidden=HIDDEN, fields=FIELDS, seed=0)
    gen = synth.CriteoSynth(b, cards=synth.CARD_REFERENCE, alpha=1.05, seed=20260101, vocab_pad=vocab)
    batches = [gen.next() for _ in range(4)]
    for i in range(warmup):
        model.step(*batches[i % 4])
    t0 = time.perf_counter()
    for i in range(steps):
        model.step(*batches[i % 4])
    dt = time.perf_counter() - t0
    return {"value": b * steps / dt, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_step": 1e3 * dt / steps,
            "sample": "BASELINE config 1 in full: batch 1000, 39 fields, vocab 200000, dim 80, fp32, %d steps after "
                      "%d warm-up" % (steps, warmup)}


def cpu_baseline_run(args, steps, warmup):
    import numpy as np
    from mindrec_b200 import synth
    from oracle import ref_c
    ref_c.use_all_cores()            # torchrun exports OMP_NUM_THREADS=1: pin the team to every host core
    cards = scaled_cards(args.vocab_scale)
    full_vocab = synth.vocab_size(cards)
    vocab = min(full_vocab, args.cpu_vocab)
    b = args.cpu_sample_batch
    model = ref_c.WideDeepCpu(vocab, EMB, hidden=HIDDEN, fields=FIELDS, seed=0)
    gen = synth.CriteoSynth(b, cards=cards, alpha=args.alpha, seed=20260101)
    batches = []
    for _ in range(4):
        ids, wts, label = gen.next()
        batches.append((np.ascontiguousarray(ids % vocab), wts, label))
    for i in range(warmup):
        model.step(*batches[i % 4])
    t0 = time.perf_counter()
    for i in range(steps):
        model.step(*batches[i % 4])
    dt = time.perf_counter() - t0
    sample = ("batch %d of the config-2 stream (Zipf %.2f, 39 fields, dim 80, MLP 3120-1024-512-256-128-1 fp32); "
              "tables held at %d rows (ids mod rows) to bound host memory; %d steps after %d warm-up" %
              (b, args.alpha, vocab, steps, warmup))
    return {"value": b * steps / dt, "unit": UNIT, "cores": ref_c.threads(), "kind": "port",
            "sample": sample, "ms_per_step": 1e3 * dt / steps}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.gpus > 1 and args.vocab_scale == 1.0:
        args.vocab_scale = float(args.gpus)      # same workload description as our arm at N > 1
    steps = max(1, min(args.steps, 20))
    warmup = max(1, min(args.warmup, 3))
    cb = cpu_baseline_run(args, steps, warmup)
    c1 = cpu_config1_run()
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "config1": c1,
        "note": "reference's CPU path as restated in oracle/ (C/OpenMP + numpy BLAS); MindSpore is not installable here",
    }
    print(json.dumps(line), file=_OUT, flush=True)


def workload_config(args, n_gpus):
    from mindrec_b200 import synth
    cards = scaled_cards(args.vocab_scale)
    return {
        "workload": "BASELINE config 2: Wide&Deep, Criteo-Kaggle vocab, dim 80, fp32 tables, batch 16000/GPU, "
                    "LazyAdam deep + FTRL wide sparse updates" + ("" if n_gpus == 1 else
                                                                  ", tables row-sharded over %d GPUs (all-to-all)" % n_gpus),
        "global_batch": args.batch * n_gpus, "batch_per_gpu": args.batch, "fields": FIELDS, "emb_dim": EMB,
        "vocab_rows": synth.vocab_size(cards), "zipf_alpha": args.alpha, "mlp": [FIELDS * EMB] + list(HIDDEN) + [1],
        "mlp_dtype": "fp16 (use_mixed_precision)", "loss_scale": 1024,
        "parallelism": "single" if n_gpus == 1 else "row-sharded embeddings (%s) + data-parallel MLP x%d" % (
            "device-driven NVLink peer stores, step in one CUDA graph" if args.exchange == "device" else "a2a", n_gpus),
        "l2": "per-step working set (>= 0.8 GB of table rows, activations and gradients) exceeds the 126 MB L2; "
              "batches rotate over a ring of 8",
    }


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  In-process NVML polling (a thread, one sample every
    2 ms: a 20-step region of ~15 ms still gets several samples); `nvidia-smi -lms` is the fallback when the NVML
    binding is missing.  start() / stop() bracket the timed region; samples outside it are dropped."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index = index
        self.p = self.f = self.thread = None
        self.samples, self.run = [], False
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = None
            try:                                      # CUDA_VISIBLE_DEVICES may renumber: match by PCI bus id
                import torch
                bus = torch.cuda.get_device_properties(index).pci_bus_id
                dom = torch.cuda.get_device_properties(index).pci_domain_id
                dev = torch.cuda.get_device_properties(index).pci_device_id
                uuid = "%08X:%02X:%02X.0" % (dom, bus, dev)
                self.h = pynvml.nvmlDeviceGetHandleByPciBusId(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._one()                               # prove the calls work before relying on them
            self.samples = []
        except Exception:
            self.nvml = None

    def _one(self):
        n = self.nvml
        mhz = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
        try:
            why = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            why = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        self.samples.append((mhz, why))

    def _poll(self):
        while self.run:
            try:
                self._one()
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        if self.nvml is not None:
            import threading
            self.run = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(self.index)], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.nvml is not None:
            try:
                self._one()                           # at least one sample taken before the region's closing sync returns
            except Exception:
                pass
            self.run = False
            if self.thread is not None:
                self.thread.join(timeout=2)
            sm = [m for m, _ in self.samples]
            reasons = sorted({name for _, why in self.samples for name, bit in self.BITS if why & bit})
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, val in zip(names, r[5:9]):
                if val.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# --------------------------------------------------------------------------------------------------
# roofline of the dominant product op, measured live (CUDA graph replay, L2 flushed, CUDA events)
# --------------------------------------------------------------------------------------------------
def measure_dominant_op(step, batches, b, reps=10):
    """mrec_sparse_lazy_adam on the step's own deep table: segment-sum of N fp16 gradient rows over the
    step's dedup result + LazyAdam update of the U unique rows (segsum_stage_kernel + rows_update_kernel<LazyAdamSink>,
    2 launches).  Algorithmic bytes (with the fp32 cast fused into the load): N*D*2 read + U*7*D*4.

    Cold-cache without a write flush: the op runs over a ring of len(batches) different batches, each with its own
    gradient buffer and dedup result (per-call working set ~0.37 GB > the 126 MB L2, so nothing of call k survives
    until call k comes round again except the Zipf-hot rows every batch shares, as in the real step).  A write
    flush would leave ~126 MB of dirty lines whose write-back is charged to the timed op (tools/kbench.py)."""
    import torch
    from mindrec_b200 import ops
    model = step.model
    d = model.emb_dim
    packed = model.embedding_table.packed is not None      # w | m | v interleaved per row ([V,3,D])
    table = model.embedding_table.kernel_arg
    m, v = (None, None) if packed else (step.optimizer_d.moment1[0], step.optimizer_d.moment2[0])
    hyper = step.optimizer_d.hyper
    sets, alg, us = [], 0, []
    for ids, wts, _ in batches:
        n = ids.numel()
        uq = ops.unique(ids, table_like=table, ws_tag="unique_roofline_%d" % len(sets))
        u = int(uq.count.item())
        g16 = torch.randn((n, d), device=ids.device, dtype=torch.float16)
        sets.append((g16, wts.reshape(-1), uq))
        alg += n * d * 2 + u * 7 * d * 4
        us.append(u)

    def ring():
        for g16, mask, uq in sets:
            ops.sparse_lazy_adam(table, m, v, hyper, g16, mask, uq)
    for _ in range(2):
        ring()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ring()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    k = len(sets)
    ms = e0.elapsed_time(e1) / (reps * k)
    alg //= k
    peak, how = hbm_peak()
    gbs = alg / ms / 1e6
    # dram__bytes_read + dram__bytes_write of the op's two kernels, from this round's ncu --set full capture
    traffic = _traffic("sparse_lazy_adam" if packed else "sparse_lazy_adam_split")
    if traffic is None:
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_roofline_traffic.json")))["traffic_bytes_per_launch"]
        except Exception:
            pass
    return {"kernel": "mrec_sparse_lazy_adam = segsum_stage_kernel<__half> + "
                      "rows_update_kernel<F8,LazyAdamSink> (256-bit rows; 2 launches, timed as one op)",
            "table_layout": "interleaved w|m|v records [V,3,D]" if packed else "split w[V,D], m[V,D], v[V,D]",
            "bound": "hbm", "achieved": round(gbs, 1), "peak": peak,
            "peak_source": "MEASURED_PEAKS.json (measured)" if how == "measured" else "fallback 6650",
            "unit": "GB/s", "frac": round(gbs / peak, 4), "traffic": traffic, "algorithmic_bytes": alg,
            "unique_rows": int(sum(us) / k), "lookups": batches[0][0].numel(), "ms": round(ms, 4),
            "how": "CUDA graph of the op over a ring of %d different batches (inputs > L2, no flush), %d replays "
                   "back to back, CUDA events, mean per call" % (k, reps)}


def _traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from this round's `ncu --set full`
    capture (profiles/r2_traffic.json, written by tools/ncu_metrics.py from the .ncu-rep of the same command)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))[name]["traffic_bytes_per_launch"]
    except Exception:
        return None


def _time_ring(calls, reps=10):
    """Mean device time (ms) of one call: the calls (each on its own buffers; together larger than the 126 MB L2, so
    no flush is needed) are captured back to back in one CUDA graph, replayed `reps` times between two CUDA events."""
    import torch
    for _ in range(2):
        for c in calls:
            c()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for c in calls:
            c()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * len(calls))


def _roof(kernel, alg_bytes, ms, traffic_key=None, **extra):
    peak, how = hbm_peak()
    gbs = alg_bytes / ms / 1e6
    out = {"kernel": kernel, "bound": "hbm", "achieved": round(gbs, 1), "peak": peak,
           "peak_source": "MEASURED_PEAKS.json (measured)" if how == "measured" else "fallback 6650", "unit": "GB/s",
           "frac": round(gbs / peak, 4), "traffic": _traffic(traffic_key) if traffic_key else None,
           "algorithmic_bytes": int(alg_bytes), "ms": round(ms, 4)}
    out.update(extra)
    return out


def measure_gather(table, zipf_ids):
    """BASELINE metric "embedding gather HBM GB/s": mrec_gather of the config-2 deep table (33.76 M x 80 fp32),
    N = 624 000 lookups, algorithmic bytes N*(8D+4).  Two id streams: the step's own Zipf(1.05) ids (most
    reads hit the L2: ~120 k distinct rows) and ids uniform over the table (every row comes from HBM) — the uniform
    figure is the honest HBM number.  Ring of 4 id sets, each call writes its own [N, D] output (200 MB)."""
    import torch
    from mindrec_b200 import ops
    dev = table.device
    v, d = table.shape[0], table.shape[-1]
    n = zipf_ids[0].numel()
    outs = [torch.empty((n, d), dtype=torch.float32, device=dev) for _ in range(2)]
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    uni = [torch.randint(0, v, zipf_ids[0].shape, device=dev, dtype=torch.int32, generator=gen) for _ in range(4)]
    alg = n * (8 * d + 4)
    res = {}
    for name, ids in (("zipf", zipf_ids), ("uniform", uni)):
        calls = [(lambda i=i, t=t: ops.gather(table, t, out=outs[i & 1])) for i, t in enumerate(ids)]
        ms = _time_ring(calls)
        res[name] = _roof("gather_rows_kernel (mrec_gather, D=80, %s ids)" % name, alg, ms, "gather_" + name,
                          lookups=n, distinct_rows=int(torch.unique(ids[0]).numel()))
    return res


def measure_c3(args, dev, steps):
    """BASELINE configs[2]: Deep&Cross, 6 cross layers, dim 80 (row = 39*80 = 3120 fp32), batch 16384, vocab 200 000,
    fp32 towers (the reference's convert_dtype=False), Adam on every weight: full training steps through
    interaction.DeepCrossTrainStep + the cross kernels in isolation (fwd 2*B*D'*4, bwd 3*B*D'*4 algorithmic bytes)."""
    import torch
    from mindrec_b200 import interaction, ops, synth
    b, d, layers, vocab = 16384, EMB, 6, 200000
    cfg = interaction.DeepCrossConfig(batch_size=b, field_size=FIELDS, vocab_size=vocab, emb_dim=d,
                                      deep_layer_dim=(1024, 1024), cross_layer_num=layers)
    model = interaction.DeepCrossModel(cfg, device=dev)
    step = interaction.DeepCrossTrainStep(interaction.DeepCrossNetWithLoss(model))
    gen = synth.CriteoSynth(b, cards=synth.CARD_REFERENCE, alpha=args.alpha, seed=20260103, vocab_pad=vocab)
    batches = [tuple(torch.from_numpy(x).to(dev) for x in gen.next()) for _ in range(4)]
    for i in range(3):
        loss = step(*batches[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = step(*batches[i % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    dp = FIELDS * d
    xs = [torch.randn((b, dp), device=dev) * 0.1 for _ in range(2)]
    dys = [torch.randn((b, dp), device=dev) * 0.1 for _ in range(2)]
    ys = [torch.empty((b, dp), device=dev) for _ in range(2)]
    ps = [torch.empty((b, layers), device=dev) for _ in range(2)]
    dw, db = torch.empty((layers, dp), device=dev), torch.empty((layers, dp), device=dev)
    w, bb = model.cross_weight, model.cross_bias
    fwd = _time_ring([(lambda i=i: ops.cross_fwd(xs[i], w, bb, y=ys[i], p=ps[i])) for i in range(2)])
    bwd = _time_ring([(lambda i=i: ops.cross_bwd(xs[i], dys[i], w, bb, ps[i], dx=ys[i], dw=dw, db=db)) for i in range(2)])
    return {"workload": "BASELINE config 3: Deep&Cross, 6 cross layers, dim 80, batch 16384, vocab 200000, fp32, "
                        "Adam (dense-equivalent) on the table", "samples_per_s": b / (ms * 1e-3), "ms_per_step": ms,
            "steps": steps, "mode": "eager", "final_loss": float(loss),
            "cross_fwd": _roof("cross_fwd_kernel (6 layers fused)", 2 * b * dp * 4, fwd, "cross_fwd"),
            "cross_bwd": _roof("cross_bwd_kernel (6 layers fused)", 3 * b * dp * 4, bwd, "cross_bwd")}


def measure_c4(args, dev, steps):
    """BASELINE configs[3]: DeepFM, FM second order over 39 fields, dim 16, batch 16384, vocab 184 965: full training
    steps through interaction.DeepFMTrainStep + the FM kernels in isolation (fwd B*F*D*4, bwd 2*B*F*D*4 bytes; a ring
    of 8 inputs of 41 MB each so that the 126 MB L2 cannot hold them)."""
    import torch
    from mindrec_b200 import interaction, ops, synth
    b, d, vocab = 16384, 16, 184965
    cfg = interaction.DeepFMConfig(batch_size=b, data_field_size=FIELDS, data_vocab_size=vocab, data_emb_dim=d)
    model = interaction.DeepFMModel(cfg, device=dev)
    step = interaction.DeepFMTrainStep(interaction.DeepFMNetWithLoss(model, l2_coef=cfg.l2_coef), lr=cfg.learning_rate,
                                       eps=cfg.epsilon, loss_scale=cfg.loss_scale)
    gen = synth.CriteoSynth(b, cards=synth.CARD_REFERENCE, alpha=args.alpha, seed=20260104, vocab_pad=vocab)
    batches = [tuple(torch.from_numpy(x).to(dev) for x in gen.next()) for _ in range(4)]
    for i in range(3):
        loss = step(*batches[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = step(*batches[i % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ring = 8
    vxs = [torch.randn((b, FIELDS, d), device=dev) * 0.1 for _ in range(ring)]
    outs = [torch.empty((b, 1), device=dev) for _ in range(ring)]
    gout = torch.randn((b, 1), device=dev)
    dvx = [torch.empty((b, FIELDS, d), device=dev) for _ in range(ring)]
    fwd = _time_ring([(lambda i=i: ops.fm_fwd(vxs[i], out=outs[i])) for i in range(ring)])
    bwd = _time_ring([(lambda i=i: ops.fm_bwd(vxs[i], gout, out=dvx[i])) for i in range(ring)])
    vol = b * FIELDS * d * 4
    return {"workload": "BASELINE config 4: DeepFM, FM second order over 39 fields, dim 16, batch 16384, vocab 184965",
            "samples_per_s": b / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "mode": "eager",
            "final_loss": float(loss),
            "fm_fwd": _roof("fm_fwd_kernel", vol + b * 4, fwd, "fm_fwd"),
            "fm_bwd": _roof("fm_bwd_kernel", 2 * vol + b * 4, bwd, "fm_bwd")}


def measure_torch_free(args, cards, host_batches, steps):
    """BASELINE config 2 through the torch-free host path (mindrec_b200.rt_wide_deep.WideDeepRT: runtime buffers and
    streams, aot kernels, cuBLASLt through mrec_rt_gemm, one CUDA graph per step) — the same step as the headline, driven
    without a torch tensor, allocator, stream or GEMM.  (This process has torch loaded for the other blocks;
    tests/test_rt_wide_deep_gpu.py runs the class with the torch import blocked.)"""
    from mindrec_b200 import runtime, synth
    from mindrec_b200.rt_wide_deep import WideDeepRT
    dev = runtime.Device(int(os.environ.get("LOCAL_RANK", "0")))
    step = WideDeepRT(dev, args.batch, FIELDS, synth.vocab_size(cards), EMB, HIDDEN, mixed=True, sens=1024.0, seed=1)
    ring = [tuple(dev.from_numpy(x.numpy()) for x in hb) for hb in host_batches]
    step.set_inputs(*ring[0])
    step.capture(warmup=2)
    r = len(ring)
    for i in range(5):
        step.train_step(*ring[i % r], next_batch=ring[(i + 1) % r])
    dev.synchronize()
    e0, e1 = runtime.Event(timing=True), runtime.Event(timing=True)
    e0.record(dev.stream)
    for i in range(5, 5 + steps):
        step.train_step(*ring[i % r], next_batch=ring[(i + 1) % r])
    e1.record(dev.stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / steps
    loss = float(step.loss_out[1].item())
    out = {"class": "mindrec_b200.rt_wide_deep.WideDeepRT", "ms_per_step": ms, "samples_per_s": args.batch / (ms * 1e-3),
           "steps": steps, "gpu_launches_per_step": step.launches_per_step, "final_loss": loss,
           "host": "ctypes over mrec_rt_* + the aot entry points; DenseLayer GEMMs: cuBLASLt via mrec_rt_gemm"}
    del step, ring
    return out


def measure_c5(args, world, rank, dev, group_one):
    """BASELINE configs[4]: the multitable model's emb128 table (dim 128, + its dim-1 wide vector) row-sharded over the
    ranks with rows-per-GPU fixed, plus a hash-sharded MapParameter (int64 Zipf keys over 2^40, admission on the second
    sighting, eviction of keys unseen for 8 steps) in the same step; 5 x 1024 DenseLayers data parallel; batch 16384 x
    (26 table fields + 26 dynamic-feature fields) per GPU; row-sparse LazyAdam / FTRL (mindrec_b200.multitable_sharded).
    Returns the block for the bench line (rank 0) — every rank must call it."""
    import torch
    import torch.distributed as dist
    from mindrec_b200 import multitable_sharded as M
    from tools import sharded_parity
    b, ft, fh = 16384, 26, 26
    out = {}
    # parity first (small table): G ranks vs the same step on a one-rank group fed the concatenated batch
    out["parity_check"] = {"fp32": sharded_parity.multitable(world, rank, dev, b, mixed=False, group_one=group_one,
                                                             n_table_fields=ft, n_hash_fields=fh, evict_filter_value=1,
                                                             evict_every=2, hash_capacity=1 << 22),
                           "fp16": sharded_parity.multitable(world, rank, dev, b, mixed=True, group_one=group_one,
                                                             n_table_fields=ft, n_hash_fields=fh, evict_filter_value=1,
                                                             evict_every=2, hash_capacity=1 << 22)}
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info(dev)
    row_bytes = (128 + 1) * 4 * 3                       # w, m, v (+ the wide vector's w, accum, linear)
    rows = min(args.c5_rows_per_gpu, int((free - (40 << 30)) // row_bytes))
    t = torch.tensor([rows], device=dev, dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    rows = int(t.item())
    if rows < (1 << 20):
        return dict(out, skipped="not enough free HBM for the config-5 table (%d rows per GPU)" % rows)
    rows_total = rows * world
    step = M.ShardedMultitableStep(b, rows_total, dev, n_table_fields=ft, n_hash_fields=fh, hash_capacity=1 << 23, seed=1)
    host = sharded_parity.c5_batches(b, ft, fh, rows_total, 40, 20260105, rank, 4, alpha=args.alpha)
    ring = [tuple(torch.from_numpy(x).to(dev) for x in hb) for hb in host]
    step.capture(*ring[0], warmup=2)
    k = max(3, min(args.steps, 20))
    for i in range(3):
        step.replay(*ring[i % 4])
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3, 3 + k):
        loss = step.replay(*ring[i % 4])[0]
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / k
    tt = torch.tensor([ms], device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    st = step.exchange_stats()
    flags = step.error_flags()
    agg = torch.tensor([float(st["nvlink_bytes_out"]), float(st["table"]["unique_keys"]), float(st["table"]["rows_owned"]),
                        float(st["hash"]["unique_keys"]), float(st["hash"]["rows_owned"]), float(flags),
                        float(len(step.hash.rk.table)), float(loss)], device=dev)
    mx = agg.clone()
    dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    a = (agg / world).tolist()
    out.update({
        "workload": "BASELINE config 5: emb128 table (dim 128) + wide vector row-sharded, %d rows per GPU = %d rows on %d "
                    "GPUs (fp32 w + LazyAdam m, v resident), MapParameter hash-sharded (int64 Zipf keys over 2^40, "
                    "permit 2, evict 8), batch 16384 x (26 + 26) fields per GPU, DenseLayers 6656-1024x5-1 fp16 data "
                    "parallel, row-sparse LazyAdam / FTRL" % (rows, rows_total, world),
        "samples_per_s": b * world / (ms * 1e-3), "ms_per_step": ms, "steps": k, "rows_per_gpu": rows,
        "rows_total": rows_total, "gpu_launches_per_step": step.launches_per_step,
        "exchange": {"nvlink_bytes_out_per_rank_per_step": int(a[0]), "table_unique_keys_per_rank": int(a[1]),
                     "table_rows_owned_per_rank": int(a[2]), "hash_unique_keys_per_rank": int(a[3]),
                     "hash_rows_owned_per_rank": int(a[4]),
                     "avg_gbs_per_direction_over_step": round(a[0] / (ms * 1e-3) / 1e9, 1),
                     "nvlink_peak_gbs_per_direction": 900, "nvlink_frac": round(a[0] / (ms * 1e-3) / 1e9 / 900, 4)},
        "error_flags": int(mx[5].item()), "hash_resident_keys_per_rank": int(a[6]), "final_loss": a[7]})
    step.close()
    del step
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from mindrec_b200 import _lib, cells, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=150))
    _lib.lib()  # fail loudly if the CUDA extension is missing

    if world > 1 and args.vocab_scale == 1.0:
        args.vocab_scale = float(world)          # weak scaling: rows per GPU fixed at the config-2 table size
    cards = scaled_cards(args.vocab_scale)
    vocab = synth.vocab_size(cards)
    b = args.batch
    if world == 1:
        cfg = cells.WideDeepConfig(batch_size=b, field_size=FIELDS, vocab_size=vocab, emb_dim=EMB,
                                   deep_layer_dim=HIDDEN, use_mixed_precision=True, sparse=True, seed=1,
                                   interleave_adam_state=(args.layout == "interleaved"))
        model = cells.WideDeepModel(cfg, device=dev)
        step = cells.TrainStepWrap(cells.NetWithLossClass(model, cfg), sens=1024.0, sparse=True, lazy_adam=True)
    else:
        step = None
        if args.exchange == "device":
            from mindrec_b200 import peer_sharded
            try:
                step = peer_sharded.PeerShardedWideDeepStep(b, vocab, EMB, HIDDEN, dev, seed=1)
            except peer_sharded.PeerMemoryUnavailable as exc:      # raised on every rank together
                if rank == 0:
                    print("bench.py: %s -> falling back to the NCCL exchange" % (exc,), file=sys.stderr)
                args.exchange = "nccl"
        if step is None:
            from mindrec_b200 import sharded
            step = sharded.build_sharded_wide_deep(b, vocab, EMB, HIDDEN, dev, seed=1)

    ring = 8
    gen = synth.CriteoSynth(b, cards=cards, alpha=args.alpha, seed=20260101, rank=rank)
    host = []
    for _ in range(ring):
        ids, wts, label = gen.next()
        host.append(tuple(torch.from_numpy(x).pin_memory() for x in (ids, wts, label)))
    devb = [tuple(x.to(dev) for x in hb) for hb in host]
    h2d_bytes = sum(x.numel() * x.element_size() for x in host[0])

    launches0 = _lib.launch_count()
    step.capture(*devb[0], warmup=3)
    per_step_launches = (_lib.launch_count() - launches0) // 4  # 3 warm-up steps + 1 captured
    per_step_launches = getattr(step, "launches_per_step", per_step_launches)   # counted while capturing, if offered

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: W warm-up + K timed graph replays -------------------------
    for i in range(max(3, args.warmup)):
        step.replay(*devb[i % ring], next_batch=devb[(i + 1) % ring])
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    w0 = max(3, args.warmup)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]      # per-step spread (p10/p50/p90)
    e0.record()
    for i in range(w0, w0 + args.steps):
        step.replay(*devb[i % ring], next_batch=devb[(i + 1) % ring])
        marks[i - w0].record()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    step_ms = None
    try:
        edges = [e0] + marks
        per = sorted(edges[j].elapsed_time(edges[j + 1]) for j in range(args.steps))
        pick = lambda q: round(per[min(len(per) - 1, int(q * len(per)))], 4)
        step_ms = {"p10": pick(0.10), "p50": pick(0.50), "p90": pick(0.90), "rank": rank}
    except Exception:                                      # the spread is informative only
        step_ms = None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = b * world / (ms * 1e-3)

    # ---- end to end through the public API: pinned host batch -> H2D -> step -> loss D2H ---------
    # Every step: the step's inputs go pinned host -> device, the step's loss comes device -> pinned host and
    # is read by the host.  The read is pipelined by one step (the loss of step i is consumed while step i+1
    # runs), as a training loop that logs its loss does; nothing is skipped or cached.
    loss_ring = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
    if hasattr(step, "_pending"):
        step._pending = None                                        # drop the look-ahead of the resident loop
        step._pending_src = None
    loss_seen = []

    def e2e_step(i):
        out = step.replay(*host[i % ring], next_batch=host[(i + 1) % ring])   # pinned H2D copies + step
        loss_ring[i & 1].copy_(out[0].reshape(1), non_blocking=True)          # loss D2H
        loss_ev[i & 1].record()

    def e2e_read(i):
        loss_ev[i & 1].synchronize()
        loss_seen.append(float(loss_ring[i & 1][0]))

    for i in range(3):
        e2e_step(i)
        e2e_read(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3, 3 + args.steps):
        e2e_step(i)
        if i > 3:
            e2e_read(i - 1)
    e2e_read(3 + args.steps - 1)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    final_loss = loss_seen[-1]
    assert len(loss_seen) == args.steps + 3

    # ---- per-phase breakdown + roofline of the dominant product kernel (eager, CUDA events) ------
    breakdown, roofline = {}, None
    if hasattr(step, "profile"):
        step.profile = cells.StepProfile()
        n_prof = min(args.steps, 20)
        for i in range(n_prof):
            step(*devb[i % ring])
        tot = step.profile.totals()
        step.profile = None
        breakdown = {k: round(statistics.median(v), 4) for k, v in tot.items()}
    roofline_gather, configs = None, {}
    if world == 1:
        roofline = measure_dominant_op(step, devb[:4], b)
        roofline_gather = measure_gather(step.model.embedding_table.kernel_arg, [x[0] for x in devb[:4]])

    exchange = None
    if world > 1 and hasattr(step.tables, "rk"):
        # NVLink traffic of the last step, from the device-side plan (a2a bytes = remote unique keys x
        # (key + row) forward and x row backward; rows are D deep floats + 1 wide float)
        rk = step.tables.rk
        bnd = rk.bounds.tolist()
        n_u, own = bnd[world], bnd[rank + 1] - bnd[rank]
        n_r = int(rk.n_r.item())
        row_b = (EMB + 1) * 4
        out_b = (n_u - own) * 4 + (n_r - own) * row_b + (n_u - own) * row_b     # keys out, rows served out, grads out
        t_b = torch.tensor([float(out_b), float(n_u), float(n_r)], device=dev)
        dist.all_reduce(t_b, op=dist.ReduceOp.SUM)
        out_b, n_u_avg, n_r_avg = (float(x) / world for x in t_b.tolist())
        exchange = {"unique_keys_per_rank": int(n_u_avg), "rows_owned_per_rank": int(n_r_avg),
                    "nvlink_bytes_out_per_rank_per_step": int(out_b),
                    "avg_gbs_per_direction_over_step": round(out_b / (ms * 1e-3) / 1e9, 1), "nvlink_peak_gbs_per_direction": 900,
                    "note": "peer stores (keys, rows, gradients) leaving one GPU in one step / step time; the "
                            "exchange is latency- and dedup-bound, not link-bound"}
    exch_err = 0
    if world > 1 and hasattr(step.tables, "error_flags"):
        # device-driven exchange: bit 0 = a peer wait timed out, bit 1 = an inbox overflowed (either voids the run)
        t = torch.tensor([step.tables.error_flags()], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        exch_err = int(t.item())
        if exch_err:
            raise SystemExit("bench.py: sharded exchange raised error flags %d (1 = wait time-out, 2 = inbox overflow)" % exch_err)

    parity = None
    if world > 1 and hasattr(step.tables, "rk") and not args.no_extra_configs:
        # multi-GPU parity of the benchmarked exchange + BASELINE config 5 (every rank takes part)
        from tools import sharded_parity
        step.close()
        del step
        torch.cuda.empty_cache()
        group_one = dist.new_group([0])
        parity = {"fp32": sharded_parity.wide_deep(world, rank, dev, b, FIELDS, EMB, HIDDEN, mixed=False, alpha=args.alpha),
                  "fp16": sharded_parity.wide_deep(world, rank, dev, b, FIELDS, EMB, HIDDEN, mixed=True, alpha=args.alpha)}
        parity["ok"] = parity["fp32"]["ok"] and parity["fp16"]["ok"]
        configs["c5"] = measure_c5(args, world, rank, dev, group_one)
        c5p = configs["c5"].get("parity_check", {})
        bad = (not parity["ok"]) or any(not v.get("ok", False) for v in c5p.values()) or configs["c5"].get("error_flags", 0)
        if bad:
            if rank == 0:
                print(json.dumps({"parity_check": parity, "c5": configs["c5"]}), file=sys.stderr, flush=True)
            raise SystemExit("bench.py: multi-GPU parity check FAILED (see stderr)")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cb = None
    if world == 1 and not args.no_cpu_baseline:
        cb_full = cpu_baseline_run(args, steps=8, warmup=2)
        cb = {k: cb_full[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cb["config1"] = cpu_config1_run()
    if world == 1 and not args.no_extra_configs:
        # the other single-GPU BASELINE configs, measured by the same run (the tables of config 2 are released first)
        del step, model
        torch.cuda.empty_cache()
        n_extra = max(3, min(args.steps, 20))
        configs["torch_free"] = measure_torch_free(args, cards, host[:4], max(20, min(args.steps, 100)))
        import gc
        gc.collect()
        configs["c3"] = measure_c3(args, dev, n_extra)
        configs["c4"] = measure_c4(args, dev, n_extra)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms, "step_ms": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "clocks": clocks,
        "e2e": {"value": b * world / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
        "gpu_launches": int(per_step_launches * args.steps),
        "gpu_launches_per_step": int(per_step_launches),
        "roofline": roofline, "roofline_gather": roofline_gather, "configs": configs, "parity_check": parity,
        "cpu_baseline": cb, "breakdown_ms": breakdown, "exchange": exchange, "final_loss": final_loss,
        "lib": _lib.version(),
    }
    print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    # stdout carries the ONE JSON line and nothing else: libraries that print there (NCCL's version banner) are sent to
    # stderr by pointing fd 1 at fd 2 for the run; the line goes to the saved descriptor
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
