"""Host (enqueue) time vs device time of the sharded step's phases, run under torchrun."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindrec_b200 import sharded, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cards = [max(3, int(c * world)) for c in synth.CARD_KAGGLE]
step = sharded.ShardedWideDeepStep(16000, synth.vocab_size(cards), 80, (1024, 512, 256, 128), dev, seed=1)
gen = synth.CriteoSynth(16000, cards=cards, seed=20260101, rank=rank)
batches = [tuple(torch.from_numpy(x).to(dev) for x in gen.next()) for _ in range(8)]
step.capture(*batches[0], warmup=3)
T = {}
def timed(name, fn, *a, **kw):
    t0 = time.perf_counter(); r = fn(*a, **kw); T[name] = T.get(name, 0.0) + time.perf_counter() - t0; return r
tb = step.tables
orig_lookup, orig_update, orig_dense, orig_du, orig_plan = tb.lookup, tb.update, step._run_dense, step._dense_update, tb.plan_batch
tb.lookup = lambda *a: timed("lookup", orig_lookup, *a)
tb.update = lambda *a: timed("update", orig_update, *a)
step._run_dense = lambda: timed("dense_graph", orig_dense)
step._dense_update = lambda: timed("allreduce_adam", orig_du)
tb.plan_batch = lambda *a, **k: timed("plan", orig_plan, *a, **k)
n = 100
for i in range(10): step.replay(*batches[i % 8], next_batch=batches[(i + 1) % 8])
torch.cuda.synchronize(); T.clear()
t0 = time.perf_counter()
for i in range(n): step.replay(*batches[i % 8], next_batch=batches[(i + 1) % 8])
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
if rank == 0:
    print("host enqueue per step %.3f ms, wall per step %.3f ms" % (1e3 * t_host / n, 1e3 * t_all / n))
    for k, v in T.items(): print("  host %-16s %.3f ms/step" % (k, 1e3 * v / n))
dist.destroy_process_group()
