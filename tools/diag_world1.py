"""One-rank triangulation: unsharded cell vs peer step (eager) vs peer step (graph, look-ahead) on the same batches."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindrec_b200 import cells, peer_sharded, synth

os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29577", RANK="0", WORLD_SIZE="1")
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
vocab, dim, b, fields, hidden = 50021, 16, 512, 39, (64, 32)
if len(sys.argv) > 1:
    vocab, dim, b, hidden = 2_000_003, 80, 16000, (1024, 512, 256, 128)
scale = vocab / synth.vocab_size(synth.CARD_KAGGLE)
cards = [max(3, int(c * scale * 0.9)) for c in synth.CARD_KAGGLE]      # every id in range
gen = synth.CriteoSynth(b, cards=cards, vocab_pad=vocab, seed=4)
batches = [tuple(torch.from_numpy(x).to(dev) for x in gen.next()) for _ in range(6)]


def run_peer(graph):
    step = peer_sharded.PeerShardedWideDeepStep(b, vocab, dim, hidden, dev, seed=3, use_mixed_precision=False, fields=fields, graph=graph)
    init = (step.tables.gather_full(), step.dense.flat.clone())
    step.capture(*batches[0], warmup=2)
    losses = []
    for s in range(1, 5):
        nxt = batches[s + 1] if (graph and s < 4) else None
        losses.append(float(step.replay(*batches[s], next_batch=nxt)[0]))
    torch.cuda.synchronize()
    out = (step.tables.gather_full(), step.dense.flat.clone(), losses, init)
    step.close()
    return out


(w_e, d_e), f_e, l_e, init = run_peer(False)
(w_g, d_g), f_g, l_g, _ = run_peer(True)
cfg = cells.WideDeepConfig(batch_size=b, field_size=fields, vocab_size=vocab, emb_dim=dim, deep_layer_dim=hidden,
                           use_mixed_precision=False, sparse=True, seed=9)
model = cells.WideDeepModel(cfg, device=dev)
model.wide_embeddinglookup.embedding_table.data.copy_(init[0][0])
model.deep_embeddinglookup.embedding_table.data.copy_(init[0][1])
model.dense.flat.copy_(init[1])
ref = cells.TrainStepWrap(cells.NetWithLossClass(model, cfg), sparse=True, lazy_adam=True)
for _ in range(2):
    ref(*batches[0])
l_r = [float(ref(*batches[s])[0]) for s in range(1, 5)]
print("loss eager", l_e)
print("loss graph", l_g)
print("loss ref  ", l_r)
print("eager == graph:", torch.equal(d_e, d_g), torch.equal(w_e, w_g), torch.equal(f_e, f_g))
for name, a, r in (("deep", d_e, model.embedding_table.data), ("wide", w_e, model.wide_embeddinglookup.embedding_table.data),
                   ("dense", f_e, model.dense.flat)):
    d = (a - r).abs().view(-1)
    top = torch.topk(d, 5)
    print(name, "max", float(d.max()), "n>1e-6", int((d > 1e-6).sum()), "of", d.numel(), "top idx", top.indices.tolist(),
          "vals", [(float(a.view(-1)[i]), float(r.view(-1)[i])) for i in top.indices.tolist()])
dist.destroy_process_group()
