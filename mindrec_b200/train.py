"""RecModel.online_train — the unbounded-stream training loop of mindspore_rec.train.RecModel
(mindspore_rec/train/rec_model.py:118-309), restated without MindSpore's Model / dataset machinery, plus table
checkpoint export / import.

The loop itself holds no arithmetic: it pulls batches from an (infinite) iterable, stages them to the device
(one pinned H2D per tensor — the "dataset sink" of the reference) and calls the train network, firing the
callback hooks in the order of rec_model.py:279-308.  Argument validation and its error messages follow the
reference (they are the only behaviour its CI tests pin: ci/st/online_learning/test_online_learning.py:54-114).
"""
import sys

import torch


class Callback:
    """Hook points used by rec_model.py:279-308 (subset of mindspore.train.callback.Callback)."""

    def on_train_begin(self, run_context): pass
    def on_train_epoch_begin(self, run_context): pass
    def on_train_step_begin(self, run_context): pass
    def on_train_step_end(self, run_context): pass
    def on_train_epoch_end(self, run_context): pass
    def on_train_end(self, run_context): pass


class _CallbackParams:
    def __init__(self):
        self.train_network = None
        self.batch_num = 1
        self.cur_epoch_num = 0
        self.cur_step_num = 0
        self.dataset_sink_mode = True
        self.net_outputs = None


class RunContext:
    def __init__(self, params):
        self._params = params
        self._stop = False

    def original_args(self):
        return self._params

    def request_stop(self):
        self._stop = True

    def get_stop_requested(self):
        return self._stop


def _check_bool(value):
    if not isinstance(value, bool):
        raise TypeError("The input value must be a bool, but got %s." % type(value).__name__)
    return value


def _check_positive_int(value):
    if isinstance(value, bool) or not isinstance(value, int) or value <= 0:
        raise ValueError("The input value must be int and must > 0, but got %r." % (value,))
    return value


class RecModel:
    """RecModel(network, ...).online_train(train_dataset, callbacks=None, dataset_sink_mode=True, sink_size=1).

    `network` is a train-step callable (e.g. cells.TrainStepWrap): network(*batch) -> outputs."""

    def __init__(self, network, loss_fn=None, optimizer=None, metrics=None, eval_network=None, eval_indexes=None,
                 amp_level="O0", boost_level="O0", device="cuda"):
        if loss_fn is not None or optimizer is not None:
            raise ValueError("pass a train-step network that already contains loss and optimizers "
                             "(TrainStepWrap), as the reference's Wide&Deep scripts do")
        self._train_network = network
        self._device = torch.device(device)

    def online_train(self, train_dataset, callbacks=None, dataset_sink_mode=True, sink_size=1, max_steps=None):
        """rec_model.py:118-190.  `max_steps` (not in the reference, whose loop never ends) bounds the run
        for tests and benchmarks; a callback may also call run_context.request_stop()."""
        _check_bool(dataset_sink_mode)
        cbs = [] if callbacks is None else (list(callbacks) if isinstance(callbacks, (list, tuple)) else [callbacks])
        params = _CallbackParams()
        params.train_network = self._train_network
        if dataset_sink_mode:
            # rec_model.py:267-271
            sink_size = _check_positive_int(sink_size)
            if sink_size != 1:
                raise ValueError("The sink_size parameter only support value of 1 currently, but got: %d" % sink_size)
            params.batch_num = sink_size
        params.dataset_sink_mode = dataset_sink_mode
        ctx = RunContext(params)
        for cb in cbs:
            cb.on_train_begin(ctx)
        stop = False
        for epoch in range(sys.maxsize):                       # rec_model.py:285-287: unbounded epochs
            params.cur_epoch_num = epoch + 1
            for cb in cbs:
                cb.on_train_epoch_begin(ctx)
            produced = False
            for batch in train_dataset:
                produced = True
                params.cur_step_num += 1
                for cb in cbs:
                    cb.on_train_step_begin(ctx)
                inputs = tuple(self._stage(x) for x in batch)
                params.net_outputs = self._train_network(*inputs)
                for cb in cbs:
                    cb.on_train_step_end(ctx)
                if ctx.get_stop_requested() or (max_steps is not None and params.cur_step_num >= max_steps):
                    stop = True
                    break
            for cb in cbs:
                cb.on_train_epoch_end(ctx)
            if stop or not produced:
                break
        for cb in cbs:
            cb.on_train_end(ctx)
        return params

    def _stage(self, x):
        if isinstance(x, torch.Tensor):
            t = x
        else:
            t = torch.as_tensor(x)
        if t.device != self._device:
            if self._device.type == "cuda" and not t.is_pinned():
                t = t.pin_memory()
            t = t.to(self._device, non_blocking=True)
        return t


# --------------------------------------------------------------------------------------------------
# table checkpoints (ModelCheckpoint / load_param_into_net stand-in; MapParameter.export_data / import_data)
# --------------------------------------------------------------------------------------------------
def export_tables(step):
    """Flat (name -> CPU tensor) state of a Wide&Deep TrainStepWrap: tables, DenseLayers and optimizer state."""
    m = step.model
    out = {
        "wide_embeddinglookup.embedding_table": m.wide_embeddinglookup.embedding_table.data,
        "deep_embeddinglookup.embedding_table": m.deep_embeddinglookup.embedding_table.data,
        "dense_layers+Wide_b": m.dense.flat,
        "ftrl.accum": step.optimizer_w.accum[0], "ftrl.linear": step.optimizer_w.linear[0],
        "adam.moment1.table": step.optimizer_d.moment1[0], "adam.moment2.table": step.optimizer_d.moment2[0],
        "adam.moment1.dense": step.optimizer_d.moment1[1], "adam.moment2.dense": step.optimizer_d.moment2[1],
        "adam.hyper": step.optimizer_d.hyper, "ftrl.hyper": step.optimizer_w.hyper,
    }
    return {k: v.detach().cpu().clone() for k, v in out.items()}


def import_tables(step, state):
    """Inverse of export_tables (shapes must match)."""
    m = step.model
    dst = {
        "wide_embeddinglookup.embedding_table": m.wide_embeddinglookup.embedding_table.data,
        "deep_embeddinglookup.embedding_table": m.deep_embeddinglookup.embedding_table.data,
        "dense_layers+Wide_b": m.dense.flat,
        "ftrl.accum": step.optimizer_w.accum[0], "ftrl.linear": step.optimizer_w.linear[0],
        "adam.moment1.table": step.optimizer_d.moment1[0], "adam.moment2.table": step.optimizer_d.moment2[0],
        "adam.moment1.dense": step.optimizer_d.moment1[1], "adam.moment2.dense": step.optimizer_d.moment2[1],
        "adam.hyper": step.optimizer_d.hyper, "ftrl.hyper": step.optimizer_w.hyper,
    }
    missing = [k for k in dst if k not in state]
    if missing:
        raise KeyError("checkpoint is missing %s" % missing)
    for k, t in dst.items():
        if tuple(state[k].shape) != tuple(t.shape):
            raise ValueError("shape mismatch for %s: %s vs %s" % (k, tuple(state[k].shape), tuple(t.shape)))
        t.copy_(state[k])


# --------------------------------------------------------------------------------------------------
# row-sharded checkpoints: every rank saves its slice, evaluation merges them (models/wide_deep/eval.py:86-107:
# load_checkpoint per slice -> merge_sliced_parameter -> load_param_into_net)
# --------------------------------------------------------------------------------------------------
_SLICED = ("wide_embeddinglookup.embedding_table", "deep_embeddinglookup.embedding_table", "ftrl.accum", "ftrl.linear",
           "adam.moment1.table", "adam.moment2.table")


def export_sharded_tables(step):
    """This rank's slice of a row-sharded Wide&Deep step (sharded.ShardedWideDeepStep / peer_sharded.PeerSharded-
    WideDeepStep): its rows of the two tables and of their optimizer state, the (replicated) DenseLayers, and the
    sharding it was written under.  One file per rank, like the reference's per-rank checkpoints."""
    tb = step.tables
    rk = getattr(tb, "rk", tb)                        # PeerShardedTables keeps the arrays on its PeerRank
    plan = tb.plan
    out = {
        "wide_embeddinglookup.embedding_table": rk.wide, "deep_embeddinglookup.embedding_table": rk.deep,
        "ftrl.accum": rk.acc, "ftrl.linear": rk.lin, "adam.moment1.table": rk.m, "adam.moment2.table": rk.v,
        "dense_layers+Wide_b": step.dense.flat, "adam.moment1.dense": step.dense_m, "adam.moment2.dense": step.dense_v,
        "adam.hyper": rk.adam_hyper, "ftrl.hyper": rk.ftrl_hyper, "adam.hyper.dense": step.dense_hyper,
    }
    state = {k: v.detach().cpu().clone() for k, v in out.items()}
    state["sharding"] = {"rank": int(tb.rank), "world": int(tb.world), "vocab_size": int(plan.vocab_size),
                         "rows_per_rank": int(plan.rows_per_rank), "layout": "mod"}
    return state


def merge_sliced_tables(slices, layout=None):
    """Per-rank slices (export_sharded_tables, any order) -> one state in export_tables' format, loadable into the
    unsharded cell with import_tables (evaluation, serving, re-sharding to another G).

    layout "mod" (this repository: owner = row mod G, local index = row div G) interleaves the slices; layout
    "contiguous" (the reference's TABLE_ROW_SLICE: rank r holds rows [r R, (r+1) R), what merge_sliced_parameter
    concatenates) appends them.  Replicated entries (DenseLayers, their moments) are taken from rank 0 after checking
    that every rank holds the same values."""
    import torch
    if not slices:
        raise ValueError("merge_sliced_tables: no slices")
    meta = [s["sharding"] for s in slices]
    world = meta[0]["world"]
    if sorted(m["rank"] for m in meta) != list(range(world)) or any(m["world"] != world for m in meta):
        raise ValueError("merge_sliced_tables: need exactly one slice of every rank 0..%d" % (world - 1))
    layout = layout or meta[0].get("layout", "mod")
    if layout not in ("mod", "contiguous"):
        raise ValueError("layout must be 'mod' or 'contiguous'")
    by_rank = sorted(slices, key=lambda s: s["sharding"]["rank"])
    vocab, rows = meta[0]["vocab_size"], meta[0]["rows_per_rank"]
    out = {}
    for key in _SLICED:
        parts = [s[key] for s in by_rank]
        if any(p.shape[0] != rows for p in parts):
            raise ValueError("slice of %s does not have rows_per_rank = %d rows" % (key, rows))
        full = torch.stack(parts, 1).reshape(rows * world, -1) if layout == "mod" else torch.cat(parts, 0)
        out[key] = full[:vocab].contiguous()
    for key in ("dense_layers+Wide_b", "adam.moment1.dense", "adam.moment2.dense", "ftrl.hyper"):
        ref = by_rank[0][key]
        for s in by_rank[1:]:
            if not torch.equal(s[key], ref):
                raise ValueError("replicated entry %s differs between ranks" % key)
        out[key] = ref.clone()
    # the unsharded cell keeps one Adam hyper block for the table and the DenseLayers (same step count on both)
    out["adam.hyper"] = by_rank[0]["adam.hyper"].clone()
    return out


def split_tables(state, world, layout="mod"):
    """Inverse of merge_sliced_tables: a full-table state (export_tables / merge_sliced_tables) -> `world` per-rank
    slices in export_sharded_tables' format, e.g. to re-shard a checkpoint written on G GPUs for G' GPUs, or to hand a
    single-GPU checkpoint to a row-sharded job.  Rows beyond the vocabulary (padding of the last rank) are zero, with
    FTRL's accumulator padded by its own initial value so that a later merge round-trips."""
    import torch
    if layout not in ("mod", "contiguous"):
        raise ValueError("layout must be 'mod' or 'contiguous'")
    world = int(world)
    vocab = state["deep_embeddinglookup.embedding_table"].shape[0]
    rows = (vocab + world - 1) // world
    pad = rows * world - vocab
    out = []
    full = {}
    for key in _SLICED:
        t = state[key]
        t = t.reshape(vocab, -1)
        fill = 1.0 if key == "ftrl.accum" else 0.0
        full[key] = torch.cat([t, torch.full((pad, t.shape[1]), fill, dtype=t.dtype)], 0) if pad else t
    for r in range(world):
        s = {}
        for key in _SLICED:
            s[key] = (full[key][r::world] if layout == "mod" else full[key][r * rows:(r + 1) * rows]).contiguous().clone()
        for key in ("dense_layers+Wide_b", "adam.moment1.dense", "adam.moment2.dense", "ftrl.hyper", "adam.hyper"):
            s[key] = state[key].clone()
        s["adam.hyper.dense"] = state.get("adam.hyper.dense", state["adam.hyper"]).clone()
        s["sharding"] = {"rank": r, "world": world, "vocab_size": int(vocab), "rows_per_rank": int(rows), "layout": layout}
        out.append(s)
    return out
