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
