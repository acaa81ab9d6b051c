"""The reference's own CI tests for RecModel.online_train (ci/st/online_learning/test_online_learning.py:54-114),
restated against mindrec_b200.train.RecModel: three argument-validation cases with the same messages, plus the
loop's callback order.  CPU only (the toy net is a plain callable)."""
import numpy as np
import pytest

from mindrec_b200.train import Callback, RecModel


class StreamingDataset:
    """Same fixture as the reference: an endless stream of ones, batches of 100 x 39."""

    def __iter__(self):
        while True:
            yield (np.ones((100, 39), dtype=np.int32),)


class Net:
    def __init__(self):
        self.calls = 0

    def __call__(self, indices):
        self.calls += 1
        return indices.sum()


def test_online_learning_api_sink_size_is_negative():
    model = RecModel(Net(), device="cpu")
    with pytest.raises(ValueError) as exc_info:
        model.online_train(StreamingDataset(), dataset_sink_mode=True, sink_size=-1)
    assert "The input value must be int and must > 0" in str(exc_info.value)


def test_online_learning_api_sink_size_not_equal_one():
    model = RecModel(Net(), device="cpu")
    with pytest.raises(ValueError) as exc_info:
        model.online_train(StreamingDataset(), dataset_sink_mode=True, sink_size=100)
    assert "The sink_size parameter only support value of 1" in str(exc_info.value)


def test_online_learning_api_data_sink_mode_not_bool():
    model = RecModel(Net(), device="cpu")
    with pytest.raises(TypeError) as exc_info:
        model.online_train(StreamingDataset(), dataset_sink_mode="valid")
    assert "The input value must be a bool, but got str" in str(exc_info.value)


def test_online_train_runs_until_stopped_and_fires_hooks_in_order():
    events = []

    class Rec(Callback):
        def on_train_begin(self, c): events.append("begin")
        def on_train_epoch_begin(self, c): events.append("epoch_begin")
        def on_train_step_begin(self, c): events.append("step_begin")
        def on_train_step_end(self, c):
            events.append("step_end")
            if c.original_args().cur_step_num == 3:
                c.request_stop()
        def on_train_epoch_end(self, c): events.append("epoch_end")
        def on_train_end(self, c): events.append("end")

    net = Net()
    params = RecModel(net, device="cpu").online_train(StreamingDataset(), callbacks=Rec())
    assert net.calls == 3 and params.cur_step_num == 3 and params.cur_epoch_num == 1
    assert events == ["begin", "epoch_begin"] + ["step_begin", "step_end"] * 3 + ["epoch_end", "end"]
    assert float(params.net_outputs) == 3900.0


def test_merge_sliced_tables_mod_and_contiguous_layouts():
    """Per-rank slices of a row-sharded checkpoint merge back into the full tables (models/wide_deep/eval.py:86-107) for
    the owner = row mod G layout of this repository and for the reference's contiguous TABLE_ROW_SLICE split."""
    import torch
    from mindrec_b200 import train
    torch.manual_seed(0)
    vocab, dim, world = 103, 4, 4
    rows = (vocab + world - 1) // world
    full = {k: torch.randn(rows * world, 1 if "wide" in k or "ftrl" in k else dim) for k in train._SLICED}
    dense = {"dense_layers+Wide_b": torch.randn(37), "adam.moment1.dense": torch.randn(37), "adam.moment2.dense": torch.randn(37),
             "ftrl.hyper": torch.randn(16), "adam.hyper": torch.randn(16), "adam.hyper.dense": torch.randn(16)}
    for layout in ("mod", "contiguous"):
        slices = []
        for r in range(world):
            s = {k: (v[r::world] if layout == "mod" else v[r * rows:(r + 1) * rows]).clone() for k, v in full.items()}
            s.update({k: v.clone() for k, v in dense.items()})
            s["sharding"] = {"rank": r, "world": world, "vocab_size": vocab, "rows_per_rank": rows, "layout": layout}
            slices.append(s)
        merged = train.merge_sliced_tables(slices[::-1])
        for k, v in full.items():
            assert torch.equal(merged[k], v[:vocab])
        assert torch.equal(merged["dense_layers+Wide_b"], dense["dense_layers+Wide_b"])
        # re-sharding: merged state -> slices for another world size -> merged again
        for g2 in (1, 3, 8):
            again = train.merge_sliced_tables(train.split_tables(merged, g2, layout=layout))
            for k in train._SLICED + ("dense_layers+Wide_b", "adam.hyper"):
                assert torch.equal(again[k], merged[k])
    slices[1]["dense_layers+Wide_b"][0] += 1.0
    with pytest.raises(ValueError, match="differs between ranks"):
        train.merge_sliced_tables(slices)
    with pytest.raises(ValueError, match="one slice of every rank"):
        train.merge_sliced_tables(slices[:-1])
