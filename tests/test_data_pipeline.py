"""Input pipeline: the reference's TFRecord layout (1000 samples per record, columns feat_ids /
feat_vals / label) written by mindrec_b200.data.write_tfrecord and read back through the native decoder; sharding,
batching and the padding function follow models/wide_deep/src/datasets.py:175-270."""
import os
import struct

import numpy as np
import pytest

from mindrec_b200 import data, synth


def _write(tmp_path, name, n, seed, lps=100):
    gen = synth.CriteoSynth(n, cards=[50] * 26, vocab_pad=2000, seed=seed)
    ids, wts, label = gen.next()
    path = os.path.join(tmp_path, name)
    assert data.write_tfrecord(path, ids, wts, label, line_per_sample=lps) == n // lps
    return ids, wts, label


def test_roundtrip_is_exact(tmp_path, built_lib):
    ids, wts, label = _write(tmp_path, "train_0.tfrecord", 700, 1)
    ds = data.TFRecordDataset(str(tmp_path), train_mode=True, batch_size=200, line_per_sample=100, shuffle=False)
    assert len(ds) == 3                                           # 7 records, 2 per batch, remainder dropped
    got = list(ds)
    assert len(got) == 3
    for i, (x, y, z) in enumerate(got):
        sl = slice(i * 200, (i + 1) * 200)
        assert x.dtype == np.int32 and y.dtype == np.float32 and z.shape == (200, 1)
        np.testing.assert_array_equal(x, ids[sl])
        np.testing.assert_array_equal(y, wts[sl])
        np.testing.assert_array_equal(z, label[sl])


def test_negative_ids_and_unpacked_lists(tmp_path, built_lib):
    """Negative ids take the 10-byte varint form; unpacked (one tag per value) lists decode too."""
    ids = np.array([[-1, 5, 2 ** 31 - 1]], dtype=np.int32).repeat(2, 0)
    wts = np.array([[0.5, -2.0, 3.25]], dtype=np.float32).repeat(2, 0)
    lab = np.array([[1.0], [0.0]], dtype=np.float32)
    path = os.path.join(tmp_path, "train_neg.tfrecord")
    data.write_tfrecord(path, ids, wts, lab, line_per_sample=2)
    f = data.TFRecordFile(path)
    out = np.empty(6, np.int32)
    assert f.parse_into(0, "feat_ids", 0, out) == 6
    np.testing.assert_array_equal(out, ids.reshape(-1))
    # hand-built Example with UNPACKED float values: Feature{2: FloatList{1: fixed32, 1: fixed32}}
    vals = b"".join(b"\x0d" + struct.pack("<f", v) for v in (1.5, 2.5))
    feat = data._len_field(2, vals)
    entry = data._len_field(1, b"label") + data._len_field(2, feat)
    ex = data._len_field(1, data._len_field(1, entry))
    L = data._c()
    head = struct.pack("<Q", len(ex))
    with open(os.path.join(tmp_path, "train_unpacked.tfrecord"), "wb") as fh:
        fh.write(head + struct.pack("<I", L.mrec_crc32c_masked(head, 8)) + ex + struct.pack("<I", L.mrec_crc32c_masked(ex, len(ex))))
    g = data.TFRecordFile(os.path.join(tmp_path, "train_unpacked.tfrecord"))
    o = np.empty(4, np.float32)
    assert g.parse_into(0, "label", 1, o) == 2 and o[:2].tolist() == [1.5, 2.5]
    with pytest.raises(IOError, match="not found"):
        g.parse_into(0, "feat_ids", 0, out)
    with pytest.raises(IOError, match="another list type"):
        g.parse_into(0, "label", 0, out)


def test_corruption_is_detected(tmp_path, built_lib):
    _write(tmp_path, "train_0.tfrecord", 200, 2)
    path = os.path.join(tmp_path, "train_0.tfrecord")
    raw = bytearray(open(path, "rb").read())
    raw[40] ^= 0xff
    bad = os.path.join(tmp_path, "train_bad.tfrecord")
    open(bad, "wb").write(raw)
    with pytest.raises(IOError, match="CRC"):
        data.TFRecordFile(bad)
    assert data.TFRecordFile(bad, check_crc=False).n == 2
    open(bad, "wb").write(raw[:-7])
    with pytest.raises(IOError, match="framing"):
        data.TFRecordFile(bad)


def test_shards_are_equal_and_disjoint_and_shuffle_is_per_epoch(tmp_path, built_lib):
    ids, _, _ = _write(tmp_path, "train_a.tfrecord", 900, 3)
    _write(tmp_path, "train_b.tfrecord", 500, 4)
    _write(tmp_path, "test_a.tfrecord", 300, 5)                     # not a training file
    seen = []
    for r in range(3):
        ds = data.create_dataset(str(tmp_path), train_mode=True, batch_size=100, line_per_sample=100, rank_size=3,
                                 rank_id=r, seed=11)
        assert len(ds) == 4                                         # 14 records -> floor(14 / 3) per shard
        seen.append([tuple(x[0].tolist()) for x, _, _ in ds])
        e0 = [tuple(x[0].tolist()) for x, _, _ in ds]               # second epoch: another order, same shard size
        assert len(e0) == 4 and e0 != seen[-1]
    flat = [s for sh in seen for s in sh]
    assert len(set(flat)) == len(flat) == 12
    test = data.create_dataset(str(tmp_path), train_mode=False, batch_size=100, line_per_sample=100)
    assert len(test) == 3 and not test.shuffle


def test_padding_func_matches_reference(tmp_path, built_lib):
    """manual_shape: ids padded to target_column with (offset + size - 1) of the slice a column belongs to, weights
    with zeros (datasets.py:175-205)."""
    ids, wts, label = _write(tmp_path, "train_0.tfrecord", 200, 6)
    manual_shape = [(0, 1000), (1000, 500), (1500, 300), (1800, 200)]
    ds = data.create_dataset(str(tmp_path), batch_size=200, line_per_sample=100, manual_shape=manual_shape,
                             target_column=40, shuffle=False)
    (x, y, z), = list(ds)
    assert x.shape == (200, 40) and y.shape == (200, 40)
    np.testing.assert_array_equal(x[:, :39], ids)
    np.testing.assert_array_equal(y[:, :39], wts)
    assert np.all(x[:, 39] == 1999) and np.all(y[:, 39] == 0.0)      # column 39 -> part 3 -> 1800 + 200 - 1
    with pytest.raises(ValueError, match="multiple"):
        data.create_dataset(str(tmp_path), batch_size=150, line_per_sample=100)
    with pytest.raises(NotImplementedError, match="MindRecord"):
        data.create_dataset(str(tmp_path), data_type=data.DataType.MINDRECORD)


@pytest.mark.gpu
def test_device_loader_double_buffers(tmp_path, cuda):
    import torch
    ids, wts, label = _write(tmp_path, "train_0.tfrecord", 1200, 7)
    ds = data.TFRecordDataset(str(tmp_path), batch_size=200, line_per_sample=100, shuffle=False)
    loader = data.DeviceLoader(ds, cuda, depth=2, epochs=2)
    assert loader.h2d_bytes_per_batch == 200 * (39 * 8 + 4)
    got = [tuple(t.clone() for t in b) for b in loader]
    assert len(got) == 12
    for i, (x, y, z) in enumerate(got):
        sl = slice((i % 6) * 200, (i % 6 + 1) * 200)
        assert x.is_cuda and torch.equal(x.cpu(), torch.from_numpy(ids[sl]))
        assert torch.equal(y.cpu(), torch.from_numpy(wts[sl])) and torch.equal(z.cpu(), torch.from_numpy(label[sl]))
