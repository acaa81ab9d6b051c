"""Input pipeline: the reference's Criteo TFRecord layout -> pinned host buffers -> device, double
buffered.

    create_dataset / DataType / TFRecordDataset      models/wide_deep/src/datasets.py:272-326,468-520
    one record = line_per_sample (1000) samples:      datasets/criteo_1tb/process_data.py:203-283
        feat_ids  int64 list [1000 * 39]   feat_vals  float list [1000 * 39]   label  float list [1000]
    a batch = batch_size / line_per_sample records -> (ids int32 [B, 39], weights float32 [B, 39], label float32 [B, 1]),
    optionally padded to `target_column` columns (the reference's _padding_func for host-device mode).

Framing and protobuf decoding are native (csrc/tfrecord.cu, host C entry points of libmindrec_b200.so reached through
ctypes on memory-mapped files): records are decoded straight into the (pinned) batch buffers.  `DeviceLoader` runs the
decode on a background thread and stages each batch to the device on its own copy stream, two batches ahead, so a
training loop that calls `next()` only ever waits for a copy that was issued during the previous step.
MindRecord (MindSpore's own container: sqlite index + paged blobs) and HDF5 are not read; `write_tfrecord` produces
the same record layout for tests and synthetic data.
"""
import ctypes
import mmap
import os
import queue
import struct
import threading

import numpy as np

from . import _lib


class DataType:
    """models/wide_deep/src/datasets.py:27-38."""
    MINDRECORD = 1
    TFRECORD = 2
    H5 = 3


_bound = False


def _c():
    global _bound
    L = _lib.lib()
    if not _bound:
        u8p = ctypes.c_void_p
        L.mrec_tfrecord_index.restype = ctypes.c_int64
        L.mrec_tfrecord_index.argtypes = [u8p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
        L.mrec_tfrecord_parse.restype = ctypes.c_int
        L.mrec_tfrecord_parse.argtypes = [u8p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
        L.mrec_crc32c_masked.restype = ctypes.c_uint32
        L.mrec_crc32c_masked.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        L.mrec_varint_pack.restype = ctypes.c_int64
        L.mrec_varint_pack.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64]
        _bound = True
    return L


# --------------------------------------------------------------------------------------------------
# writer (tests, synthetic data): the record layout of process_data.py
# --------------------------------------------------------------------------------------------------
def _varint(n):
    out = bytearray()
    while True:
        b = n & 0x7f
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _len_field(field, payload):
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _feature_entry(name, kind_field, packed):
    lst = _len_field(1, packed)                       # FloatList / Int64List { 1: packed values }
    feature = _len_field(kind_field, lst)             # Feature { 2: float_list | 3: int64_list }
    entry = _len_field(1, name.encode()) + _len_field(2, feature)
    return _len_field(1, entry)                       # Features { 1: map entry }


def encode_example(feat_ids, feat_vals, label):
    """One serialized tf.train.Example holding line_per_sample samples (flattened)."""
    L = _c()
    ids = np.ascontiguousarray(feat_ids, dtype=np.int32).reshape(-1)
    buf = np.empty(ids.size * 10, dtype=np.uint8)
    n = L.mrec_varint_pack(ids.ctypes.data, ids.size, buf.ctypes.data, buf.size)
    feats = (_feature_entry("feat_ids", 3, buf[:n].tobytes()) +
             _feature_entry("feat_vals", 2, np.ascontiguousarray(feat_vals, dtype="<f4").tobytes()) +
             _feature_entry("label", 2, np.ascontiguousarray(label, dtype="<f4").tobytes()))
    return _len_field(1, feats)                       # Example { 1: Features }


def write_tfrecord(path, ids, wts, label, line_per_sample=1000):
    """Write samples (ids [N,F] int32, wts [N,F] f32, label [N,1]) as N / line_per_sample records; a ragged tail is
    dropped, as the reference's writer does (process_data.py:262-266)."""
    L = _c()
    n = (ids.shape[0] // line_per_sample) * line_per_sample
    with open(path, "wb") as f:
        for r0 in range(0, n, line_per_sample):
            sl = slice(r0, r0 + line_per_sample)
            data = encode_example(ids[sl], wts[sl], label[sl])
            head = struct.pack("<Q", len(data))
            f.write(head)
            f.write(struct.pack("<I", L.mrec_crc32c_masked(head, 8)))
            f.write(data)
            f.write(struct.pack("<I", L.mrec_crc32c_masked(data, len(data))))
    return n // line_per_sample


# --------------------------------------------------------------------------------------------------
# reader
# --------------------------------------------------------------------------------------------------
class TFRecordFile:
    """A memory-mapped TFRecord file with its record index."""

    def __init__(self, path, check_crc=True):
        self.path = path
        self._f = open(path, "rb")
        size = os.fstat(self._f.fileno()).st_size
        self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ) if size else None
        self._view = np.frombuffer(self._mm, dtype=np.uint8) if size else np.empty(0, np.uint8)
        self.base = self._view.ctypes.data if size else 0
        L = _c()
        n = L.mrec_tfrecord_index(self.base, size, None, None, 0, 0)
        if n < 0:
            raise IOError("%s: corrupt TFRecord framing at byte %d" % (path, -n - 1))
        self.offsets = np.empty(max(n, 1), dtype=np.int64)
        self.lengths = np.empty(max(n, 1), dtype=np.int64)
        m = L.mrec_tfrecord_index(self.base, size, self.offsets.ctypes.data, self.lengths.ctypes.data, n, int(check_crc))
        if m < 0:
            raise IOError("%s: CRC mismatch in the record at byte %d" % (path, -m - 1))
        self.n = int(n)

    def parse_into(self, i, name, kind, out):
        """Decode feature `name` of record i into the numpy array `out` (int32 for kind 0, float32 for kind 1)."""
        cnt = ctypes.c_int64(0)
        rc = _c().mrec_tfrecord_parse(self.base + int(self.offsets[i]), int(self.lengths[i]), name.encode(), kind,
                                      out.ctypes.data, out.size, ctypes.byref(cnt))
        if rc != 0:
            raise IOError("%s record %d: %s" % (self.path, i, _lib.last_error()))
        return cnt.value


class TFRecordDataset:
    """ds.TFRecordDataset(files, schema(feat_ids, feat_vals, label), num_shards, shard_id, shard_equal_rows=True)
    .batch(batch_size / line_per_sample, drop_remainder=True).map(_padding_func) of datasets.py:220-270, as one iterable
    of (ids int32 [B, C], weights float32 [B, C], label float32 [B, 1]) numpy batches (C = 39, or target_column when
    manual_shape is given)."""

    def __init__(self, data_dir, train_mode=True, batch_size=1000, line_per_sample=1000, rank_size=None, rank_id=None,
                 manual_shape=None, target_column=40, field_size=39, shuffle=None, seed=0, check_crc=True):
        if batch_size % line_per_sample:
            raise ValueError("batch_size must be a multiple of line_per_sample (%d)" % line_per_sample)
        prefix = "train" if train_mode else "test"
        files = []
        if isinstance(data_dir, (list, tuple)):
            files = list(data_dir)
        else:
            for dirpath, _, names in os.walk(data_dir):                      # datasets.py:237-241
                files += [os.path.join(dirpath, n) for n in names if prefix in n and "tfrecord" in n]
        if not files:
            raise FileNotFoundError("no %s*tfrecord* file under %r" % (prefix, data_dir))
        self.files = [TFRecordFile(p, check_crc) for p in sorted(files)]
        self.batch_size, self.lps, self.field_size = batch_size, line_per_sample, field_size
        self.records_per_batch = batch_size // line_per_sample
        self.shuffle = train_mode if shuffle is None else shuffle
        self.seed, self.epoch = seed, 0
        index = [(fi, ri) for fi, f in enumerate(self.files) for ri in range(f.n)]
        self._index = index
        self.rank_size, self.rank_id = rank_size, rank_id
        if (rank_size is None) != (rank_id is None):
            raise ValueError("rank_size and rank_id go together")
        if rank_size is not None and not 0 <= rank_id < rank_size:
            raise ValueError("rank_id must be in [0, rank_size)")
        # _padding_func (datasets.py:175-215)
        self.columns = field_size
        self._fill = None
        if manual_shape:
            offs = [item[0] + item[1] for item in manual_shape]
            part = int(target_column / len(offs))
            self._fill = np.asarray([offs[i // part] - 1 for i in range(field_size, target_column)], dtype=np.int32)
            self.columns = target_column

    def _epoch_records(self):
        order = np.arange(len(self._index))
        if self.shuffle:
            np.random.default_rng(self.seed + self.epoch).shuffle(order)
        if self.rank_size is not None:                     # shard_equal_rows: every shard gets floor(n / G) records
            per = len(order) // self.rank_size
            order = order[self.rank_id::self.rank_size][:per]
        return order

    def __len__(self):
        n = len(self._index) if self.rank_size is None else len(self._index) // self.rank_size
        return n // self.records_per_batch                 # drop_remainder=True

    def alloc_batch(self, pinned=False):
        b, c = self.batch_size, self.columns
        if pinned:
            import torch
            t = (torch.empty((b, c), dtype=torch.int32).pin_memory(), torch.empty((b, c), dtype=torch.float32).pin_memory(),
                 torch.empty((b, 1), dtype=torch.float32).pin_memory())
            return t, tuple(x.numpy() for x in t)
        a = (np.empty((b, c), np.int32), np.empty((b, c), np.float32), np.empty((b, 1), np.float32))
        return a, a

    def fill_batch(self, records, arrays):
        """Decode `records` (indices into the global record list) into the batch arrays."""
        ids, wts, lab = arrays
        f, lps = self.field_size, self.lps
        padded = self.columns != f
        tmp_i = np.empty(lps * f, np.int32) if padded else None
        tmp_w = np.empty(lps * f, np.float32) if padded else None
        for j, gi in enumerate(records):
            fi, ri = self._index[gi]
            rec = self.files[fi]
            rows = slice(j * lps, (j + 1) * lps)
            if padded:
                ni, nw = rec.parse_into(ri, "feat_ids", 0, tmp_i), rec.parse_into(ri, "feat_vals", 1, tmp_w)
                ids[rows, :f] = tmp_i.reshape(lps, f)
                wts[rows, :f] = tmp_w.reshape(lps, f)
                ids[rows, f:] = self._fill[None, :]
                wts[rows, f:] = 0.0
            else:   # a record's 1000 x 39 values are exactly the rows of the (contiguous) batch slice
                ni = rec.parse_into(ri, "feat_ids", 0, ids[rows].reshape(-1))
                nw = rec.parse_into(ri, "feat_vals", 1, wts[rows].reshape(-1))
            nl = rec.parse_into(ri, "label", 1, lab[rows].reshape(-1))
            if ni != lps * f or nw != lps * f or nl != lps:
                raise IOError("%s record %d: expected %d x %d values, got ids %d vals %d labels %d"
                              % (rec.path, ri, lps, f, ni, nw, nl))

    def batches(self):
        """Record-index lists of this epoch's batches."""
        order = self._epoch_records()
        k = self.records_per_batch
        return [order[i:i + k] for i in range(0, len(order) - k + 1, k)]

    def __iter__(self):
        for recs in self.batches():
            out, arrays = self.alloc_batch()
            self.fill_batch(recs, arrays)
            yield out
        self.epoch += 1


def create_dataset(data_dir, train_mode=True, batch_size=1000, data_type=DataType.TFRECORD, line_per_sample=1000,
                   rank_size=None, rank_id=None, manual_shape=None, target_column=40, **kw):
    """datasets.py:468-520 (same argument names)."""
    if data_type == DataType.TFRECORD:
        return TFRecordDataset(data_dir, train_mode, batch_size, line_per_sample, rank_size, rank_id, manual_shape,
                               target_column, **kw)
    if data_type == DataType.MINDRECORD:
        raise NotImplementedError("MindRecord is MindSpore's own container format; convert with the reference's "
                                  "process_data.py --file_type tfrecord, or write_tfrecord()")
    raise NotImplementedError("only DataType.TFRECORD is read (got %r)" % (data_type,))


class DeviceLoader:
    """Decode on a background thread into a ring of pinned host batches and copy each to the device on a dedicated
    copy stream, `depth` batches ahead of the consumer.  Iterating yields (ids, weights, label) device tensors; the
    consumer's stream is made to wait for the copy's event, the host never blocks on a copy that was issued early
    enough.  A yielded batch stays valid until `depth` further batches have been taken."""

    def __init__(self, dataset, device, depth=2, epochs=1):
        import torch
        self.ds, self.device, self.depth, self.epochs = dataset, torch.device(device), max(2, int(depth)), epochs
        self._torch = torch
        n = self.depth + 2
        self._host = [dataset.alloc_batch(pinned=True) for _ in range(n)]
        self._dev = [tuple(torch.empty_like(t, device=self.device) for t in self._host[i][0]) for i in range(n)]
        self._copy = torch.cuda.Stream(device=self.device)
        self._free = queue.Queue()
        for i in range(n):
            self._free.put(i)
        self._ready = queue.Queue(maxsize=n)
        self._thread = None
        self._error = None
        self.h2d_bytes_per_batch = sum(t.numel() * t.element_size() for t in self._host[0][0])

    def _produce(self):
        try:
            for _ in range(self.epochs):
                for recs in self.ds.batches():
                    slot = self._free.get()
                    if slot is None:
                        return
                    self.ds.fill_batch(recs, self._host[slot][1])
                    self._ready.put(slot)
                self.ds.epoch += 1
        except Exception as exc:                           # noqa: BLE001 - re-raised in the consumer
            self._error = exc
        self._ready.put(None)

    def __iter__(self):
        torch = self._torch
        self._thread = threading.Thread(target=self._produce, daemon=True)
        self._thread.start()
        inflight = []                                      # (slot, event) staged to the device, oldest first
        done = False
        held = []
        while True:
            while not done and len(inflight) < self.depth:
                slot = self._ready.get()
                if slot is None:
                    done = True
                    break
                with torch.cuda.stream(self._copy):
                    for d, h in zip(self._dev[slot], self._host[slot][0]):
                        d.copy_(h, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                inflight.append((slot, ev))
            if not inflight:
                break
            slot, ev = inflight.pop(0)
            torch.cuda.current_stream().wait_event(ev)
            held.append((slot, ev))
            if len(held) > 1:                              # the batch handed out before the previous one is reusable:
                old, old_ev = held.pop(0)                  # its consumer kernels were enqueued before this point
                done_ev = torch.cuda.Event()
                done_ev.record()
                self._copy.wait_event(done_ev)             # device buffer: next copy into it waits for its readers
                old_ev.synchronize()                       # pinned buffer: its copy (issued 2 batches ago) has left
                self._free.put(old)
            yield self._dev[slot]
        self._free.put(None)
        if self._error is not None:
            raise self._error
