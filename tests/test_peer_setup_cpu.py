"""Host logic of the device-driven exchange that needs no GPU: the collectively agreed failure of the peer-memory
set-up (2 ranks over gloo: one rank cannot allocate -> EVERY rank raises PeerMemoryUnavailable instead of one rank
hanging the others in a barrier) and the packed layout of the dedup result the look-ahead plan copies."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mindrec_b200 import ops, peer_sharded


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeArena:
    def __init__(self, fail_open):
        self.fail_open = fail_open

    def exchange(self):
        if self.fail_open:
            raise RuntimeError("cudaIpcOpenMemHandle failed (simulated)")
        return {}


class _FakeRank:
    def connect(self, base):
        self.base = base


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import unittest.mock as mock
        with mock.patch.object(torch.cuda, "synchronize", lambda *a, **k: None):
            cpu = torch.device("cpu")

            def bad_alloc():
                if rank == 1:
                    raise RuntimeError("mrec_peer_alloc failed (simulated)")
                return _FakeRank()
            for name, make, arena in (("alloc", bad_alloc, _FakeArena(False)),
                                      ("open", _FakeRank, _FakeArena(fail_open=(rank == 0))),
                                      ("ok", _FakeRank, _FakeArena(False))):
                try:
                    peer_sharded._connect_collectively(make, arena, None, cpu)
                    ret["%s%d" % (name, rank)] = "connected"
                except peer_sharded.PeerMemoryUnavailable as exc:
                    ret["%s%d" % (name, rank)] = "unavailable: %s" % exc
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_peer_memory_failure_is_agreed_by_all_ranks():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        ret = dict(ret)
    for r in range(world):
        assert ret["alloc%d" % r].startswith("unavailable"), ret      # rank 1 failed -> both ranks back off
        assert ret["open%d" % r].startswith("unavailable"), ret       # rank 0 failed -> both ranks back off
        assert ret["ok%d" % r] == "connected", ret


def test_packed_unique_result_is_one_buffer_with_aligned_fields():
    for n in (1, 5, 64, 1001):
        a = ops.UniqueResult(n, torch.int32, "cpu", packed=True)
        b = ops.UniqueResult(n, torch.int32, "cpu", packed=True)
        fields = a.outputs()
        assert [f.numel() for f in fields] == [n, n, 1, n, n + 1, n]
        base = a.flat.data_ptr()
        spans = sorted((f.data_ptr() - base, f.numel() * 4) for f in fields)
        for (o0, l0), (o1, _) in zip(spans, spans[1:]):
            assert o0 + l0 <= o1                                       # no overlap
        for f in (a.uniq, a.inverse, a.perm, a.seg_of, a.seg_start):
            assert (f.data_ptr() - base) % 16 == 0                     # vector loads in the kernels
        for i, f in enumerate(b.outputs()):
            f.fill_(i + 1)
        a.copy_from(b)                                                 # ONE copy moves every field
        for fa, fb in zip(a.outputs(), b.outputs()):
            assert torch.equal(fa, fb)
    c = ops.UniqueResult(7, torch.int64, "cpu", packed=True)          # int64 keys: not packed, field-wise copy
    assert c.flat is None
    d = ops.UniqueResult(7, torch.int64, "cpu")
    d.uniq.fill_(3)
    d.inverse.zero_(); d.count.fill_(1); d.perm.zero_(); d.seg_start.zero_(); d.seg_of.zero_()
    c.copy_from(d)
    assert int(c.uniq[0]) == 3 and int(c.count) == 1
