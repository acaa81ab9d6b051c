"""K2 parity: radix sort / unique vs the numpy oracle — bit-exact on every index output."""
import numpy as np
import pytest
import torch

from mindrec_b200 import ops
from oracle import ref_numpy as R

pytestmark = pytest.mark.gpu


def _check(uq, ids, bound=None):
    n = ids.size
    uniq, inverse, perm, seg_start = R.unique_sorted(ids, bound)
    u = int(uq.count.item())
    assert u == uniq.size
    np.testing.assert_array_equal(uq.uniq[:u].cpu().numpy(), uniq)
    np.testing.assert_array_equal(uq.inverse.cpu().numpy(), inverse)
    np.testing.assert_array_equal(uq.perm.cpu().numpy(), perm)
    np.testing.assert_array_equal(uq.seg_start[:u + 1].cpu().numpy(), seg_start)
    seg_of = np.repeat(np.arange(u), np.diff(seg_start)).astype(np.int32)
    np.testing.assert_array_equal(uq.seg_of.cpu().numpy(), seg_of)
    assert n == seg_start[-1]


@pytest.mark.parametrize("n", [1, 2, 31, 2048, 2049, 100003])
@pytest.mark.parametrize("idt", [np.int32, np.int64])
def test_unique_full_width(cuda, n, idt):
    rng = np.random.default_rng(n)
    ids = rng.integers(-50, 5000, size=n).astype(idt)
    if idt == np.int64 and n > 2:
        ids[0] = 2**40 + 7
        ids[1] = -(2**35)
    uq = ops.unique(torch.from_numpy(ids).to(cuda))
    _check(uq, ids)


@pytest.mark.parametrize("vocab", [100, 200000, 33762616])
def test_unique_bounded_zipf(cuda, vocab):
    rng = np.random.default_rng(vocab)
    n = 16000 * 39 if vocab > 1000 else 5000
    ids = (rng.zipf(1.05, size=n) % vocab).astype(np.int32)
    table_like = torch.empty((vocab, 0), device=cuda)
    uq = ops.unique(torch.from_numpy(ids).to(cuda), table_like=table_like)
    _check(uq, ids, bound=vocab)


def test_unique_bounded_collapses_out_of_range(cuda):
    vocab = 1000
    ids = np.array([5, 1000, 7, -3, 5, 2000, 999], dtype=np.int32)
    uq = ops.unique(torch.from_numpy(ids).to(cuda), table_like=torch.empty((vocab, 1), device=cuda))
    _check(uq, ids, bound=vocab)
    u = int(uq.count.item())
    assert uq.uniq[:u].cpu().tolist() == [5, 7, 999, 1000]


def test_unique_all_equal_and_all_distinct(cuda):
    ids = np.full(70000, 42, dtype=np.int32)
    _check(ops.unique(torch.from_numpy(ids).to(cuda)), ids)
    ids = np.random.default_rng(0).permutation(70000).astype(np.int32)
    _check(ops.unique(torch.from_numpy(ids).to(cuda)), ids)


def test_unique_empty(cuda):
    uq = ops.unique(torch.zeros(0, dtype=torch.int32, device=cuda))
    assert int(uq.count.item()) == 0


def test_unique_invariants_at_full_size(cuda):
    """Order-free invariants at BASELINE config 2 size: uniq[inverse] == ids, ascending, count."""
    ids = torch.randint(0, 33762616, (16000 * 39,), device=cuda, dtype=torch.int32)
    ids[:16000 * 13] = torch.arange(13, device=cuda, dtype=torch.int32).repeat(16000)
    uq = ops.unique(ids, table_like=torch.empty((33762616, 0), device=cuda))
    u = int(uq.count.item())
    assert torch.equal(uq.uniq[uq.inverse.long()], ids)
    assert bool((uq.uniq[1:u] > uq.uniq[:u - 1]).all())
    assert u == torch.unique(ids).numel()
    assert torch.equal(ids[uq.perm.long()], uq.uniq[uq.seg_of.long()])


@pytest.mark.parametrize("n", [1, 7, 5000, 70001])
def test_unique_first_occurrence_order(cuda, n):
    rng = np.random.default_rng(n)
    ids = rng.integers(0, max(2, n // 3), size=n).astype(np.int32)
    uniq, inverse, count = ops.unique_first(torch.from_numpy(ids).to(cuda))
    r_uniq, r_inv = R.unique_first(ids)
    u = int(count.item())
    assert u == r_uniq.size
    np.testing.assert_array_equal(uniq[:u].cpu().numpy(), r_uniq)
    np.testing.assert_array_equal(inverse.cpu().numpy(), r_inv)


def test_unique_first_docs_example(cuda):
    """MindSpore Unique docs example: [1,2,5,2] -> ([1,2,5],[0,1,2,1]) (holds for both orders)."""
    ids = torch.tensor([1, 2, 5, 2], dtype=torch.int32, device=cuda)
    uniq, inverse, count = ops.unique_first(ids)
    assert uniq[:int(count.item())].tolist() == [1, 2, 5] and inverse.tolist() == [0, 1, 2, 1]
    uq = ops.unique(ids)
    assert uq.uniq[:int(uq.count.item())].tolist() == [1, 2, 5] and uq.inverse.tolist() == [0, 1, 2, 1]


@pytest.mark.parametrize("n,bits", [(1_300_000, 21), (1_300_000, 26), (800_000, 22), (2_500_000, 24)])
def test_unique_large_tiles_and_waves(cuda, n, bits):
    """The pass kernels pick 8, 12 or 16 keys per thread from N (one wave of 148 tiles when possible, several waves
    beyond 1.2 M keys); every variant against the oracle, with and without a device-side valid count."""
    rng = np.random.default_rng(n + bits)
    vocab = (1 << bits) - 77
    ids = (rng.zipf(1.05, size=n) % vocab).astype(np.int32)
    ids[::7] = rng.integers(0, vocab, size=ids[::7].size)
    table_like = torch.empty((vocab, 0), device=cuda)
    d_ids = torch.from_numpy(ids).to(cuda)
    _check(ops.unique(d_ids, table_like=table_like), ids, bound=vocab)
    for nv in (n // 5, n - 3):                           # static inbox, device-side count (owner-side dedup)
        n_valid = torch.tensor([nv], dtype=torch.int32, device=cuda)
        uq = ops.unique(d_ids, table_like=table_like, n_valid=n_valid)
        uniq, inverse, perm, seg_start = R.unique_sorted(ids[:nv], vocab)
        u = int(uq.count.item())
        assert u == uniq.size
        np.testing.assert_array_equal(uq.uniq[:u].cpu().numpy(), uniq)
        np.testing.assert_array_equal(uq.inverse[:nv].cpu().numpy(), inverse)
        np.testing.assert_array_equal(uq.perm[:nv].cpu().numpy(), perm)
        np.testing.assert_array_equal(uq.seg_start[:u + 1].cpu().numpy(), seg_start)
