"""K1 parity: CUDA gather (through the aot C-ABI) vs the numpy oracle — bit-exact."""
import numpy as np
import pytest
import torch

from mindrec_b200 import _lib, ops
from oracle import ref_numpy as R

pytestmark = pytest.mark.gpu


def _table(v, d, seed=0):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((v, d)) * 0.01).astype(np.float32)


@pytest.mark.parametrize("dim", [4, 16, 32, 64, 80, 128, 100, 27, 1, 3])
@pytest.mark.parametrize("idt", [torch.int32, torch.int64])
def test_gather_matches_oracle(cuda, dim, idt):
    v, n = 5000, 3001  # ragged: not a multiple of any tile size
    tab = _table(v, dim)
    rng = np.random.default_rng(1)
    ids = rng.integers(0, v, size=n)
    out = ops.gather(torch.from_numpy(tab).to(cuda), torch.from_numpy(ids).to(cuda).to(idt))
    np.testing.assert_array_equal(out.cpu().numpy(), R.gather(tab, ids))


@pytest.mark.parametrize("dim", [16, 80, 128, 27])
def test_gather_masked_matches_oracle(cuda, dim):
    v, b, f = 20000, 777, 39
    tab = _table(v, dim, 2)
    rng = np.random.default_rng(3)
    ids = rng.integers(0, v, size=(b, f)).astype(np.int32)
    mask = rng.random((b, f)).astype(np.float32)
    mask[rng.random((b, f)) < 0.1] = 0.0
    out = ops.gather_masked(torch.from_numpy(tab).to(cuda), torch.from_numpy(ids).to(cuda),
                            torch.from_numpy(mask).to(cuda))
    assert out.shape == (b, f * dim)
    np.testing.assert_array_equal(out.cpu().numpy(), R.gather_masked(tab, ids, mask))


def test_gather_unaligned_ids_take_the_non_bulk_path(cuda):
    v, dim = 3000, 80
    tab = _table(v, dim, 4)
    ids = np.random.default_rng(5).integers(0, v, size=1000).astype(np.int32)
    d_ids = torch.from_numpy(np.concatenate([[0], ids]).astype(np.int32)).to(cuda)[1:]  # 4-B offset
    assert d_ids.data_ptr() % 16 != 0
    out = ops.gather(torch.from_numpy(tab).to(cuda), d_ids.contiguous() if False else d_ids)
    np.testing.assert_array_equal(out.cpu().numpy(), R.gather(tab, ids))


def test_gather_out_of_range_rows_are_zero_and_flagged(cuda):
    v, dim = 100, 80
    tab = _table(v, dim, 6)
    ids = np.array([0, 99, 100, -1, 5, 2**31 - 1], dtype=np.int32)
    flag = torch.zeros(1, dtype=torch.int32, device=cuda)
    out = ops.gather(torch.from_numpy(tab).to(cuda), torch.from_numpy(ids).to(cuda), oob_flag=flag)
    ref = R.gather(tab, ids)
    np.testing.assert_array_equal(out.cpu().numpy(), ref)
    assert not ref[2].any() and not ref[3].any()
    assert int(flag.item()) == 1
    flag.zero_()
    ops.gather(torch.from_numpy(tab).to(cuda), torch.from_numpy(ids[:2]).to(cuda), oob_flag=flag)
    assert int(flag.item()) == 0


def test_gather_empty(cuda):
    tab = torch.zeros((10, 80), device=cuda)
    out = ops.gather(tab, torch.zeros(0, dtype=torch.int32, device=cuda))
    assert out.shape == (0, 80)


def test_gather_reduce_matches_oracle(cuda):
    v, b, f = 20000, 1001, 39
    rng = np.random.default_rng(7)
    tab = (rng.standard_normal((v, 1)) * 0.01).astype(np.float32)
    ids = rng.integers(0, v, size=(b, f)).astype(np.int32)
    mask = rng.random((b, f)).astype(np.float32)
    bias = np.array([0.25], dtype=np.float32)
    out = ops.gather_reduce(torch.from_numpy(tab).to(cuda), torch.from_numpy(ids).to(cuda),
                            torch.from_numpy(mask).to(cuda), torch.from_numpy(bias).to(cuda))
    ref = R.gather_reduce(tab, ids, mask, bias)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-7)  # fp32 sum of 39 terms


def test_gather_full_size_roundtrip_property(cuda):
    """BASELINE config 2 lookup shape (16000 x 39, D=80) on a smaller vocab: gather(arange) is the
    identity and gather(perm)[inverse perm] restores the table (size-independent property)."""
    v, dim = 700000, 80
    tab = torch.randn((v, dim), device=cuda)
    perm = torch.randperm(v, device=cuda).to(torch.int32)
    g = ops.gather(tab, perm)
    inv = torch.empty_like(perm)
    inv[perm.long()] = torch.arange(v, device=cuda, dtype=torch.int32)
    back = ops.gather(g, inv)
    assert torch.equal(back, tab)
    ids = torch.randint(0, v, (16000, 39), device=cuda, dtype=torch.int32)
    out = ops.gather(tab, ids)
    assert torch.equal(out, tab[ids.long()])


def test_wrong_device_is_rejected(built_lib):
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.gather(torch.zeros((4, 4)), torch.zeros(2, dtype=torch.int32))


@pytest.mark.parametrize("dim", [16, 80, 27])
def test_gather_masked_fp16_output_equals_cast_after(cuda, dim):
    """fp16 `out` fuses DenseLayer's Cast(x, float16): bit-identical to rounding the fp32 product."""
    v, b, f = 9000, 333, 39
    tab = _table(v, dim, 8)
    rng = np.random.default_rng(9)
    ids = rng.integers(0, v, size=(b, f)).astype(np.int32)
    mask = rng.random((b, f)).astype(np.float32)
    out = ops.gather_masked(torch.from_numpy(tab).to(cuda), torch.from_numpy(ids).to(cuda),
                            torch.from_numpy(mask).to(cuda), out_dtype=torch.float16)
    assert out.dtype == torch.float16
    ref = R.gather_masked(tab, ids, mask).astype(np.float16)
    np.testing.assert_array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("b,s,dim", [(1, 1, 64), (333, 12, 64), (100, 7, 128), (64, 30, 16)])
def test_gather_pool_matches_oracle_and_backward_maps_to_sparse_update(cuda, b, s, dim):
    """Forward: fused mean-pooled lookup (multitable a15).  Backward: the pooled gradient broadcast over the S
    slots with weight mask/S is exactly mrec_segment_sum with div = S."""
    v = 4000
    rng = np.random.default_rng(b * 7 + s)
    tab = _table(v, dim, 11)
    ids = rng.integers(0, v, size=(b, s)).astype(np.int32)
    mask = (rng.random((b, s)) < 0.7).astype(np.float32)
    d_ids, d_mask = torch.from_numpy(ids).to(cuda), torch.from_numpy(mask).to(cuda)
    out = ops.gather_pool(torch.from_numpy(tab).to(cuda), d_ids, d_mask)
    ref = R.gather_pool(tab, ids, mask)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-6 * np.abs(tab).max())
    gout = rng.standard_normal((b, dim)).astype(np.float32)
    uq = ops.unique(d_ids)
    gs = ops.segment_sum(torch.from_numpy(gout).to(cuda), (d_mask / s).reshape(-1), uq, dim=dim)
    uniq, inverse, _, _ = R.unique_sorted(ids)
    gref = R.segment_sum(gout, inverse, uniq.size, (mask / s).reshape(-1), div=s)
    np.testing.assert_allclose(gs[:uniq.size].cpu().numpy(), gref, rtol=1e-5, atol=1e-5 * np.abs(gref).max())
