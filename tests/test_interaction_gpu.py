"""K7 / K8 parity: FM second-order and DCN cross stack (fwd + bwd) vs the numpy oracle, 1e-5 relative."""
import numpy as np
import pytest
import torch

from mindrec_b200 import ops
from oracle import ref_numpy as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("b,f,d", [(1, 39, 16), (257, 39, 16), (100, 39, 80), (33, 7, 6), (64, 39, 128), (5, 3, 1)])
def test_fm_forward_backward(cuda, b, f, d):
    rng = np.random.default_rng(b * 100 + d)
    vx = (rng.standard_normal((b, f, d)) * 0.5).astype(np.float32)
    g = rng.standard_normal((b, 1)).astype(np.float32)
    dvx_in = torch.from_numpy(vx).to(cuda)
    out = ops.fm_fwd(dvx_in)
    ref = R.fm_forward(vx)
    # fm is a difference of two O(sum vx^2) terms: fp32 tolerance is relative to that magnitude
    scale = 0.5 * np.square(vx.astype(np.float64)).sum(axis=(1, 2)).max()
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-5 * scale)
    dv = ops.fm_bwd(dvx_in, torch.from_numpy(g).to(cuda))
    rdv = R.fm_backward(vx, g)
    np.testing.assert_allclose(dv.cpu().numpy(), rdv, rtol=1e-5, atol=1e-5 * np.abs(rdv).max())


def test_fm_pairwise_identity_at_full_size(cuda):
    """BASELINE config 4 shape (16384 x 39 x 16): fm == sum_{i<j} <v_i, v_j> via a batched Gram matrix."""
    vx = torch.randn((16384, 39, 16), device=cuda) * 0.3
    out = ops.fm_fwd(vx).double().view(-1)
    gram = torch.bmm(vx.double(), vx.double().transpose(1, 2))
    ref = (gram.sum((1, 2)) - gram.diagonal(dim1=1, dim2=2).sum(1)) * 0.5
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("b,dp,layers", [(1, 3120, 6), (301, 3120, 6), (64, 1053, 6), (17, 40, 1), (50, 624, 3),
                                         (9, 8000, 8), (40, 2100, 2)])
def test_cross_stack_forward_backward(cuda, b, dp, layers):
    rng = np.random.default_rng(b + dp + layers)
    x0 = (rng.standard_normal((b, dp)) * 0.1).astype(np.float32)
    w = (rng.standard_normal((layers, dp)) * 0.05).astype(np.float32)
    bb = (rng.standard_normal((layers, dp)) * 0.05).astype(np.float32)
    gy = rng.standard_normal((b, dp)).astype(np.float32)
    dx0, dw_, db_ = (torch.from_numpy(a).to(cuda) for a in (x0, w, bb))
    y, p = ops.cross_fwd(dx0, dw_, db_)
    ry, _, rs = R.cross_forward(x0, w, bb)
    np.testing.assert_allclose(y.cpu().numpy(), ry, rtol=1e-5, atol=1e-5 * np.abs(ry).max())
    np.testing.assert_allclose(p.cpu().numpy(), x0.astype(np.float64) @ w.astype(np.float64).T, rtol=1e-4,
                               atol=1e-5 * np.abs(rs).max())
    gx, gw, gb = ops.cross_bwd(dx0, torch.from_numpy(gy).to(cuda), dw_, db_, p)
    rx, rw, rb = R.cross_backward(x0, w, bb, gy)
    for got, ref in ((gx, rx), (gw, rw), (gb, rb)):
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=1e-4, atol=2e-5 * np.abs(ref).max())


def test_cross_stack_equals_six_naive_layers_at_full_size(cuda):
    """BASELINE config 3 shape: fused stack vs 6 x (x0 * (xl @ w) + b + xl) in float64 on the device."""
    b, dp, layers = 16384, 3120, 6
    x0 = torch.randn((b, dp), device=cuda) * 0.1
    w = torch.randn((layers, dp), device=cuda) * 0.02
    bb = torch.randn((layers, dp), device=cuda) * 0.02
    y, _ = ops.cross_fwd(x0, w, bb)
    xl = x0.double()
    for l in range(layers):
        xl = x0.double() * (xl @ w[l].double())[:, None] + bb[l].double() + xl
    torch.testing.assert_close(y.double(), xl, rtol=1e-5, atol=1e-5 * float(xl.abs().max()))


def test_cross_backward_is_deterministic(cuda):
    b, dp, layers = 4000, 3120, 6
    x0 = torch.randn((b, dp), device=cuda) * 0.1
    w = torch.randn((layers, dp), device=cuda) * 0.02
    bb = torch.randn((layers, dp), device=cuda) * 0.02
    gy = torch.randn((b, dp), device=cuda)
    _, p = ops.cross_fwd(x0, w, bb)
    a = [t.clone() for t in ops.cross_bwd(x0, gy, w, bb, p)]
    c = [t.clone() for t in ops.cross_bwd(x0, gy, w, bb, p)]
    assert all(torch.equal(u, v) for u, v in zip(a, c))
